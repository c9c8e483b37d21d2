#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2aa.log; : > $L
timeout 600 python tools/bench_egfr_sens.py 2048 16384 >> $L 2>&1
cat $L
