#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2p.log; : > $L
for c in 1e-4 1e-3 1e-2 1e-1; do timeout 300 python tools/bench_ensemble.py 32768 6 $c >> $L 2>&1; done
cat $L
