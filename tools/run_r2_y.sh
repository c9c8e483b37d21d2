#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2y.log; : > $L
timeout 1200 python -m pytest tests/test_gpu_stiff_sens.py -m gpu -q -x -s 2>&1 | tail -50 >> $L
cat $L
