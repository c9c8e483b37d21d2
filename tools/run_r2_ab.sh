#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ab.log; : > $L
timeout 1200 python -m pytest tests/test_gpu_dense_output.py -m gpu -q -x -s 2>&1 | tail -40 >> $L
cat $L
