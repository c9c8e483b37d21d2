#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2u.log; : > $L
timeout 1500 python -m pytest tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -30 >> $L
cat $L
