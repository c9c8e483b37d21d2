#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/check2.log; : > $L
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/fp64_latency.cu && /tmp/fp64_latency >> $L 2>&1
G='^B |rror|SBM_CT_PROF model|SBM_CT_PROF column warp 0'
echo "== default (2 newton steps)" >> $L; python tools/prof_ct.py 65536 2>&1 | grep -E "$G" >> $L
echo "== prof" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" python tools/prof_ct.py 65536 2>&1 | grep -E "$G" | tail -3 >> $L
python -m pytest tests -m gpu -x -q -k "reciprocal or cta_transposed" 2>&1 | tail -3 >> $L
cat $L
