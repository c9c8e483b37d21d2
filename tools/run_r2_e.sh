#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2e.log; : > $L
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_bench_parity.py tests/test_gpu_epilogue.py -m gpu -q -x 2>&1 | tail -8 >> $L
echo "== JTM=1 65536" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== JTM=0 65536" >> $L; SBM_CT_JTM=0 timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== JTM=1 PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
cat $L
