#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2k.log; : > $L
echo "== LAG split 65536" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
timeout 900 python -m pytest tests/test_gpu_epilogue.py tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_determinism.py -m gpu -q -x -s 2>&1 | grep -E "max abs diff|C4 \||passed|failed|Error|assert" | tail -14 >> $L
cat $L
