#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2h.log; : > $L
echo "== LAG=1 16384" >> $L; timeout 120 python tools/prof_ct.py 16384 2>&1 | grep -E "^B |rror" >> $L
echo "== LAG=1 65536" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== LAG=0 65536" >> $L; SBM_CT_LAG=0 timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_bench_parity.py tests/test_gpu_parity.py tests/test_gpu_determinism.py -m gpu -q -x -s 2>&1 | grep -E "C4 \||passed|failed|Error|assert" | tail -12 >> $L
echo "== LAG=1 PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
cat $L
