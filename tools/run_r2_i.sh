#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2i.log; : > $L
echo "== LAG=1 65536" >> $L; timeout 300 python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_determinism.py tests/test_gpu_configs.py -m gpu -q -x -s 2>&1 | grep -E "C4 \||passed|failed|Error|assert" | tail -12 >> $L
echo "== LAG=1 PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
for v in "-DSBM_WHATIF_NO_SYNCWARP" "-DSBM_WHATIF_NO_SYNCWARP -DSBM_CT_BUF_ARG=1"; do for lag in 0 1; do
echo "== LAG=$lag $v 16384" >> $L; SBM_CT_LAG=$lag SBM_KERNEL_DEFINES="$v" timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror" >> $L
done; done
echo "== LAG=0 16384 (reference for the experiment)" >> $L; SBM_CT_LAG=0 timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror" >> $L
echo "== LAG=1 16384" >> $L; timeout 300 python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror" >> $L
cat $L
