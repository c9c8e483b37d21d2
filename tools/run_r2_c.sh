#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2c.log; : > $L
python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_configs.py -m gpu -q -s 2>&1 | grep -E "C4 \||passed|failed|Error|assert" | head -20 >> $L
echo "== default 65536" >> $L; python tools/prof_ct.py 65536 2>&1 | grep -E "^B |rror" >> $L
echo "== PROF" >> $L; SBM_KERNEL_DEFINES="-DSBM_CT_PROF=0" python tools/prof_ct.py 16384 0,0,0 1e-10 1 2>&1 | grep -E "^B |rror|SBM_CT_PROF" | tail -5 >> $L
cat $L
