#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2al.log; : > $L
run() { echo "== $1 | $2" >> $L; SBM_KERNEL_DEFINES="$1" timeout 600 python tools/prof_ct.py ${2:-65536} 0,0,0 1e-10 2 2>&1 | grep -E "^B |rror|SBM_CT_PROF" >> $L; }
run "" 65536
run "" 65536
timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_determinism.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
