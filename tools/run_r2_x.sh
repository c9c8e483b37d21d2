#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2x.log; : > $L
run() { echo "== $1 | base=$2 shares=$3 | $4" >> $L; SBM_CT_HEAD_COLBASE=$2 SBM_CT_HEAD_SHARES=$3 SBM_KERNEL_DEFINES="$1" timeout 600 python tools/prof_ct.py ${4:-65536} 0,0,0 1e-10 2 2>&1 | grep -E "^B |rror|SBM_CT_PROF" >> $L; }
run "" 200 4 65536
run "-DSBM_CTL_HELP=1" 200 4 65536
run "-DSBM_CTL_HELP=1" 0 4 65536
run "-DSBM_CTL_HELP=1" 0 6 65536
run "-DSBM_CTL_HELP=1" 100 3 65536
run "-DSBM_CTL_HELP=1 -DSBM_CT_PROF=10" 0 4 16384
SBM_CT_HEAD_COLBASE=0 SBM_KERNEL_DEFINES="-DSBM_CTL_HELP=1" timeout 900 python -m pytest tests/test_gpu_bench_parity.py tests/test_gpu_determinism.py -m gpu -q -x 2>&1 | tail -4 >> $L
cat $L
