#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2am.log; : > $L
timeout 600 python tools/bench_egfr_sens.py 2048 >> $L 2>&1
SBM_KERNEL_DEFINES="-DSBM_JG_SMEM=0" timeout 600 python tools/bench_egfr_sens.py 2048 >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_ensembles.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
