#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ae.log; : > $L
run() {
echo "== PUSH=$1 defines=$2" >> $L
SBM_PUSH_NV=${SBM_PUSH_NV:-5} SBM_PLAIN_PUSH=$1 SBM_KERNEL_DEFINES="$2" timeout 600 python - >> $L 2>&1 <<PY
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from tools import workloads as wl
proj, _ = wl.c5_project()
sm, nominal, _ = wl.nominal_of('egfr_signalling')
th = torch.from_numpy(wl.make_thetas(nominal, 131072, 5, 0.1)).cuda()
(ms_i, ms_l), cnt, failed = bench._time_eval(proj, th, False, 'dp5', 2)
print('C5 ms', round(ms_i, 2), 'solves/s', round(131072 / ms_i * 1e3), 'steps/solve', cnt['steps'] / 131072, 'failed', failed)
PY
}
run 1 ""
SBM_PUSH_NV=4 run 1 "-DSBM_PUSH_YREG"
cat $L
