#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2aj.log; : > $L
for v in 24 25; do
echo "== SBM_PUSH_YREG_MAX=$v" >> $L
SBM_PUSH_YREG_MAX=$v timeout 900 python - >> $L 2>&1 <<PY
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from tools import workloads as wl
proj = wl.c5_project()[0]
sm, nominal, _ = wl.nominal_of('egfr_signalling')
th = torch.from_numpy(wl.make_thetas(nominal, 131072, 5, 0.1)).cuda()
(ms_i, ms_l), cnt, failed = bench._time_eval(proj, th, False, 'dp5', 3)
print('c5 states-only ms', round(ms_i, 2), 'solves/s', round(131072 / ms_i * 1e3), 'failed', failed)
PY
done
cat $L
