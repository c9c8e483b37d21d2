#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ad.log; : > $L
for v in 0 1; do
echo "== SBM_PLAIN_PUSH=$v" >> $L
SBM_PLAIN_PUSH=$v timeout 600 python - >> $L 2>&1 <<PY
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from tools import workloads as wl
proj, _ = wl.c5_project()
sm, nominal, _ = wl.nominal_of('egfr_signalling')
th = torch.from_numpy(wl.make_thetas(nominal, 131072, 5, 0.1)).cuda()
(ms_i, ms_l), cnt, failed = bench._time_eval(proj, th, False, 'dp5', 2)
print('C5 ms', ms_i, 'solves/s', 131072 / ms_i * 1e3, 'steps/solve', cnt['steps'] / 131072, 'rej', cnt['rejects'], 'rhs', cnt['rhs_evals'], 'failed', failed)
PY
done
SBM_PLAIN_PUSH=1 timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_ensembles.py -m gpu -q -x 2>&1 | tail -5 >> $L
cat $L
