#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12 > gpurun_out/r20_tests.log; tail -8 gpurun_out/r20_tests.log
python - <<'PY' > gpurun_out/r20_egfr.log 2>&1
import sys, numpy as np
sys.path.insert(0, '.')
import sysbio_modeling_b200 as sb
from sysbio_modeling_b200 import model_files
mod = model_files.module('egfr_signalling')
om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'))
nominal = np.array([mod.nominal[k] for k in om.param_order])
t = np.concatenate([[0.0], np.linspace(50, 600, 12)])
for nb in (4096, 131072):
    P = nominal * np.exp(0.1 * np.random.default_rng(5).standard_normal((nb, len(nominal))))
    om.simulate_batch(P[:64], t, mod.initial_conditions)
    best = 1e30
    for _ in range(3):
        om.simulate_batch(P, t, mod.initial_conditions); best = min(best, om.last_solve_ms)
    print('C5 egfr plain dp5', nb, 'ms', best, 'solves/s', nb / best * 1e3, 'steps/solve', om.last_counters['steps'] / nb, 'failed', int((om.last_status != 0).sum()))
PY
cat gpurun_out/r20_egfr.log
