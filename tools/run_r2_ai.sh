#!/bin/bash
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
L=gpurun_out/r2ai.log; : > $L
timeout 900 python - >> $L 2>&1 <<PY
import sys, torch, numpy as np
sys.path.insert(0, '.')
import bench
from tools import workloads as wl
for name, mk, nom, B in (('c4', lambda: wl.c4_project()[0], 'mapk_kholodenko', 65536), ('c5', lambda: wl.c5_project()[0], 'egfr_signalling', 131072)):
    proj = mk()
    sm, nominal, _ = wl.nominal_of(nom)
    th = torch.from_numpy(wl.make_thetas(nominal, B, 5, 0.1)).cuda()
    (ms_i, ms_l), cnt, failed = bench._time_eval(proj, th, False, 'dp5', 2)
    print(name, 'states-only ms', round(ms_i, 2), 'solves/s', round(B / ms_i * 1e3), 'steps/solve', cnt['steps'] / B, 'traj', cnt['trajectories'], 'failed', failed)
import sysbio_modeling_b200 as sb
from sysbio_modeling_b200 import model_files
om = sb.OdeModel.from_model_file(model_files.path('michaelis_menten'), model_name='mm')
P = np.array([1e-3, 0.001, 0.01, 0.01, 0.001]) * np.exp(0.3 * np.random.default_rng(8).standard_normal((262144, 5)))
t = np.linspace(0, 500.0, 20)
for _ in range(2):
    y = om.simulate_batch(P, t, np.array([1.0, 0.0]))
print('mm 262144 ms', round(om.last_solve_ms, 2), 'steps', om.last_counters['steps'] / 262144)
PY
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_bench_parity.py tests/test_ensembles.py tests/test_gpu_full_size.py -m gpu -q -x 2>&1 | tail -3 >> $L
cat $L
