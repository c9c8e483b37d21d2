"""Throughput of the other BASELINE configs (they are parity-test cases, not the bench line): C2 reversible MM with
sensitivities, C3 Robertson through the Rosenbrock kernel.  Device time of the integrator launch (CUDA events)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import sysbio_modeling_b200 as sb
from sysbio_modeling_b200 import model_files

out = {}
# ---- C2: reversible Michaelis-Menten (4 states, 4 parameters -> 20 augmented states), 16 output times
mod = model_files.module('reversible_mm')
om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
nominal = np.array([mod.nominal[k] for k in om.param_order])
t = np.concatenate([[0.0], np.linspace(0.25, 4, 16)])
for nb in (4096, 262144):
    P = nominal * np.exp(0.3 * np.random.default_rng(2).standard_normal((nb, 4)))
    om.calc_jacobian_batch(P[:64], t, mod.initial_conditions)
    best = 1e30
    for _ in range(3):
        om.calc_jacobian_batch(P, t, mod.initial_conditions)
        best = min(best, om.last_solve_ms)
    info = om.compiled.info()
    out['C2_reversible_mm_sens_%d' % nb] = {
        'solves_per_s': nb / best * 1e3, 'ms': best, 'failed': int((om.last_status != 0).sum()),
        'steps_per_solve': om.last_counters['steps'] / nb, 'rtol_atol': om.rtol,
        'layout': info['sens_layout'], 'warps': info['ct_warps'], 'cols': info['ct_cols'], 'regs': info['regs_dp5_ct']}
# ---- C3: Robertson (3 states), 12 log-spaced times, residual path only (states), RODAS4
mod = model_files.module('robertson')
om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=1e-6, atol=1e-10)
nominal = np.array([mod.nominal[k] for k in om.param_order])
t = np.concatenate([[0.0], np.logspace(-2, 3, 12)])
for nb in (65536, 1048576):
    P = nominal * np.exp(0.5 * np.random.default_rng(3).standard_normal((nb, 3)))
    om.simulate_batch(P[:64], t, mod.initial_conditions)
    best = 1e30
    for _ in range(3):
        y = om.simulate_batch(P, t, mod.initial_conditions)
        best = min(best, om.last_solve_ms)
    out['C3_robertson_rodas4_%d' % nb] = {
        'solves_per_s': nb / best * 1e3, 'ms': best, 'failed': int((om.last_status != 0).sum()),
        'steps_per_solve': om.last_counters['steps'] / nb, 'lu_per_solve': om.last_counters['lu'] / nb,
        'mass_conservation_max_err': float(np.abs(y.sum(axis=2) - 1.0).max()), 'rtol': 1e-6, 'atol': 1e-10}
print(json.dumps(out, indent=1))
