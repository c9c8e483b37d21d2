"""Sensitivity solve of the EGFR-scale model (25 states, 52 parameters: beyond the CTA-transposed layout, so the warp-group
kernel sbm_dp5_sens runs): device time of one scaled-Jacobian evaluation of B parameter sets."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tools import workloads as wl
from sysbio_modeling_b200 import _lib
proj, _ = wl.c5_project()
sm, nominal, _ = wl.nominal_of('egfr_signalling')
info = proj._model.compiled.info()
print({k: info[k] for k in info if 'sens' in k or k in ('group_lanes', 'cols_per_lane', 'block')})
for B in [int(x) for x in (sys.argv[1:] or ['2048', '16384'])]:
    th = torch.from_numpy(wl.make_thetas(nominal, B, 5, 0.1)).cuda()
    bufs = {}
    proj.evaluate_device(th, True, out=bufs)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        proj.evaluate_device(th, True, out=bufs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    cnt = dict(zip(_lib.COUNTER_NAMES, [int(x) for x in bufs['counters'].cpu().numpy()[:5]]))
    cs = proj._model.compiled.cuda
    flops = cnt['rhs_evals'] * cs.flops_sens_rhs + cnt['steps'] * 57 * 25 * 53
    print('EGFR sens B', B, 'ms', round(best, 2), 'jacobian evals/s', round(B / best * 1e3), 'steps/solve', cnt['steps'] / B,
          'TFLOP/s', round(flops / best / 1e9, 2), 'failed', int((bufs['status'] != 0).sum()))
