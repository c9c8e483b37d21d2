"""Timing of the loss epilogue and the normal-equation kernel on C4's shape (65,536 sets of 40 x 22), CUDA events."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tools import workloads as wl
from sysbio_modeling_b200 import engine
B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
proj, _ = wl.c4_project(1e-6)
sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
theta = torch.from_numpy(wl.make_thetas(nominal, B, 4, 0.25)).cuda()
bufs = {}
proj.evaluate_device(theta, True, out=bufs)
jac0, sims0 = bufs['jac'].clone(), bufs['sims'].clone()
t = proj.plan_tables
best = 1e9
for _ in range(5):
    bufs['jac'].copy_(jac0); bufs['sims'].copy_(sims0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    proj.device_plan.loss_device(theta, bufs['sims'], bufs['jac'], bufs['resid'], bufs['ssr'], bufs['sf'], bufs['sfgrad'], bufs['grad'])
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
byts = B * (2 * t.n_rows * t.n_proj + 3 * t.n_rows + 3 * t.n_proj + 4) * 8
print('loss ms %.4f  algorithmic GB/s %.0f  (NBUF=%s WARPS=%s)' % (best, byts / best / 1e6, os.environ.get('SBM_LOSS_NBUF'), os.environ.get('SBM_LOSS_WARPS')))
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); engine.normal_equations_device(bufs['jac'], bufs['resid']); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print('gram ms %.4f' % best)
