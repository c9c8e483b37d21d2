"""Experiment driver: time the C4 sensitivity solve, print the counters (with SBM_KERNEL_DEFINES=-DSBM_CT_PROF=<warp>
the kernel prints the clocks of its barrier intervals).  SBM_PROF_MAXSTEPS bounds the steps per trajectory (what-if builds
whose results are wrong must still terminate)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from tools import workloads as wl
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
shape = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '0,0,0').split(','))
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
proj = wl._c4_project(tol, {k: np.ones_like(wl.C4_T) for k in wl.C4_OBS}, shape)
proj.max_steps = int(os.environ.get('SBM_PROF_MAXSTEPS', '0'))
info = proj._model.compiled.info()
dev = torch.device('cuda')
theta = torch.from_numpy(wl.make_thetas(nominal, B, 4, 0.25)).to(dev)
bufs = {}
proj.evaluate_device(theta, with_jacobian=True, out=bufs)
torch.cuda.synchronize()
opts = proj._opts('dp5')
best = 1e30
for k in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    bufs['counters'].zero_()
    e0.record()
    proj.device_plan.solve_device(theta, opts, bufs['sims'].zero_(), bufs['jac'], bufs['status'], bufs['counters'])
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
c = bufs['counters'].cpu().numpy().astype(np.int64)
cuda_src = proj._model.compiled.cuda
flops = c[2] * cuda_src.flops_sens_rhs + c[0] * 57 * 184
flops_survey = c[2] * 1004 + c[0] * 57 * 184
ck = float(bufs['jac'].double().nan_to_num().abs().sum().item())
print('B', B, 'ms', round(best, 3), 'solves/s', round(B / best * 1e3), 'TF', round(flops / best / 1e9, 3),
      'TF(1004)', round(flops_survey / best / 1e9, 3), 'steps', c[0], 'rej', c[1], 'failed', int((bufs['status'] != 0).sum()),
      'ns/step/32', round(best * 1e6 * 148 * 32 / max(c[0], 1), 1), 'jacsum', repr(ck))
k = 'ct' if info['sens_layout'] == 2 else 'sens'
print({x: info[x] for x in info if k in x or x == 'sens_layout'})
