"""Experiment driver: time the C4 sensitivity solve and print all 8 raw counters (with -DSBM_CT_PROF=1:
[3] = state-warp total clocks, [5] = sens-warp barrier-wait clocks, [6] = state-warp barrier-wait clocks)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
shape = tuple(int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else '0,0,0').split(','))
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-10
sm, nominal, y0 = bench.mapk_model_info()
proj = bench.build_project(tol, None, shape)
info = proj._model.compiled.info()
dev = torch.device('cuda')
theta = torch.from_numpy(bench.make_thetas(nominal, B, 4)).to(dev)
bufs = {}
proj.evaluate_device(theta, with_jacobian=True, out=bufs)
torch.cuda.synchronize()
opts = proj._opts('dp5')
best = 1e30
for k in range(3):
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
print('B', B, 'ms', round(best, 3), 'solves/s', round(B / best * 1e3), 'TF', round(flops / best / 1e9, 3),
      'steps', c[0], 'rej', c[1], 'failed', int((bufs['status'] != 0).sum()))
print('counters', c.tolist())
k = 'ct' if info['sens_layout'] == 2 else 'sens'
print({x: info[x] for x in info if k in x or x == 'sens_layout'})
if c[3] > 0 and info['sens_layout'] == 2:
    W = info['ct_warps']
    print('state warp: wait/total = %.3f ; sens warps: wait/total = %.3f' % (c[6] / c[3], c[5] / (c[3] * max(W - 1, 1))))
