"""Device time of the batched Levenberg-Marquardt step (C ABI sbm_lm_step) on C4-shaped data, against the HBM roofline,
and its share of one LM iteration of the C4 project (solve + loss epilogue + step + bookkeeping)."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from sysbio_modeling_b200 import engine
import bench

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
out = {}
for R, P in ((40, 22), (16, 4), (35, 32)):
    g = torch.Generator(device='cuda').manual_seed(1)
    J = torch.randn((B, R, P), dtype=torch.float64, device='cuda', generator=g)
    r = torch.randn((B, R), dtype=torch.float64, device='cuda', generator=g)
    lam = torch.full((B,), 1e-2, dtype=torch.float64, device='cuda')
    ds = torch.zeros((B, P), dtype=torch.float64, device='cuda')
    bufs = engine.lm_step_device(J, r, lam, ds)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device='cuda')
    best = 1e30
    for _ in range(5):
        flush.fill_(1.0)                                     # J (461 MB at C4) exceeds L2 anyway
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        engine.lm_step_device(J, r, lam, ds, *bufs)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    nbytes = B * (8 * (R * (P + 1) + 1 + P) + 8 * (2 * P + 2) + 4)
    out['lm_step_%dx%d' % (R, P)] = {'param_sets': B, 'ms': best, 'algorithmic_GB': nbytes / 1e9, 'GB_per_s': nbytes / best / 1e6,
                                    'steps_per_s': B / best * 1e3, 'failed': int((bufs[3] != 0).sum())}
peaks = os.path.join(ROOT, 'MEASURED_PEAKS.json')
if os.path.exists(peaks):
    out['hbm_peak'] = json.load(open(peaks))
# one LM iteration of the C4 project
if len(sys.argv) > 2:
    sm, nominal, y0 = bench.mapk_model_info()
    proj = bench.build_project(1e-10, None, (0, 0, 0))
    nb = int(sys.argv[2])
    theta = bench.make_thetas(nominal, nb, 4)
    import time
    proj.lm_fit_batch(theta[:256], max_iter=1)
    torch.cuda.synchronize()
    t0 = time.time()
    res = proj.lm_fit_batch(theta, max_iter=3)
    torch.cuda.synchronize()
    out['c4_lm_fit_3_iterations'] = {'param_sets': nb, 'wall_s': time.time() - t0,
                                     'ssr_reduced': int((res['ssr'] < 1e300).sum())}
print(json.dumps(out, indent=1))
