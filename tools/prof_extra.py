"""Launches for the ncu captures of the kernels added late in round 2: 'c3jac' = the Robertson project's scaled Jacobian through
sbm_rodas4_sens (131,072 sets), 'c2wide' = the C2 project (4,096 experiments, 2 sets) through the row-parallel loss kernels,
'egfr' = the EGFR-scale model's scaled Jacobian through the warp-group kernel sbm_dp5_sens (1,024 sets)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools import workloads as wl
which = sys.argv[1]
if which == 'c3jac':
    proj, _ = wl.c3_project()
    sm, nominal, _ = wl.nominal_of('robertson')
    th = torch.from_numpy(wl.make_thetas(nominal, 131072, 3, 0.5)).cuda()
elif which == 'c2wide':
    proj, _ = wl.c2_project(4096)
    th = torch.from_numpy(wl.c2_theta(4096, 2)).cuda()
else:
    proj, _ = wl.c5_project()
    sm, nominal, _ = wl.nominal_of('egfr_signalling')
    th = torch.from_numpy(wl.make_thetas(nominal, 1024, 5, 0.1)).cuda()
bufs = {}
for _ in range(2):
    proj.evaluate_device(th, with_jacobian=True, out=bufs)
torch.cuda.synchronize()
print(which, 'failed', int((bufs['status'] != 0).sum()))
