"""One evaluation of C3 (Robertson, RODAS4, residuals) and C5 (EGFR-scale, plain DP5, SSR) - the launches ncu profiles
for profiles/r02_*_rodas4_plain / dp5_plain (tools/run_ncu_r02.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tools import workloads as wl
which = sys.argv[1] if len(sys.argv) > 1 else 'c3'
B = int(sys.argv[2]) if len(sys.argv) > 2 else (262144 if which == 'c3' else 32768)
if which == 'c3':
    proj, _ = wl.c3_project()
    sm, nominal, _ = wl.nominal_of('robertson')
    th = torch.from_numpy(wl.make_thetas(nominal, B, 3, 0.5)).cuda()
    solver = 'rodas4'
else:
    proj, _ = wl.c5_project()
    sm, nominal, _ = wl.nominal_of('egfr_signalling')
    th = torch.from_numpy(wl.make_thetas(nominal, B, 5, 0.1)).cuda()
    solver = 'dp5'
bufs = {}
for _ in range(3):
    proj.evaluate_device(th, with_jacobian=False, solver=solver, out=bufs)
torch.cuda.synchronize()
print(which, B, 'failed', int((bufs['status'] != 0).sum()))
