"""GPU tests of the loss epilogue kernels (csrc/sbm_epilogue.cuh): the warp-per-set kernel with the TMA-staged Jacobian
tile against the CTA-per-set kernel (same arithmetic, different reduction order), the two-phase form (group sums ->
finish) the experiment shard uses against the fused call, and the per-set normal equations against numpy."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def c4_eval():
    import torch
    from tools import workloads as wl
    proj, _ = wl.c4_project(1e-8)
    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    theta = torch.from_numpy(wl.make_thetas(nominal, 777, 4, 0.25)).cuda()
    return proj, theta


def _clone(o):
    return {k: v.clone() for k, v in o.items()}


def test_tile_kernel_matches_generic_kernel(c4_eval):
    import torch
    proj, theta = c4_eval
    a = _clone(proj.evaluate_device(theta, True))
    proj.loss_generic_kernel = True
    try:
        b = _clone(proj.evaluate_device(theta, True))
    finally:
        proj.loss_generic_kernel = False
    torch.cuda.synchronize()
    assert torch.isfinite(a['jac']).all()
    # same formulas, different association of the products and sums (the warp-per-set kernel stages d / sigma^2 and
    # s / sigma^2 per row, multiplies by 1 / sigma, reduces over 32 lanes instead of 128 threads): agreement to rounding,
    # amplified where the scale-factor gradient cancels (j1 / p2 - 2 p1 j2 / p2^2)
    for k in ('resid', 'ssr', 'sf', 'sfgrad', 'jac', 'grad'):
        scale = b[k].abs().max().item()
        err = (a[k] - b[k]).abs().max().item()
        print(k, 'max abs diff %.3e of scale %.3e' % (err, scale))
        assert torch.allclose(a[k], b[k], rtol=1e-9, atol=1e-11 * scale), (k, err, scale)
    # residual-only path (no Jacobian tile)
    ar = _clone(proj.evaluate_device(theta, False))
    proj.loss_generic_kernel = True
    try:
        br = _clone(proj.evaluate_device(theta, False))
    finally:
        proj.loss_generic_kernel = False
    for k in ('resid', 'ssr', 'sf'):
        assert torch.allclose(ar[k], br[k], rtol=1e-9, atol=1e-11 * br[k].abs().max().item()), k


@pytest.mark.parametrize('generic', [False, True])
def test_two_phase_loss_equals_fused(c4_eval, generic):
    """sums -> finish == fused, and summing the sums of two row subsets (what the all-reduce of the experiment shard
    does) gives the sums of the whole."""
    import torch
    from sysbio_modeling_b200 import _lib
    proj, theta = c4_eval
    plan, t = proj.device_plan, proj.plan_tables
    flags = _lib.SBM_FLAG_LOSS_GENERIC if generic else 0
    B, R, P, G = theta.shape[0], t.n_rows, t.n_proj, t.n_groups
    ref = {}
    proj.loss_generic_kernel = generic
    try:
        ref = _clone(proj.evaluate_device(theta, True))
    finally:
        proj.loss_generic_kernel = False
    o = {}
    sims = torch.zeros((B, R), dtype=torch.float64, device='cuda')
    jac = torch.empty((B, R, P), dtype=torch.float64, device='cuda')
    status = torch.zeros((B, t.E), dtype=torch.int32, device='cuda')
    counters = torch.zeros(8, dtype=torch.int64, device='cuda')
    plan.solve_device(theta, proj._opts('dp5'), sims, jac, status, counters)
    sums = torch.zeros((B, G, 3 + 2 * P), dtype=torch.float64, device='cuda')
    plan.loss_sums_device(theta, sims, jac, sums, flags=flags)
    resid = torch.empty((B, R), dtype=torch.float64, device='cuda')
    ssr = torch.empty((B,), dtype=torch.float64, device='cuda')
    sf = torch.empty((B, G), dtype=torch.float64, device='cuda')
    sfgrad = torch.empty((B, G, P), dtype=torch.float64, device='cuda')
    grad = torch.empty((B, P), dtype=torch.float64, device='cuda')
    plan.loss_finish_device(theta, sims, jac, sums, resid, ssr, sf, sfgrad, grad, flags=flags)
    torch.cuda.synchronize()
    for k, v in (('resid', resid), ('ssr', ssr), ('sf', sf), ('sfgrad', sfgrad), ('jac', jac), ('grad', grad)):
        assert torch.equal(v, ref[k]), (k, (v - ref[k]).abs().max().item())


def test_normal_equations(c4_eval):
    import torch
    from sysbio_modeling_b200 import engine
    proj, theta = c4_eval
    o = proj.evaluate_device(theta, True)
    packed = engine.normal_equations_device(o['jac'], o['resid'])
    A, g, rr = engine.unpack_normal_equations(packed, 22)
    torch.cuda.synchronize()
    J, r = o['jac'].cpu().numpy(), o['resid'].cpu().numpy()
    A0 = np.einsum('bri,brj->bij', J, J)
    g0 = np.einsum('bri,br->bi', J, r)
    assert np.allclose(A.cpu().numpy(), A0, rtol=1e-12, atol=1e-12 * np.abs(A0).max())
    assert np.allclose(g.cpu().numpy(), g0, rtol=1e-12, atol=1e-12 * np.abs(g0).max())
    assert np.allclose(g.cpu().numpy(), o['grad'].cpu().numpy(), rtol=1e-11, atol=1e-12 * np.abs(g0).max())
    assert np.allclose(rr.cpu().numpy(), 2 * o['ssr'].cpu().numpy(), rtol=1e-12)
    # odd shapes: 1 column, 32 columns, rows not a multiple of anything
    for R, P in ((5, 1), (7, 32), (33, 13)):
        J = torch.randn((50, R, P), dtype=torch.float64, device='cuda')
        r = torch.randn((50, R), dtype=torch.float64, device='cuda')
        A, g, rr = engine.unpack_normal_equations(engine.normal_equations_device(J, r), P)
        assert torch.allclose(A, torch.einsum('bri,brj->bij', J, J), rtol=1e-12, atol=1e-12)
        assert torch.allclose(g, torch.einsum('bri,br->bi', J, r), rtol=1e-12, atol=1e-12)


def _eval_with_env(proj, theta, with_jac, wide):
    """One evaluation through the kernels that keep J in global memory: wide = '1' the row-parallel kernels
    (sbm_loss_wide_*), '0' the CTA-per-set kernel."""
    import os
    import torch
    old = os.environ.get('SBM_LOSS_WIDE')
    os.environ['SBM_LOSS_WIDE'] = wide
    proj.loss_generic_kernel = True
    try:
        o = _clone(proj.evaluate_device(theta, with_jac))
        torch.cuda.synchronize()
    finally:
        proj.loss_generic_kernel = False
        if old is None:
            del os.environ['SBM_LOSS_WIDE']
        else:
            os.environ['SBM_LOSS_WIDE'] = old
    return o


def _assert_close_outputs(a, b, keys):
    import torch
    for k in keys:
        if k not in a or a[k] is None or a[k].numel() == 0:
            continue
        scale = max(b[k].abs().max().item(), 1e-300)
        err = (a[k] - b[k]).abs().max().item()
        print(k, 'max abs diff %.3e of scale %.3e' % (err, scale))
        assert torch.allclose(a[k], b[k], rtol=1e-10, atol=1e-12 * scale), (k, err, scale)


def test_row_parallel_kernels_match_cta_per_set_kernel(c4_eval, golden):
    """The wide path (few sets, big Jacobians: C2 through the Project) on shapes the other kernels also take: C4 (two
    scale-factor groups, 22 columns -> 32 x 4 thread shape), the reference's two-measurement project with parameter
    priors, and the 1-state project with the log loss and a scale-factor prior."""
    import torch
    from fixtures import project_1state, project_mm_two_measures
    from sysbio_modeling_b200 import LogSquareLossFunction
    proj, theta = c4_eval
    keys = ('resid', 'ssr', 'sf', 'sfgrad', 'jac', 'grad', 'sims')
    _assert_close_outputs(_eval_with_env(proj, theta, True, '1'), _eval_with_env(proj, theta, True, '0'), keys)
    _assert_close_outputs(_eval_with_env(proj, theta, False, '1'), _eval_with_env(proj, theta, False, '0'), ('resid', 'ssr', 'sf'))
    rng = np.random.default_rng(3)
    p2, _ = project_mm_two_measures(golden)
    th = torch.from_numpy(np.log(0.01) + 0.2 * rng.standard_normal((5, p2.n_project_params))).cuda()
    _assert_close_outputs(_eval_with_env(p2, th, True, '1'), _eval_with_env(p2, th, True, '0'), keys)
    p3, _ = project_1state(loss=LogSquareLossFunction)
    p3.set_scale_factor_log_prior('Variable_1', np.log(3.5), 0.4)
    th = torch.from_numpy(np.log(0.01) + 0.2 * rng.standard_normal((3, p3.n_project_params))).cuda()
    _assert_close_outputs(_eval_with_env(p3, th, True, '1'), _eval_with_env(p3, th, True, '0'), keys)
    # a NaN simulation poisons the whole set (squared_loss_function.py:28-32), the other sets are untouched
    thb = th.clone()
    thb[1, 0] = float('nan')
    a, b = _eval_with_env(p3, thb, True, '1'), _eval_with_env(p3, thb, True, '0')
    assert torch.isinf(a['resid'][1]).all() and torch.isinf(b['resid'][1]).all()
    assert torch.isfinite(a['resid'][0]).all() and torch.isfinite(a['resid'][2]).all()


def test_c2_project_wide_jacobian_block_structure_and_gradient():
    """C2 through the Project at reduced size (256 experiments x 16 rows, 1,024 project columns: the shape the
    row-parallel kernels are chosen for by default), with and without the scale-factor group shared by every experiment:
    J^T r from the kernels equals J^T r formed from their outputs, an experiment's rows touch its own four columns only
    (no groups) and the default choice equals the CTA-per-set kernel."""
    import torch
    from tools import workloads as wl
    E = 256
    for shared in (False, True):
        proj, _ = wl.c2_project(E, shared_scale_factor=shared)
        th = torch.from_numpy(wl.c2_theta(E, 2)).cuda()
        o = _clone(proj.evaluate_device(th, True))
        torch.cuda.synchronize()
        assert o['jac'].shape == (2, 16 * E, 4 * E)
        g = torch.einsum('brc,br->bc', o['jac'], o['resid'])
        assert torch.allclose(o['grad'], g, rtol=1e-10, atol=1e-10 * g.abs().max().item())
        assert torch.allclose(o['ssr'], 0.5 * (o['resid'] ** 2).sum(1), rtol=1e-12)
        if not shared:
            assert torch.count_nonzero(o['jac'][:, :16, 4:]) == 0
            assert torch.count_nonzero(o['jac'][:, :16, :4]) > 0
        b = _eval_with_env(proj, th, True, '0')
        _assert_close_outputs(o, b, ('resid', 'ssr', 'sf', 'sfgrad', 'jac', 'grad') if shared else ('resid', 'ssr', 'jac', 'grad'))
