"""Batched Levenberg-Marquardt: the CPU restatement is pinned by identities and by
scipy.optimize.leastsq (the reference's consumer, tests/test_Project.py:202-213, 351-357) on the reference's fixture
projects; the CUDA step (C ABI sbm_lm_step) and Project.lm_fit_batch are compared with it on the GPU."""
import numpy as np
import pytest

import sbm_oracle as O
from fixtures import project_mm_sum, project_1state, theta_by_name


def _problem(rng, n_rows, n_proj, scale=1.0):
    J = rng.standard_normal((n_rows, n_proj)) * np.exp(scale * rng.standard_normal(n_proj))[None, :]
    r = rng.standard_normal(n_rows)
    return J, r


def test_oracle_step_identities():
    rng = np.random.default_rng(0)
    for n_rows, n_proj in ((40, 22), (9, 3), (19, 5), (35, 32)):
        J, r = _problem(rng, n_rows, n_proj)
        # Gauss-Newton limit: lam -> 0 gives the least-squares solution of J delta = -r
        d0, pred0, gmax, info = O.lm_step(J, r, 0.0)
        ref = np.linalg.lstsq(J, -r, rcond=None)[0]
        assert info == 0 and np.allclose(d0, ref, rtol=1e-8, atol=1e-10)
        assert np.isclose(gmax, np.abs(J.T @ r).max())
        # on a linear problem the predicted reduction is the actual one, for any damping
        for lam in (1e-3, 1.0, 50.0):
            d, pred, _, info = O.lm_step(J, r, lam)
            assert info == 0
            actual = 0.5 * r @ r - 0.5 * np.sum((r + J @ d) ** 2)
            assert np.isclose(pred, actual, rtol=1e-10, atol=1e-13) and pred > 0
            assert np.linalg.norm(d) < np.linalg.norm(d0)            # damping shortens the step
    # a column without influence falls back to plain Levenberg damping instead of a singular system
    J, r = _problem(rng, 12, 4)
    J[:, 2] = 0.0
    d, pred, _, info = O.lm_step(J, r, 1e-3)
    assert info == 0 and d[2] == 0.0
    # rank deficient without damping, non-finite input
    assert O.lm_step(J, r, 0.0)[3] == 1
    J[0, 0] = np.nan
    assert O.lm_step(J, r, 1.0)[3] == 1


def test_oracle_fit_agrees_with_scipy_leastsq_on_the_mm_sum_project(golden):
    """Reference tests/test_Project.py:351-357 drives leastsq with the project's callbacks; the restated LM reaches the
    same objective from the same perturbed start (the parameters themselves are not identifiable on this data)."""
    from scipy.optimize import leastsq
    proj, oracle = project_mm_sum(golden)
    g = golden['project_mm_sum']
    names = sorted(oracle.project_param_idx, key=lambda n: oracle.project_param_idx[n]['Global'])
    th = theta_by_name(names, g['project_order'], g['points'][0]['theta'])
    pert = dict(zip(['vmax', 'km', 'k_synt_s', 'k_deg_s', 'k_deg_p'], [0.3, -0.2, 0.1, 0.25, -0.15]))
    x0 = th + np.array([pert[n] for n in names])
    xs, ier = leastsq(oracle.residuals, x0, Dfun=oracle.jacobian, ftol=1e-12, xtol=1e-12, maxfev=5000)
    ssr_scipy = oracle.sum_square_residuals(xs)
    x, ssr, n_iter, conv = O.lm_fit(oracle.residuals, oracle.jacobian, x0, max_iter=200)
    assert ssr < 1e-2 * oracle.sum_square_residuals(x0)
    assert ssr <= ssr_scipy * 1.05 + 1e-9
    assert n_iter < 200


@pytest.mark.gpu
def test_cuda_step_matches_the_restatement():
    import torch
    from sysbio_modeling_b200 import engine
    rng = np.random.default_rng(1)
    for n_rows, n_proj, B in ((40, 22, 257), (9, 3, 64), (19, 5, 33), (35, 32, 40), (6, 8, 5), (50, 17, 130)):
        Js, rs, lams = [], [], []
        for b in range(B):
            J, r = _problem(rng, n_rows, n_proj, 0.7)
            if b % 11 == 3:
                J[:, b % n_proj] = 0.0                       # a parameter without influence
            Js.append(J), rs.append(r), lams.append(10.0 ** rng.uniform(-6, 3))
        Js[1][2, 0] = np.nan                                 # non-finite input -> info 1, zero step
        if B > 4:
            Js[4][:, 1] = 0.0                                # a zero column ...
            lams[4] = 0.0                                    # ... and no damping -> not positive definite
        J_d = torch.from_numpy(np.array(Js)).cuda()
        r_d = torch.from_numpy(np.array(rs)).cuda()
        l_d = torch.from_numpy(np.array(lams)).cuda()
        delta, pred, gmax, info = engine.lm_step_device(J_d, r_d, l_d)
        torch.cuda.synchronize()
        delta, pred, gmax, info = delta.cpu().numpy(), pred.cpu().numpy(), gmax.cpu().numpy(), info.cpu().numpy()
        for b in range(B):
            d, p, gm, inf = O.lm_step(Js[b], rs[b], lams[b])
            assert info[b] == inf, (n_rows, n_proj, b)
            if inf == 0:
                # floating point: same arithmetic in a different summation order; the systems are conditioned up to
                # ~1e6 here, hence 1e-8 relative to the step's norm
                assert np.linalg.norm(delta[b] - d) <= 1e-8 * np.linalg.norm(d) + 1e-13, (n_rows, n_proj, b)
                assert abs(pred[b] - p) <= 1e-9 * abs(p) + 1e-14
                assert abs(gmax[b] - gm) <= 1e-12 * gm
            else:
                assert np.all(delta[b] == 0) and pred[b] == 0


@pytest.mark.gpu
def test_lm_fit_batch_against_the_restatement_and_scipy(golden):
    """Multistart fit of the Michaelis-Menten 'sum' project through the CUDA path: every start reaches the objective the
    CPU restatement and scipy's leastsq reach from the same start."""
    from scipy.optimize import leastsq
    proj, oracle = project_mm_sum(golden)
    g = golden['project_mm_sum']
    mine = [n.split(' ')[0] for n in proj.get_ordered_project_params()]
    th = theta_by_name(mine, g['project_order'], g['points'][0]['theta'])
    rng = np.random.default_rng(5)
    starts = th[None, :] + 0.25 * rng.standard_normal((5, len(th)))
    starts[0] = th + np.array([dict(zip(['vmax', 'km', 'k_synt_s', 'k_deg_s', 'k_deg_p'], [0.3, -0.2, 0.1, 0.25, -0.15]))[n]
                               for n in mine])
    out = proj.lm_fit_batch(starts, max_iter=200)
    ssr0 = np.array([proj.calc_sum_square_residuals(s) for s in starts])
    assert out['theta'].shape == starts.shape and np.all(out['ssr'] <= ssr0 * (1 + 1e-6) + 1e-10)
    perm = [oracle.project_param_idx[n]['Global'] for n in mine]

    def to_oracle(x):
        xo = np.zeros_like(x)
        xo[perm] = x
        return xo
    for b in range(len(starts)):
        # the device result is what the product's own objective says at the returned point (the residual-only call
        # runs the plain integrator, whose steps differ from the sensitivity integrator's: tolerance-level differences
        # in each residual, i.e. ~1e-11 absolute in SSR near the 4e-7 floor of this fixture)
        assert abs(out['ssr'][b] - proj.calc_sum_square_residuals(out['theta'][b])) <= 1e-6 * out['ssr'][b] + 1e-10
        x0 = to_oracle(starts[b])
        _, ssr_cpu, _, _ = O.lm_fit(oracle.residuals, oracle.jacobian, x0, max_iter=200)
        xs, ier = leastsq(oracle.residuals, x0, Dfun=oracle.jacobian, ftol=1e-12, xtol=1e-12, maxfev=5000)
        ssr_scipy = oracle.sum_square_residuals(xs)
        # flat, non-identifiable valley: compare objectives, not parameters
        assert out['ssr'][b] <= 1.05 * min(ssr_cpu, ssr_scipy) + 2e-7, (b, out['ssr'][b], ssr_cpu, ssr_scipy)


@pytest.mark.gpu
def test_lm_fit_batch_with_scale_factors_1state():
    """TestProject's one-state model with a scale-factor group (reference tests/test_Project.py:29-76): SSR only goes
    down, failed starts stay where they are."""
    proj, oracle = project_1state()
    th0 = np.log(np.array([0.01, 0.01, 0.001]))
    rng = np.random.default_rng(7)
    starts = th0[None, :] + 0.3 * rng.standard_normal((8, 3))
    ssr0 = np.array([proj.calc_sum_square_residuals(s) for s in starts])
    out = proj.lm_fit_batch(starts, max_iter=60)
    assert np.all(out['ssr'] <= ssr0 * (1 + 1e-6) + 1e-10), (out['ssr'], ssr0)
    assert np.all(out['n_iter'] >= 1), out
    for b in range(len(starts)):
        assert abs(out['ssr'][b] - proj.calc_sum_square_residuals(out['theta'][b])) <= 1e-6 * out['ssr'][b] + 1e-10
