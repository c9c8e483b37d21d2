"""Batched Metropolis ensemble sampler (reference project/Ensembles.py:55-178, 193-198, 226-270):
W walkers in lock-step on the GPU against the CPU restatement of ONE reference walker fed with the same random numbers."""
import numpy as np
import pytest

import sbm_oracle as O
from fixtures import project_1state


def test_sampling_matrix_is_the_reference_construction():
    from sysbio_modeling_b200.project import ensembles
    rng = np.random.default_rng(3)
    for n in (3, 5, 22):
        A = rng.standard_normal((2 * n, n)) * np.exp(2.0 * rng.standard_normal(n))[None, :]
        H = A.T @ A
        for cutoff, T, scale in ((0, 1.0, 1.0), (1e-3, 2.5, 0.5)):
            S = ensembles.sampling_matrix(H, cutoff, T, scale)
            assert np.allclose(S, O.sampling_matrix(H, cutoff, T, scale), rtol=1e-12, atol=0)
            if cutoff == 0:
                # without truncation the proposal covariance is T * scale^2 / n * (H/2)^-1 (Ensembles.py:243-256)
                assert np.allclose(S @ S.T, np.linalg.inv(0.5 * H) * T * scale ** 2 / n, rtol=1e-6)
    ac = ensembles.autocorrelation(np.sin(np.arange(200) / 5.0))
    assert ac[0] == 1.0 and ac[1] < 1.0


def _hessian(oracle, theta):
    J = oracle.jacobian(theta)
    return J.T @ J


@pytest.mark.gpu
@pytest.mark.parametrize('with_sf_prior', [False, True])
def test_walkers_follow_the_restated_reference_walker(with_sf_prior):
    """TestProject's one-state model (scale-factor group 'Variable_1', reference tests/test_Project.py:29-76).  With a
    log prior on the scale factor the free energy carries the SloppyCell entropy integral; without one the reference's
    entropy raises, so the project is built without scale factors."""
    from sysbio_modeling_b200.project import ensembles
    proj, oracle = project_1state()
    if with_sf_prior:
        proj.set_scale_factor_log_prior('Variable_1', np.log(3.5), 0.4)
        oracle.set_scale_factor_log_prior('Variable_1', np.log(3.5), 0.4)
    T = 2e-3 if with_sf_prior else 1e-3
    th0 = np.log(np.array([0.01, 0.01, 0.001]))
    W, steps = 6, 10
    rng = np.random.default_rng(11)
    normals = rng.standard_normal((steps, W, 3))
    uniforms = rng.random((steps, W))
    starts = th0[None, :] + 0.05 * rng.standard_normal((W, 3))
    H = _hessian(oracle, th0)
    if not with_sf_prior:
        with pytest.raises(TypeError):
            proj.free_energy(th0, T)                       # the reference's entropy needs a prior (None - float)
        return
    ens, fes, ratio = ensembles.ensemble_log_params_batch(proj, starts, hess=H, steps=steps, temperature=T,
                                                          step_scale=0.7, normals=normals, uniforms=uniforms)
    assert ens.shape == (steps + 1, W, 3) and fes.shape == (steps + 1, W) and ratio.shape == (W,)
    S = O.sampling_matrix(H, 0, T, 0.7)
    fe = lambda th: oracle.free_energy(th, T)
    moved = 0
    for w in range(W):
        e_ref, f_ref, r_ref = O.ensemble_walker(fe, starts[w], S, normals[:, w], uniforms[:, w], T)
        # the free energies agree to the integration tolerance (SSR ~1e-3 here; 1e-6 relative as for SSR elsewhere);
        # accept / reject follows unless a uniform number sits within that of the acceptance probability
        assert np.allclose(fes[:, w], f_ref, rtol=1e-6, atol=1e-9), (w, fes[:, w], f_ref)
        assert np.allclose(ens[:, w], e_ref, rtol=0, atol=1e-12)
        assert abs(ratio[w] - r_ref) < 1e-12
        moved += int(r_ref > 0)
    assert moved >= W // 2                                  # the chains actually move
    # the host API of one walker (Project.free_energy, reference :874-892) sees the same numbers
    assert abs(proj.free_energy(ens[-1, 0], T) - fes[-1, 0]) <= 1e-6 * abs(fes[-1, 0]) + 1e-9


@pytest.mark.gpu
def test_device_generator_and_stationarity_smoke():
    """Device random numbers: every walker keeps a finite free energy, acceptance ratios are sane, the same seed gives the
    same ensemble and another seed a different one."""
    from sysbio_modeling_b200.project import ensembles
    proj, oracle = project_1state()
    proj.set_scale_factor_log_prior('Variable_1', np.log(3.5), 0.4)
    th0 = np.log(np.array([0.01, 0.01, 0.001]))
    a = ensembles.ensemble_log_params_batch(proj, th0, steps=12, temperature=2e-3, seeds=5, n_walkers=64, skip_elems=1)
    b = ensembles.ensemble_log_params_batch(proj, th0, steps=12, temperature=2e-3, seeds=5, n_walkers=64, skip_elems=1)
    c = ensembles.ensemble_log_params_batch(proj, th0, steps=12, temperature=2e-3, seeds=6, n_walkers=64, skip_elems=1)
    assert a[0].shape == (7, 64, 3) and np.all(np.isfinite(a[1]))
    assert np.array_equal(a[0], b[0]) and not np.array_equal(a[0], c[0])
    assert 0.02 < a[2].mean() < 0.98


# ---- pinned against the real reference (tests/golden/make_golden_ensembles.py ran its function sources here) ----
def _golden_ens():
    import json
    import os
    return json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'reference_golden_ensembles.json')))


def test_sampling_matrix_and_autocorrelation_match_the_reference():
    from sysbio_modeling_b200.project import ensembles
    g = _golden_ens()
    for case in g['sampling_matrix']:
        H, ref = np.array(case['hessian']), np.array(case['samp_mat'])
        for fn in (O.sampling_matrix, ensembles.sampling_matrix):
            S = fn(H, case['cutoff'], case['temperature'], case['step_scale'])
            assert np.allclose(S, ref, rtol=1e-10, atol=1e-13 * np.abs(ref).max())
    ac = g['autocorrelation']
    assert np.allclose(ensembles.autocorrelation(ac['series']), ac['ac'], rtol=1e-10, atol=1e-12)


def test_scale_factor_entropy_matches_the_reference():
    """linear_scale_factor.py:63-80 (QUADPACK over the whole line in the reference): the restatement's windowed QUADPACK
    call, the product's host rule (composite Gauss-Legendre) and - on a GPU, test below - the torch rule."""
    from sysbio_modeling_b200.project.loss_functions.squared_loss.linear_scale_factor import LinearScaleFactor
    from scipy.integrate import quad
    for case in _golden_ens()['entropy']:
        s, d, sg = np.array(case['sim']), np.array(case['data']), np.array(case['std'])
        sds, sde = np.sum(s * s / sg ** 2), np.sum(s * d / sg ** 2)
        T, mu, sig = case['temperature'], case['log_prior'], case['log_prior_sigma']
        h = LinearScaleFactor(mu, sig)
        ent = h.calc_scale_factor_entropy(sds, sde, T)
        assert abs(ent - case['entropy']) <= 1e-8 * max(1.0, abs(case['entropy'])), (case['temperature'], ent, case['entropy'])
        B = sde / sds
        assert abs(B - case['sf']) <= 1e-12 * abs(B)
        lB = np.log(B)
        w = max(np.sqrt(T / (sds * B * B)), sig)
        pts = sorted([0.0, mu - lB])
        ans, _ = quad(lambda u: np.exp(-sds / (2 * T) * (np.exp(u) * B - B) ** 2 - (u + lB - mu) ** 2 / (2 * sig ** 2)),
                      pts[0] - 12 * w, pts[1] + 12 * w, points=pts, limit=1000, epsabs=0, epsrel=1e-12)
        assert abs(np.log(ans) - case['entropy']) <= 1e-8 * max(1.0, abs(case['entropy']))
