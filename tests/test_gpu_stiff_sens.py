"""The stiff model's Jacobian path: RODAS4 with forward sensitivities (csrc/kernels/sbm_ode_kernels.cuh,
sbm_rodas4_sens) against the reference's way of getting them - LSODA on the sensitivity-augmented system
(ode_model.py:83-126; oracle/sbm_oracle.py:48-64) at a tolerance well below the kernel's - against finite differences of the
kernel's own states, and through the Project (C3: scaled Jacobian and gradient of the Robertson project)."""
import numpy as np
import pytest

import sbm_oracle as O
from fixtures import python_rhs

pytestmark = pytest.mark.gpu

T_ROB = np.concatenate([[0.0], np.logspace(-2, 3, 12)])


def _robertson(rtol, atol):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('robertson')
    om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=rtol, atol=atol)
    sm, m, s = python_rhs('robertson')
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    return om, mod, sm, m, s, nominal


def test_robertson_sensitivities_vs_lsoda_augmented_system():
    om, mod, sm, m, s, nominal = _robertson(1e-8, 1e-12)
    rng = np.random.default_rng(11)
    P = nominal * np.exp(0.4 * rng.standard_normal((16, 3)))
    y0 = np.array(mod.initial_conditions, dtype=float)
    y, S = om.calc_jacobian_batch(P, T_ROB, y0)
    assert (om.last_status == 0).all()
    assert om.last_counters['lu'] > 0                              # the Rosenbrock kernel ran, not Dormand-Prince
    assert S.shape == (16, len(T_ROB), 3, 3)
    worst = 0.0
    for b in range(16):
        z0 = np.concatenate([y0, np.zeros(9)])
        Sr = O.calc_jacobian(s, 3, P[b], T_ROB, z0, rtol=1e-12, atol=1e-14).reshape(len(T_ROB), 3, 3)
        yr = O.simulate(m, 3, P[b], T_ROB, y0, rtol=1e-12, atol=1e-14)
        assert np.all(np.abs(y[b] - yr) <= 5e-6 * np.abs(yr) + 1e-10), (b, np.abs(y[b] - yr).max())
        # sensitivities span 12 decades (dy/dk3 ~ 1e-9, dy/dk1 ~ 1): relative to the size of each parameter's column
        scale = np.abs(Sr).max(axis=(0, 1), keepdims=True)
        err = np.abs(S[b] - Sr) / scale
        worst = max(worst, err.max())
        assert err.max() < 2e-5, (b, err.max())
    print('worst sensitivity error relative to the column scale: %.2e' % worst)
    # single-vector reference API: [t, n_vars * n_free], column i * n_free + j
    S1 = om.calc_jacobian(P[0], T_ROB, np.concatenate([y0, np.zeros(9)]))
    assert np.array_equal(S1, S[0].reshape(len(T_ROB), 9))


def test_robertson_sensitivities_vs_finite_differences_of_the_kernel():
    om, mod, sm, m, s, nominal = _robertson(1e-10, 1e-14)
    y0 = np.array(mod.initial_conditions, dtype=float)
    y, S = om.calc_jacobian_batch(nominal[None, :], T_ROB, y0)
    for j in range(3):
        d = 1e-5 * nominal[j]
        Pp, Pm = nominal.copy(), nominal.copy()
        Pp[j] += d
        Pm[j] -= d
        yp = om.simulate_batch(np.stack([Pp, Pm]), T_ROB, y0)
        fd = (yp[0] - yp[1]) / (2 * d)
        scale = np.abs(fd).max()
        assert np.abs(S[0, :, :, j] - fd).max() <= 2e-5 * scale, (j, np.abs(S[0, :, :, j] - fd).max(), scale)


def test_c3_project_jacobian_and_gradient_on_the_stiff_solver():
    from tools import workloads as wl
    proj, make_oracle = wl.c3_project(rtol=1e-8, atol=1e-12)
    oracle = make_oracle()
    sm, nominal, y0 = wl.nominal_of('robertson')
    thetas = wl.make_thetas(nominal, 64, 3, 0.4)
    out = proj.jacobian_batch(thetas)
    assert (out['status'] == 0).all()
    assert out['counters']['lu'] > 0 if isinstance(out.get('counters'), dict) else True
    assert out['jac'].shape == (64, 24, 3)
    g = np.einsum('brc,br->bc', out['jac'], out['resid'])
    assert np.allclose(out['grad'], g, rtol=1e-10, atol=1e-12 * np.abs(g).max())
    for b in (0, 1, 31, 63):
        J0 = oracle.jacobian(thetas[b])
        assert np.allclose(out['jac'][b], J0, rtol=2e-4, atol=2e-5 * np.abs(J0).max()), (b, np.abs(out['jac'][b] - J0).max())
        g0 = oracle.rss_gradient(thetas[b])
        assert np.allclose(out['grad'][b], g0, rtol=1e-3, atol=1e-4 * np.abs(g0).max()), (b, out['grad'][b], g0)
    # the reference API on one vector
    Jdf = proj.calc_project_jacobian(thetas[0])
    assert np.allclose(Jdf.values, out['jac'][0], rtol=1e-12, atol=0)
    # a batched Levenberg-Marquardt fit of the stiff project converges from perturbed starts
    fit = proj.lm_fit_batch(thetas[:8], max_iter=30)
    assert np.all(fit['ssr'] <= 0.5 * (out['resid'][:8] ** 2).sum(axis=1) + 1e-9)


def test_rodas4_sens_refused_where_unsupported():
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files, _lib
    om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'), solver='rodas4')
    mod = model_files.module('egfr_signalling')
    sm, m, s = python_rhs('egfr_signalling')
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    with pytest.raises(_lib.SbmError):
        om.calc_jacobian_batch(nominal[None, :], np.array([0.0, 1.0]), np.array(mod.initial_conditions, dtype=float))


def test_rodas4_sens_agrees_with_dormand_prince_on_the_mapk_model():
    """Two integrators, one model: on the (non-stiff) Kholodenko cascade the Rosenbrock kernel with sensitivities and the
    CTA-transposed Dormand-Prince kernel agree to the tolerance (8 states, 22 parameters: 176 sensitivity components)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    from tools import workloads as wl
    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    P = np.exp(wl.make_thetas(nominal, 8, 4, 0.25))
    t = np.concatenate([[0.0], wl.C4_T])
    ro = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), solver='rodas4', rtol=1e-8, atol=1e-8)
    dp = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), solver='dp5', rtol=1e-10, atol=1e-10)
    y1, S1 = ro.calc_jacobian_batch(P, t, y0)
    assert (ro.last_status == 0).all() and ro.last_counters['lu'] > 0
    y2, S2 = dp.calc_jacobian_batch(P, t, y0)
    assert (dp.last_status == 0).all()
    assert np.all(np.abs(y1 - y2) <= 1e-6 * (1 + np.abs(y2)))
    scale = np.abs(S2).max(axis=(1, 2), keepdims=True)          # per set and parameter
    assert (np.abs(S1 - S2) / scale).max() < 1e-5, (np.abs(S1 - S2) / scale).max()


def test_solver_auto_picks_by_measurement():
    """``Project(solver='auto')``: the stiff Robertson project ends up on the Rosenbrock kernels (the explicit kernel
    exhausts its probe budget), the non-stiff Michaelis-Menten project on Dormand-Prince; results equal the explicit choice."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    from tools import workloads as wl
    from fixtures import project_mm_two_measures
    import json, os
    ref, _ = wl.c3_project(rtol=1e-8, atol=1e-12)
    om = sb.OdeModel.from_model_file(model_files.path('robertson'), rtol=1e-8, atol=1e-12)
    mapping = {k: ('direct', v) for k, v in wl.C3_OBS.items()}
    auto = sb.Project(om, wl.c3_experiments(ref.workload_data), {}, mapping, solver='auto', rtol=1e-8, atol=1e-12)
    sm, nominal, y0 = wl.nominal_of('robertson')
    thetas = wl.make_thetas(nominal, 32, 3, 0.4)
    a, b = auto.residuals_batch(thetas), ref.residuals_batch(thetas)
    assert auto.auto_solver_report[False]['choice'] == 'rodas4', auto.auto_solver_report
    assert auto.auto_solver_report[False]['dp5']['failed'] > 0            # the explicit kernel ran out of steps
    assert np.array_equal(a['resid'], b['resid'])
    ja, jb = auto.jacobian_batch(thetas), ref.jacobian_batch(thetas)
    assert auto.auto_solver_report[True]['choice'] == 'rodas4'
    assert np.array_equal(ja['jac'], jb['jac'])
    with open(os.path.join(os.path.dirname(__file__), 'golden', 'reference_golden.json')) as fh:
        golden = json.load(fh)
    p_dp, _ = project_mm_two_measures(golden)
    p_auto, _ = project_mm_two_measures(golden, solver='auto')
    th = np.log(0.01) + 0.2 * np.random.default_rng(3).standard_normal((16, p_dp.n_project_params))
    ra, rb = p_auto.jacobian_batch(th), p_dp.jacobian_batch(th)
    assert p_auto.auto_solver_report[True]['choice'] == 'dp5', p_auto.auto_solver_report
    assert np.array_equal(ra['jac'], rb['jac']) and np.array_equal(ra['resid'], rb['resid'])


def test_rodas4_sens_failure_status_and_shared_columns():
    """A step budget that cannot be met: status MAXSTEPS, NaN rows -> inf residuals (squared_loss_function.py:28-32) on that
    set only; and two experiments that share every parameter (their rows accumulate into the same project columns)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    from tools import workloads as wl
    ref, _ = wl.c3_project(rtol=1e-8, atol=1e-12)
    sm, nominal, y0 = wl.nominal_of('robertson')
    thetas = wl.make_thetas(nominal, 8, 3, 0.3)
    ref.max_steps = 5
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        out = ref.jacobian_batch(thetas)
    assert (out['status'] == 1).all()                         # SBM_ERR_MAXSTEPS
    assert np.isinf(out['resid']).all()
    ref.max_steps = 0
    out = ref.jacobian_batch(thetas)
    assert (out['status'] == 0).all() and np.isfinite(out['jac']).all()
    # two identical experiments under different names: same parameters (all Global), twice the rows
    om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=1e-8, atol=1e-12)
    mapping = {k: ('direct', v) for k, v in wl.C3_OBS.items()}
    e1 = wl.c3_experiments(ref.workload_data)[0]
    e2 = wl.c3_experiments(ref.workload_data)[0]
    e2.name = 'Robertson_again'
    both = sb.Project(om, [e1, e2], {}, mapping, solver='rodas4', rtol=1e-8, atol=1e-12)
    o2 = both.jacobian_batch(thetas)
    assert o2['jac'].shape == (8, 48, 3)
    assert np.allclose(o2['jac'][:, :24], out['jac'], rtol=1e-12, atol=0) and np.allclose(o2['jac'][:, 24:], out['jac'], rtol=1e-12, atol=0)
    assert np.allclose(o2['grad'], 2 * out['grad'], rtol=1e-10, atol=1e-12 * np.abs(out['grad']).max())
