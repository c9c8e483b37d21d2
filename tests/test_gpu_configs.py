"""GPU parity on the BASELINE configs beyond the reference's own fixtures (C2 reversible MM with experiment-local
parameters, C3 Robertson through the Rosenbrock kernel, C4 MAPK at reduced batch) and on the loss-function front ends
(reference tests/test_Loss_Functions.py run through the CUDA epilogue)."""
import numpy as np
import pandas as pd
import pytest

import sbm_oracle as O
from fixtures import python_rhs

pytestmark = pytest.mark.gpu


def test_c3_robertson_rodas4_vs_lsoda():
    """Stiff Robertson kinetics: RODAS4 kernel (rtol 1e-6, atol 1e-10) against LSODA at tight tolerance."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('robertson')
    om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=1e-6, atol=1e-10)
    sm, m, s = python_rhs('robertson')
    rng = np.random.default_rng(3)
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    P = nominal * np.exp(0.5 * rng.standard_normal((32, 3)))
    t = np.concatenate([[0.0], np.logspace(-2, 3, 12)])
    y = om.simulate_batch(P, t, mod.initial_conditions)
    assert (om.last_status == 0).all()
    assert om.last_counters['lu'] > 0
    for b in range(32):
        yr = O.simulate(m, 3, P[b], t, np.array(mod.initial_conditions), rtol=1e-11, atol=1e-13)
        assert np.all(np.abs(y[b] - yr) <= 5e-5 * np.abs(yr) + 1e-8), (b, np.abs(y[b] - yr).max())
    # mass conservation y1 + y2 + y3 = 1 is a linear invariant: Rosenbrock methods keep it to round-off
    assert np.abs(y.sum(axis=2) - 1.0).max() < 1e-12
    # tighter tolerance -> closer to the reference solution
    om2 = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=1e-9, atol=1e-13)
    y2 = om2.simulate_batch(P[:4], t, mod.initial_conditions)
    yr = np.array([O.simulate(m, 3, P[b], t, np.array(mod.initial_conditions), rtol=1e-12, atol=1e-14) for b in range(4)])
    assert np.all(np.abs(y2 - yr) <= 2e-7 * np.abs(yr) + 1e-11)


def test_c2_reversible_mm_local_parameters_project():
    """C2 shape at reduced size: experiment-local parameters, non-zero initial conditions, one observable."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('reversible_mm')
    sm, m, s = python_rhs('reversible_mm')
    E = 24
    rng = np.random.default_rng(2)
    t = np.linspace(0.25, 4, 16)
    nominal = np.array([mod.nominal[k] for k in sm.param_order])

    def exps():
        out = []
        r = np.random.default_rng(22)
        for k in range(E):
            data = 0.5 + 0.4 * r.random(16)
            out.append(sb.Experiment('exp%03d' % k, sb.TimecourseMeasurement('P', data, t, 0.05 * data + 1e-3),
                                     initial_conditions=mod.initial_conditions))
        return out
    settings = {'Local': list(sm.param_order)}
    om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
    proj = sb.Project(om, exps(), dict(settings), {'P': ('direct', 3)})
    oracle = O.OracleProject(m, s, 4, sm.param_order, exps(), dict(settings), {'P': ('direct', 3)})
    assert proj.n_project_params == 4 * E == oracle.n_project_params
    theta = np.tile(np.log(nominal), E) + 0.3 * rng.standard_normal(4 * E)
    r = np.asarray(proj.residuals(theta))
    r0 = oracle.residuals(theta)
    assert np.allclose(r, r0, rtol=1e-6, atol=1e-6)
    J = proj.calc_project_jacobian(theta).values
    J0 = oracle.jacobian(theta)
    assert J.shape == (16 * E, 4 * E)
    assert np.allclose(J, J0, rtol=1e-6, atol=1e-7)
    # block structure: an experiment's rows only touch its own four columns
    assert np.count_nonzero(J[:16, 4:]) == 0
    g = np.asarray(proj.calc_rss_gradient(theta))
    assert np.allclose(g, oracle.rss_gradient(theta), rtol=1e-6, atol=1e-5)
    ssr = proj.calc_sum_square_residuals(theta)
    assert abs(ssr - oracle.sum_square_residuals(theta)) <= 1e-6 * ssr


def test_c4_mapk_sensitivities_vs_lsoda_and_fd():
    """MAPK cascade (184 augmented states): warp-group kernel vs the reference path at the reference's tolerance."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('mapk_kholodenko')
    sm, m, s = python_rhs('mapk_kholodenko')
    om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'))
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    rng = np.random.default_rng(4)
    P = nominal * np.exp(0.25 * rng.standard_normal((6, 22)))
    # the reference integrates on its 1000-node grid and reads the nodes at / after the measurement times
    # (base_project.py:418-420, utils.py:18-21); LSODA also needs the fine grid to stay inside its
    # per-interval step budget on this oscillator
    grid = np.linspace(0, 3000, 1000)
    idx = np.concatenate([[0], np.searchsorted(grid, np.linspace(150, 3000, 20))])
    t = grid[idx]
    y0 = np.array(mod.initial_conditions)
    y, S = om.calc_jacobian_batch(P, t, y0)
    assert (om.last_status == 0).all()
    for b in range(6):
        yr, Sr = O.calc_jacobian(s, 8, P[b], grid, np.concatenate([y0, np.zeros(176)]), rtol=1e-10, atol=1e-10,
                                 return_states=True)
        yr, Sr = yr[idx], Sr[idx]
        # LSODA's own global error at 1e-10 is ~1e-7 relative on this oscillating system (40-500 x tol on the worst
        # component); the comparison with a converged reference (DOP853 at 1e-13) within 10 x tol for the states and in
        # the weighted RMS norm for the sensitivities is tests/test_gpu_bench_parity.py::test_c4_states_and_sensitivities_within_10x_tolerance
        assert np.all(np.abs(y[b] - yr) <= 1e-6 * (1 + np.abs(yr)))
        assert np.all(np.abs(S[b].reshape(21, 176) - Sr) <= 1e-6 * (1 + np.abs(Sr)))
    # conservation: each of the three kinase pools is conserved, so the sensitivities of a pool sum to zero
    assert np.abs(y[:, :, 0] + y[:, :, 1] - 100.0).max() < 1e-9
    assert np.abs(S[:, :, 0, :] + S[:, :, 1, :]).max() < 1e-7
    # central finite differences of the plain solve w.r.t. two parameters
    for j in (0, 14):
        h = 1e-5 * P[0, j]
        Pp, Pm = P[:1].copy(), P[:1].copy()
        Pp[0, j] += h
        Pm[0, j] -= h
        fd = (om.simulate_batch(Pp, t, y0)[0] - om.simulate_batch(Pm, t, y0)[0]) / (2 * h)
        assert np.allclose(S[0, :, :, j], fd, rtol=1e-4, atol=1e-4 * np.abs(fd).max())


def test_tolerance_controls_error_dp5():
    """Error of the DP5 kernel scales with the requested tolerance (closed form of the 1-state model)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    t = np.linspace(0, 100, 11)
    kd, ks = 0.03, 0.5
    exact = ks / kd * (1 - np.exp(-kd * t))
    errs = []
    for tol in (1e-4, 1e-7, 1e-10, 1e-12):
        om = sb.OdeModel.from_model_file(model_files.path('simple_synth_deg'), rtol=tol, atol=tol)
        y = om.simulate(np.array([kd, ks]), t)[:, 0]
        errs.append(np.abs(y - exact).max())
        assert errs[-1] <= 10 * tol * (1 + exact.max())
    assert errs[0] > errs[1] > errs[2]


# ---- loss-function front ends (reference tests/test_Loss_Functions.py) ------------------------------
def _lin_square_frames():
    t = np.linspace(0, 100, 101)
    sim = pd.DataFrame({'mean': np.hstack((2 * t, t ** 2)), 'timecourse': np.hstack((t, t))})
    sim.index = pd.MultiIndex.from_tuples([('Experiment 1', 'Lin')] * 101 + [('Experiment 2', 'Square')] * 101)
    return sim


def test_loss_no_scale_factors_and_scale_factors():
    from sysbio_modeling_b200 import SquareLossFunction, NormalizedSquareLossFunction
    sim = _lin_square_frames()
    rng = np.random.default_rng(0)
    measures = sim.copy()
    measures.insert(1, 'std', np.ones(len(sim)))
    noise = rng.standard_normal(len(sim))
    measures['mean'] -= noise
    lf = SquareLossFunction()
    res = lf.residuals(sim, measures)
    assert np.allclose(res, noise)
    assert abs(lf.evaluate(sim, measures) - 0.5 * np.sum(noise ** 2)) < 1e-9
    measures['std'] = np.abs(rng.standard_normal(len(sim))) + 0.1
    assert np.allclose(lf.residuals(sim, measures), (sim['mean'] - measures['mean']) / measures['std'])
    # normalized loss = square loss / mean (test_normalized_loss_function)
    m2 = sim.copy()
    m2.insert(1, 'std', np.ones(len(sim)))
    m2['mean'] += 5.0
    r_sq = np.asarray(SquareLossFunction().residuals(sim, m2))
    r_n = np.asarray(NormalizedSquareLossFunction().residuals(sim, m2))
    assert np.allclose(r_n * m2['mean'].values, r_sq)
    # exact scale factors 2.0 / 3.6 (test_scale_factors_loss)
    lf = SquareLossFunction(sf_groups=['Lin', 'Square'])
    m3 = sim.copy()
    m3.insert(1, 'std', np.ones(len(sim)))
    m3.loc[(slice(None), 'Lin'), 'mean'] *= 2.0
    m3.loc[(slice(None), 'Square'), 'mean'] *= 3.6
    lf.update_scale_factors(sim, m3)
    assert abs(lf.scale_factors['Lin'].sf - 2.0) < 1e-14 and abs(lf.scale_factors['Square'].sf - 3.6) < 1e-14
    scaled = lf.scale_sim_values(sim)
    assert np.allclose(scaled['mean'], m3['mean'])
    assert np.allclose(lf.residuals(sim, m3), 0, atol=1e-9)
    lf2 = SquareLossFunction(frozenset(['Variable_1', 'Variable_2']))
    assert 'Variable_1' in lf2.scale_factors and 'Variable_2' in lf2.scale_factors


def _model3(p, t):
    a, b, c = p
    return a * np.sin(b * t) - c * t ** 2.0


def _frames3(vals, t):
    sim = pd.DataFrame({'mean': vals, 'timecourse': t})
    sim.index = pd.MultiIndex.from_tuples([('Experiment_1', 'Val')] * len(t))
    return sim


def test_square_loss_sf_gradient_and_jacobian_vs_fd():
    from sysbio_modeling_b200 import SquareLossFunction
    t = np.linspace(0, 10, 11)
    p1 = np.array([0.3, 0.5, 1.3])
    sim = _frames3(_model3(p1, t), t)
    jac = pd.DataFrame(np.stack([np.sin(p1[1] * t), p1[0] * t * np.cos(p1[1] * t), -t ** 2], axis=1),
                       columns=['a', 'b', 'c'], index=sim.index)
    measures = sim.copy()
    measures.insert(1, 'std', np.ones(11))
    assert np.array_equal(SquareLossFunction().jacobian(sim, measures, jac).values, jac.values)
    lf = SquareLossFunction(sf_groups=['Val'])
    measures['mean'] = measures['mean'] * 5 - np.random.default_rng(1).standard_normal(11)

    def sf_of(p):
        lf.update_scale_factors(_frames3(_model3(p, t), t), measures)
        return np.array([lf.scale_factors['Val'].sf])
    lf.update_scale_factors_gradient(sim, measures, jac)
    assert np.allclose(lf.scale_factors['Val'].gradient, O.central_difference(sf_of, p1)[0], rtol=1e-5)
    Jfd = O.central_difference(lambda p: np.asarray(lf.residuals(_frames3(_model3(p, t), t), measures)), p1)
    assert np.allclose(lf.jacobian(sim, measures, jac).values, Jfd, rtol=1e-5, atol=1e-7)


def test_log_loss_and_sf_priors():
    from sysbio_modeling_b200 import SquareLossFunction, LogSquareLossFunction
    sim = _lin_square_frames()
    sim = sim[sim['mean'] != 0]
    measures = sim.copy()
    measures.insert(1, 'std', np.ones(len(sim)))
    measures['mean'] /= 2.3
    res = LogSquareLossFunction().residuals(sim, measures)
    assert np.allclose(res, np.log(2.3))
    lf = LogSquareLossFunction(sf_groups=['Lin', 'Square'])
    assert np.allclose(lf.residuals(sim, measures), 0, atol=1e-12)
    assert abs(lf.scale_factors['Lin'].sf - 1 / 2.3) < 1e-12
    # log-loss Jacobian vs finite differences (test_log_squared_loss_jacobian)
    t = np.linspace(0, 11, 11)
    f = lambda p: p[0] + t * p[2] * p[1] ** 2
    p1 = np.array([0.3, 0.5, 1.3])
    sim3 = _frames3(f(p1), t)
    jac = pd.DataFrame(np.stack([np.ones(11), 2 * t * p1[2] * p1[1], t * p1[1] ** 2], axis=1), columns=list('abc'),
                       index=sim3.index)
    m3 = sim3.copy()
    m3.insert(1, 'std', np.ones(11))
    m3['mean'] = m3['mean'] * 5 + np.abs(np.random.default_rng(2).standard_normal(11))
    llf = LogSquareLossFunction(sf_groups=['Val'])
    Jfd = O.central_difference(lambda p: np.asarray(llf.residuals(_frames3(f(p), t), m3)), p1)
    assert np.allclose(llf.jacobian(sim3, m3, jac).values, Jfd, rtol=1e-5, atol=1e-8)
    # scale-factor priors (test_scale_factor_priors): residual of the prior row = (log 5 - 1)/2
    lf = SquareLossFunction(sf_groups=['Lin', 'Square'])
    m = sim.copy()
    m.insert(1, 'std', np.ones(len(sim)))
    m['mean'] *= 5
    lf.set_scale_factor_priors('Lin', 1.0, 2.0)
    assert lf.scale_factors['Lin'].log_prior == 1.0 and lf.scale_factors['Lin'].log_prior_sigma == 2.0
    with pytest.raises(KeyError):
        lf.residuals(sim, m)
    prior = pd.DataFrame({'mean': [1.0], 'timecourse': [np.nan]},
                         index=pd.MultiIndex.from_tuples([('~~SF_Prior', '~Lin')]))
    sim_p = pd.concat([sim, prior], axis=0)
    m_p = sim_p.copy()
    m_p.insert(1, 'std', np.ones(len(sim_p)))
    m_p['mean'] *= 5
    m_p.loc[('~~SF_Prior', '~Lin'), :] = np.array([1.0, 2.0, np.nan])
    res = np.asarray(lf.residuals(sim_p, m_p))
    assert abs(res[-1] - (np.log(5) - 1) / 2.0) < 1e-12


def test_device_reciprocal_accuracy(tmp_path):
    """sbm_rcp (MUFU seed + Newton steps, no slow-path branch) stays within 1 ulp of IEEE division."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'check_device_math')
    subprocess.run(['nvcc', '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-o', exe,
                    os.path.join(root, 'tests', 'cuda', 'check_device_math.cu')], check=True)
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    print(out)
    vals = dict(zip(out.split()[0::2], [float(v) for v in out.split()[1::2]]))
    assert vals['sbm_rcp_ulp'] <= 1.0
    assert vals['seed_relerr'] < 1e-5
    # branch-free log / exp of the generated model code: within 2 ulp of libm (itself <= 1 ulp) on positive normal
    # arguments; never a wrong finite number outside (NaN for log, Inf / 0 for exp)
    assert vals['sbm_log_ulp'] <= 2.0 and vals['sbm_log_near1_ulp'] <= 2.0 and vals['sbm_exp_ulp'] <= 2.0
    assert vals['pow_relerr'] < 5e-15
    for k in ('log_zero', 'log_neg', 'log_inf', 'log_subnormal', 'exp_nan'):
        assert np.isnan(vals[k]), k
    assert vals['log_one'] == 0.0 and vals['exp_zero'] == 1.0
    assert vals['exp_800'] == np.inf and vals['exp_1e12'] == np.inf
    assert vals['exp_m800'] == 0.0 and vals['exp_m1e12'] == 0.0
    assert abs(vals['exp_m740'] / np.exp(-740.0) - 1) < 1e-3         # subnormal result


def test_cta_transposed_layout_matches_warp_group_layout():
    """The two sensitivity kernels (lane = column warp groups, and lane = trajectory / warp = column group with a
    state warp) integrate the same augmented system: states and sensitivities agree to the solver tolerance, on the
    MAPK cascade (8 warps per CTA) and on the 2-state Michaelis-Menten fixture (single-warp CTAs)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    for name, nb, tgrid in (('mapk_kholodenko', 70, np.concatenate([[0.0], np.linspace(150, 3000, 20)])),
                            ('michaelis_menten', 37, np.linspace(0, 30, 13))):
        mod = model_files.module(name)
        out = {}
        for layout in (1, 2):
            om = sb.OdeModel.from_model_file(model_files.path(name), sens_layout=layout)
            assert om.compiled.info()['sens_layout'] == layout
            nom = getattr(mod, 'nominal', {'vmax': 1.0, 'km': 0.5, 'k_synt_s': 0.3, 'k_deg_s': 0.01, 'k_deg_p': 0.05})
            nominal = np.array([nom[k] for k in om.param_order])
            P = nominal * np.exp(0.25 * np.random.default_rng(11).standard_normal((nb, len(nominal))))
            y0 = np.array(getattr(mod, 'initial_conditions', np.zeros(om.n_vars)), dtype=float)
            out[layout] = om.calc_jacobian_batch(P, tgrid, y0)
            assert (om.last_status == 0).all()
            assert om.last_counters['trajectories'] == nb
        (y1, S1), (y2, S2) = out[1], out[2]
        assert np.all(np.abs(y1 - y2) <= 1e-8 * (1 + np.abs(y1)))
        assert np.all(np.abs(S1 - S2) <= 1e-7 * (1 + np.abs(S1)))
    # other shapes of the same kernel on the reversible Michaelis-Menten model: 4 column warps x 1 column and
    # 2 x 2 (each plus the model warp, S1 / S4 in tensor memory) against one thread per trajectory (registers)
    mod = model_files.module('reversible_mm')
    tgrid = np.concatenate([[0.0], np.linspace(0.25, 4, 16)])
    res = {}
    for cols in (0, 1, 2):
        om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'), sens_layout=2, cols_per_lane=cols)
        info = om.compiled.info()
        assert info['sens_layout'] == 2 and info['ct_warps'] == {0: 1, 1: 5, 2: 3}[cols]
        nominal = np.array([mod.nominal[k] for k in om.param_order])
        P = nominal * np.exp(0.3 * np.random.default_rng(14).standard_normal((100, 4)))
        res[cols] = om.calc_jacobian_batch(P, tgrid, np.array(mod.initial_conditions))
        assert (om.last_status == 0).all()
    for cols in (1, 2):
        assert np.all(np.abs(res[cols][0] - res[0][0]) <= 1e-9 * (1 + np.abs(res[0][0])))
        assert np.all(np.abs(res[cols][1] - res[0][1]) <= 1e-8 * (1 + np.abs(res[0][1])))


def test_c5_egfr_scale_ssr_ensemble():
    """C5 shape at reduced size: 25-species EGFR-scale network, many parameter sets ("walkers"), residuals / SSR only
    (no sensitivities) - plain DP5 kernel vs the reference path (LSODA), and the batched SSR vs the oracle project."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('egfr_signalling')
    sm, m, s = python_rhs('egfr_signalling')
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    y0 = np.array(mod.initial_conditions)
    rng = np.random.default_rng(5)
    om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'))
    t = np.linspace(0, 600, 13)
    P = nominal * np.exp(0.2 * rng.standard_normal((8, len(nominal))))
    y = om.simulate_batch(P, t, y0)
    assert (om.last_status == 0).all()
    for b in range(8):
        yr = O.simulate(m, 25, P[b], t, y0, rtol=1e-11, atol=1e-11)
        assert np.all(np.abs(y[b] - yr) <= 1e-7 * (1 + np.abs(yr))), (b, np.abs(y[b] - yr).max())
    # conserved pools: Ras (15, 16), Raf (17, 18), MEK (19..21), ERK (22..24)
    assert np.abs(y[:, :, 15] + y[:, :, 16] - 120.0).max() < 1e-8
    assert np.abs(y[:, :, 22] + y[:, :, 23] + y[:, :, 24] - 200.0).max() < 1e-8
    # ensemble SSR: data generated at the nominal parameters, 512 walkers around them
    tm = np.linspace(50, 600, 12)
    ynom = O.simulate(m, 25, nominal, np.concatenate([[0.0], tm]), y0, rtol=1e-11, atol=1e-11)[1:]

    def exps():
        return [sb.Experiment('EGF_stimulation',
                              [sb.TimecourseMeasurement('ERKPP', ynom[:, 24], tm, 0.05 * ynom[:, 24] + 0.1),
                               sb.TimecourseMeasurement('RasT', ynom[:, 16], tm, 0.05 * ynom[:, 16] + 0.1)],
                              initial_conditions=mod.initial_conditions)]
    mapping = {'ERKPP': ('direct', 24), 'RasT': ('direct', 16)}
    proj = sb.Project(om, exps(), {}, mapping)
    oracle = O.OracleProject(m, s, 25, sm.param_order, exps(), {}, mapping)
    theta = np.log(nominal)[None, :] + 0.1 * rng.standard_normal((512, len(nominal)))
    out = proj.residuals_batch(theta)
    assert (out['status'] == 0).all()
    assert np.allclose(out['ssr'], 0.5 * (out['resid'] ** 2).sum(axis=1), rtol=1e-12)
    for b in (0, 17, 511):
        ssr0 = oracle.sum_square_residuals(theta[b])
        assert abs(out['ssr'][b] - ssr0) <= 1e-6 * ssr0 + 1e-9, (b, out['ssr'][b], ssr0)


def test_cta_transposed_kernel_failure_status_and_queue_refill():
    """CTA-transposed sensitivity kernel: (1) a step budget that is too small gives status max_steps and NaN rows for
    every trajectory; (2) more trajectories than one wave of persistent CTAs (lanes refill from the global queue):
    every trajectory is integrated exactly once and matches a stand-alone solve."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('mapk_kholodenko')
    y0 = np.array(mod.initial_conditions)
    t = np.concatenate([[0.0], np.linspace(150, 3000, 20)])
    om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), max_steps=5)
    assert om.compiled.info()['sens_layout'] == 2
    nominal = np.array([mod.nominal[k] for k in om.param_order])
    P = nominal * np.exp(0.25 * np.random.default_rng(12).standard_normal((45, 22)))
    y, S = om.calc_jacobian_batch(P, t, y0)
    assert (om.last_status == 1).all()                      # SBM_ERR_MAXSTEPS
    assert np.isnan(y[:, 1:]).all() and np.isnan(S[:, 1:]).all()
    # queue refill: 148 SMs x 32 lanes = 4,736 trajectories in flight; take more, at a loose tolerance
    om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), rtol=1e-6, atol=1e-6)
    nb = 4736 + 1000 + 37
    P = nominal * np.exp(0.25 * np.random.default_rng(13).standard_normal((nb, 22)))
    y, S = om.calc_jacobian_batch(P, t, y0)
    assert (om.last_status == 0).all()
    assert om.last_counters['trajectories'] == nb
    assert np.isfinite(y).all() and np.isfinite(S).all()
    for b in (0, 31, 4735, 4736, nb - 1):
        y1, S1 = om.calc_jacobian_batch(P[b:b + 1], t, y0)
        assert np.array_equal(y1[0], y[b]) and np.array_equal(S1[0], S[b])      # same arithmetic, any lane


def test_c5_egfr_sensitivities_vs_lsoda_and_fd():
    """EGFR-scale network WITH sensitivities (25 states x 52 parameters = 1,325 augmented states): beyond the
    CTA-transposed layout (N <= 24), so the warp-group kernel integrates it - against the reference path (LSODA at 1e-10
    on the numba-compiled generated sensitivity RHS) and central finite differences of the plain solve.  This is the
    Jacobian behind the sampler's Hessian J^T J (reference Ensembles.py:226-258)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('egfr_signalling')
    sm, m, s = python_rhs('egfr_signalling')
    try:
        import numba
        s = numba.jit("void(f8[:], f8, f8[:], f8[:])", nopython=True)(s)
    except Exception:
        pass
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    y0 = np.array(mod.initial_conditions)
    om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'))
    assert om.compiled.info()['sens_layout'] == 1
    rng = np.random.default_rng(55)
    P = nominal * np.exp(0.2 * rng.standard_normal((3, len(nominal))))
    t = np.linspace(0, 600, 7)
    y, S = om.calc_jacobian_batch(P, t, y0)
    assert (om.last_status == 0).all() and S.shape == (3, 7, 25, 52)
    NF = 52
    for b in range(2):
        yr, Sr = O.calc_jacobian(s, 25, P[b], t, np.concatenate([y0, np.zeros(25 * NF)]), rtol=1e-10, atol=1e-10,
                                 return_states=True)
        assert np.all(np.abs(y[b] - yr) <= 1e-6 * (1 + np.abs(yr)))
        assert np.all(np.abs(S[b].reshape(7, 25 * NF) - Sr) <= 2e-6 * (1 + np.abs(Sr))), np.abs(S[b].reshape(7, -1) - Sr).max()
    for j in (0, 17, 51):
        h = 1e-5 * P[0, j]
        Pp, Pm = P[:1].copy(), P[:1].copy()
        Pp[0, j] += h
        Pm[0, j] -= h
        fd = (om.simulate_batch(Pp, t, y0)[0] - om.simulate_batch(Pm, t, y0)[0]) / (2 * h)
        assert np.allclose(S[0, :, :, j], fd, rtol=2e-4, atol=2e-4 * max(np.abs(fd).max(), 1e-12))


def test_push_form_kernel_matches_flat_stage_loop():
    """States-only Dormand-Prince: the push-form kernel (default) against the flat stage loop (SBM_PLAIN_PUSH=0) on the
    EGFR-scale model - same tableau, same controller: same step counts, results equal to rounding."""
    import os
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    from tools import workloads as wl
    sm, nominal, y0 = wl.nominal_of('egfr_signalling')
    P = np.exp(wl.make_thetas(nominal, 64, 5, 0.1))
    t = np.concatenate([[0.0], wl.C5_T])
    om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'), rtol=1e-9, atol=1e-9)
    om.dense_output = 2
    old = os.environ.get('SBM_PLAIN_PUSH')
    try:
        os.environ['SBM_PLAIN_PUSH'] = '1'
        ya = om.simulate_batch(P, t, y0)
        ca = dict(om.last_counters)
        assert (om.last_status == 0).all()
        os.environ['SBM_PLAIN_PUSH'] = '0'
        yb = om.simulate_batch(P, t, y0)
        cb = dict(om.last_counters)
    finally:
        if old is None:
            os.environ.pop('SBM_PLAIN_PUSH', None)
        else:
            os.environ['SBM_PLAIN_PUSH'] = old
    assert abs(ca['steps'] - cb['steps']) <= 2 and abs(ca['rejects'] - cb['rejects']) <= 2, (ca, cb)
    assert np.all(np.abs(ya - yb) <= 1e-9 * (1 + np.abs(yb)))
