"""Generate golden vectors by running the REAL reference (FedericoV/SysBio_Modeling) in this container.

/root/reference is Python 2 and cannot be imported as is.  This script makes a throw-away copy under
/tmp, applies the mechanical py2->py3 / pandas-3 patches listed below (no numerical change), imports
it and records its outputs on the reference's own test fixtures.  The GPU box has no /root/reference,
so the numbers are committed as ``tests/golden/reference_golden.json``.

Run:  python tests/golden/make_golden.py        (only works where /root/reference exists)

Patches:
  1. implicit relative imports -> explicit        (model/ode_model.py:3, project/base_project.py:9,
                                                   loss_functions/squared_loss/__init__.py:4-6)
  2. print statements -> print()                  (base_project.py:195,1058-1062)
  3. dict.keys() list concatenation               (base_project.py:218-222)
  4. DataFrame.sortlevel -> sort_index            (base_project.py:340)
  5. ``.values[...] = x`` on frames (read-only under pandas copy-on-write) -> ``.iloc``/``.loc``
  6. ``.ix`` -> ``.loc``
  7. drop(level=0) of absent labels -> errors='ignore'
  8. ``is 'all'`` -> ``== 'all'``;  ``zip(*x)[0]`` -> ``list(zip(*x))[0]`` (base_project.py:1079)
"""
import json
import os
import re
import shutil
import sys

import numpy as np

REF = '/root/reference'
TMP = '/tmp/sbm_ref_golden'
PKG = os.path.join(TMP, 'SysBio_Modeling')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_golden.json')


def sub(path, pairs, regex=False):
    p = os.path.join(PKG, path)
    s = open(p).read()
    for a, b in pairs:
        if regex:
            s2 = re.sub(a, b, s)
        else:
            s2 = s.replace(a, b)
        if s2 == s:
            raise RuntimeError('patch did not apply in %s: %r' % (path, a))
        s = s2
    open(p, 'w').write(s)


def make_patched_copy():
    if os.path.exists(TMP):
        shutil.rmtree(TMP)
    shutil.copytree(REF, PKG)
    sub('model/ode_model.py', [('from abstract_model import', 'from .abstract_model import')])
    sub('project/base_project.py', [
        ('from loss_functions.squared_loss import', 'from .loss_functions.squared_loss import'),
        ('print "The following parameters are global because no settings were specified: %s " % par_str',
         'print("The following parameters are global because no settings were specified: %s " % par_str)'),
        ('exp_fixed_pars = experiment.fixed_parameters.keys()',
         'exp_fixed_pars = list(experiment.fixed_parameters.keys())'),
        ('measurements_df.sortlevel(inplace=True)', 'measurements_df = measurements_df.sort_index()'),
        ('self._simulations_df.values[:] = 0\n\n        jac_cols', 'self._simulations_df.iloc[:, :] = 0.0\n\n        jac_cols'),
        ('self._simulations_df.values[res_idx:res_idx+extra_residuals, :] = temp',
         'self._simulations_df.iloc[res_idx:res_idx+extra_residuals, :] = temp'),
        ('self._simulations_df.values[residual_idx, 0] = log_p_value',
         'self._simulations_df.iloc[residual_idx, 0] = log_p_value'),
        ('self._model_jacobian_df.values[res_idx:(res_idx + extra_residual), p_project_idx] += measure_jac[:,',
         'self._model_jacobian_df.iloc[res_idx:(res_idx + extra_residual), p_project_idx] += measure_jac[:,'),
        ('self._model_jacobian_df.values[residual_idx, p_idx] = 1',
         'self._model_jacobian_df.iloc[residual_idx, p_idx] = 1'),
        ('self._simulations_df.values[:] = 0\n        self._model_jacobian_df.values[:] = 0',
         'self._simulations_df.iloc[:, :] = 0.0\n        self._model_jacobian_df.iloc[:, :] = 0.0'),
        ("if exp_subset is 'all':", "if exp_subset == 'all':"),
        ('return zip(*project_params)[0]', 'return list(zip(*project_params))[0]'),
        ("if settings is 'all':", "if settings == 'all':"),
        ('out.drop(["~Prior", "~~SF_Prior"], level=0, axis=0, inplace=True)',
         'out.drop(["~Prior", "~~SF_Prior"], level=0, axis=0, inplace=True, errors="ignore")'),
    ])
    # remaining py2 print statements in base_project (plot/debug helpers, not on the hot path)
    p = os.path.join(PKG, 'project/base_project.py')
    s = open(p).read()
    s = re.sub(r'^(\s*)print (?!\()(.+)$', r'\1print(\2)', s, flags=re.M)
    open(p, 'w').write(s)
    sub('project/loss_functions/squared_loss/__init__.py', [
        ('from squared_loss_function import', 'from .squared_loss_function import'),
        ('from log_squared_loss_function import', 'from .log_squared_loss_function import'),
        ('from normalized_squared_loss_function import', 'from .normalized_squared_loss_function import'),
    ])
    sub('project/loss_functions/abstract_loss_function.py', [
        ('simulations.loc[("~~SF_Prior", "~%s" % measure_name)].values[:, 0] = np.log(sf.sf)',
         'simulations.loc[("~~SF_Prior", "~%s" % measure_name), "mean"] = np.log(sf.sf)'),
        ('simulations_jacobian.ix[', 'simulations_jacobian.loc['),
    ])
    sub('project/loss_functions/squared_loss/squared_loss_function.py', [
        ('scaled_jacobian.values[:, :] = np.inf', 'scaled_jacobian.iloc[:, :] = np.inf'),
        ('scaled_jacobian.ix[', 'scaled_jacobian.loc['),
    ])
    sub('project/loss_functions/squared_loss/linear_scale_factor.py', [
        ('import scipy\n', 'import scipy\nimport scipy.integrate\n'),
    ])
    # tests package __init__ imports test modules that need nose; blank it
    open(os.path.join(PKG, 'tests', '__init__.py'), 'w').write('')
    open(os.path.join(PKG, 'tests', 'test_utils', '__init__.py'), 'w').write('')
    # the py2-only modules we never import must at least not be imported by package __init__
    open(os.path.join(PKG, '__init__.py'), 'w').write('')
    open(os.path.join(PKG, 'symbolic', '__init__.py'), 'w').write('')
    sys.path.insert(0, TMP)


def fixtures_project_1state():
    """tests/test_Project.py:29-76 (TestProject.setUpClass) data, verbatim numbers."""
    t1 = np.array([0., 11.11111111, 22.22222222, 33.33333333, 44.44444444, 55.55555556, 66.66666667,
                   77.77777778, 88.88888889, 100.])
    low = np.array([0., 0.11049608, 0.21977125, 0.32783897, 0.43471258, 0.54995561,
                    0.65437492, 0.75764044, 0.85976492, 0.9516258]) * 3.75
    t2 = np.array([5.05050505, 9.09090909, 12.12121212, 16.16161616, 19.19191919, 23.23232323,
                   26.26262626, 30.3030303, 33.33333333, 37.37373737, 40.4040404, 44.44444444,
                   47.47474747, 50.50505051, 54.54545455, 57.57575758, 61.61616162, 64.64646465,
                   68.68686869, 71.71717172, 75.75757576, 78.78787879, 82.82828283, 85.85858586,
                   89.8989899, 92.92929293])
    high = np.array([0.04925086, 0.08689927, 0.11415396, 0.14923229, 0.17462643, 0.20731014,
                     0.23097075, 0.26142329, 0.28346868, 0.31184237, 0.33238284, 0.3588196,
                     0.37795787, 0.39652489, 0.42042171, 0.43772125, 0.45998675, 0.47610533,
                     0.49685087, 0.51186911, 0.53119845, 0.54519147, 0.56320128, 0.57623905,
                     0.59301942, 0.60516718]) * 3.75
    return t1, low, t2, high


def main():
    make_patched_copy()
    from SysBio_Modeling.model.ode_model import OdeModel
    from SysBio_Modeling.project import Project
    from SysBio_Modeling.experiment import Experiment
    from SysBio_Modeling.measurement import TimecourseMeasurement
    from SysBio_Modeling.tests.test_utils.jittable_model import model as m1
    from SysBio_Modeling.tests.test_utils.sens_jittable_model import sens_model as s1
    from SysBio_Modeling.tests.test_utils.jittable_mm_model import model as m2, ordered_params as op2
    from SysBio_Modeling.tests.test_utils.sens_jittable_mm_model import sens_model as s2
    from SysBio_Modeling.tests.test_utils.simple_model_settings import settings as settings1
    from scipy.integrate import odeint
    import scipy
    import pandas

    G = {'versions': {'scipy': scipy.__version__, 'numpy': np.__version__, 'pandas': pandas.__version__}}

    # ---- OdeModel level (tests/test_OdeModel.py:20-52) -------------------------------------
    om = OdeModel(m1, s1, 1, ['k_deg', 'k_synt'], use_jit=False)
    p = np.array([0.001, 0.01])
    t = np.linspace(0, 100, 10)
    G['odemodel_1state'] = {
        'p': p.tolist(), 't': t.tolist(),
        'simulate': om.simulate(p, t).tolist(),
        'calc_jacobian': om.calc_jacobian(p, t, np.zeros(3)).tolist(),
        'test_OdeModel_desired': [0, 0.11049612, 0.21977129, 0.32783901, 0.43471262, 0.54040533,
                                  0.64493016, 0.74830004, 0.85052773, 0.95162583],
    }
    om2 = OdeModel(m2, s2, 2, op2, use_jit=False)
    p2 = np.array([1e-3, 0.001, 0.01, 0.01, 0.001])
    t2 = np.linspace(0, 100, 20)
    G['odemodel_mm'] = {
        'p': p2.tolist(), 't': t2.tolist(), 'param_order': list(op2),
        'simulate': om2.simulate(p2, t2).tolist(),
        'calc_jacobian': om2.calc_jacobian(p2, t2, np.zeros(12)).tolist(),
    }

    # ---- Project level, 1-state fixture with shared sf (tests/test_Project.py:29-176) ------
    t1, low, t2_, high = fixtures_project_1state()
    e_low = Experiment('Low_Deg_Exp', TimecourseMeasurement('Variable_1', low, t1),
                       experiment_settings={'Deg_Rate': 'Low'})
    e_high = Experiment('High_Deg_Exp', TimecourseMeasurement('Variable_1', high, t2_),
                        experiment_settings={'Deg_Rate': 'High'})
    import copy
    proj = Project(om, [e_low, e_high], copy.deepcopy(settings1), {'Variable_1': ('direct', 0)},
                   sf_groups=[frozenset(['Variable_1'])])
    idx = proj.project_param_idx
    cols = {'k_synt': idx['k_synt']['Global'], 'High': idx['Group_1'][('High',)], 'Low': idx['Group_1'][('Low',)]}

    def theta_of(ks, kh, kl):
        v = np.zeros(3)
        v[cols['k_synt']], v[cols['High']], v[cols['Low']] = ks, kh, kl
        return np.log(v)

    out = {'columns': cols, 'n_project_residuals': proj.n_project_residuals, 'points': []}
    for (ks, kh, kl) in [(0.01, 0.01, 0.001), (0.05, 0.02, 0.003)]:
        th = theta_of(ks, kh, kl)
        r = np.asarray(proj.residuals(th))
        sims = proj.get_simulations()
        sf = proj.scale_factors['Variable_1'].sf
        J = proj.calc_project_jacobian(th)
        Jm = proj.get_model_jacobian_df()
        g = np.asarray(proj.calc_rss_gradient(th))
        ssr = proj.calc_sum_square_residuals(th)
        out['points'].append({
            'k': [ks, kh, kl], 'theta': th.tolist(), 'residuals': r.tolist(),
            'index': [list(x) for x in sims.index.tolist()],
            'sims': sims.values[:, 0].tolist(), 'sim_times': sims.values[:, 1].tolist(),
            'sf': float(sf), 'jacobian': np.asarray(J.values).tolist(),
            'model_jacobian': np.asarray(Jm.values).tolist(),
            'rss_gradient': g.tolist(), 'ssr': float(ssr)})
    G['project_1state_sf'] = out

    # ---- Project level, MM 'sum' mapping, no sf (tests/test_Project.py:263-349) ------------
    def michaelis(y, t, p):
        vmax, km, k_synt_s, k_deg_s, k_deg_p = p
        v = vmax * y[0] / (km + y[0])
        return [k_synt_s - v - k_deg_s * y[0], v - k_deg_p * y[1]]

    pmm = np.array([1e-3, 0.001, 0.01, 0.01, 0.001])
    tmm = np.linspace(0, 100, 20)
    sim = odeint(michaelis, [0.0, 0.0], tmm, args=(pmm,))
    total = TimecourseMeasurement('Total', np.sum(sim, axis=1), tmm)
    pj = Project(om2, [Experiment('Standard', total)], {}, {'Total': ('sum', [0, 1])})
    pd_ = {n: {'Global': v} for n, v in zip(op2, pmm)}
    th0 = np.log(pj.project_param_dict_to_vect(pd_))
    order = sorted(pj.project_param_idx, key=lambda n: pj.project_param_idx[n]['Global'])
    delta = {'vmax': 0.25, 'km': 0.1, 'k_synt_s': 0.3, 'k_deg_s': -0.2, 'k_deg_p': -0.15}
    th1 = th0 + np.array([delta[n] for n in order])
    out = {'project_order': order, 'data': np.sum(sim, axis=1).tolist(), 't': tmm.tolist(), 'points': []}
    for th in (th0, th1):
        r = np.asarray(pj.residuals(th))
        sims = pj.get_simulations()
        J = pj.calc_project_jacobian(th)
        Jv = np.array(J.values, copy=True)
        g = np.asarray(pj.calc_rss_gradient(th))
        out['points'].append({'theta': th.tolist(), 'residuals': r.tolist(),
                              'sims': sims.values[:, 0].tolist(), 'sim_times': sims.values[:, 1].tolist(),
                              'jacobian': Jv.tolist(), 'rss_gradient': g.tolist(),
                              'ssr': float(pj.calc_sum_square_residuals(th))})
    G['project_mm_sum'] = out

    # ---- Project level, two measurements in one experiment, shared sf group + priors --------
    # (tests/test_Project.py:359-375 set-up).  Residuals only: the reference's Jacobian for this
    # shape hits the row-offset defect (base_project.py:467,488) and a ValueError in
    # squared_loss_function.py:103 (hstack of 2-D blocks), recorded below.
    sub_m = TimecourseMeasurement('Substrate', sim[:, 0], tmm)
    prod_m = TimecourseMeasurement('Product', sim[:, 1], tmm)
    pj2 = Project(om2, [Experiment('Substrate Experiment', [sub_m, prod_m])], {},
                  {'Substrate': ('direct', 0), 'Product': ('direct', 1)},
                  sf_groups=[frozenset(['Substrate', 'Product'])])
    pj2.set_parameter_log_prior('k_synt_s', 'Global', np.log(0.01), 0.5)
    pj2.set_parameter_log_prior('k_deg_s', 'Global', np.log(0.01), 0.5)
    pj2.set_parameter_log_prior('k_deg_p', 'Global', np.log(0.01), 0.5)
    order2 = sorted(pj2.project_param_idx, key=lambda n: pj2.project_param_idx[n]['Global'])
    th = np.log(pj2.project_param_dict_to_vect(pd_)) + np.array([delta[n] for n in order2])
    r = np.asarray(pj2.residuals(th))
    sims = pj2.get_simulations(include_priors=True)
    out = {'project_order': order2, 'theta': th.tolist(), 'residuals': r.tolist(),
           'index': [list(x) for x in sims.index.tolist()], 'sims': sims.values[:, 0].tolist(),
           'sf': float(pj2.scale_factors['Product'].sf), 'substrate': sim[:, 0].tolist(),
           'product': sim[:, 1].tolist(), 't': tmm.tolist()}
    try:
        pj2.calc_project_jacobian(th)
        out['jacobian_error'] = None
    except Exception as exc:      # documented reference defect
        out['jacobian_error'] = type(exc).__name__
    G['project_mm_two_measures'] = out

    # ---- Project level, two measurements, NO sf: records the Jacobian row-offset defect ----
    pj3 = Project(om2, [Experiment('Substrate Experiment', [sub_m, prod_m])], {},
                  {'Substrate': ('direct', 0), 'Product': ('direct', 1)})
    order3 = sorted(pj3.project_param_idx, key=lambda n: pj3.project_param_idx[n]['Global'])
    th = np.log(pj3.project_param_dict_to_vect(pd_)) + np.array([delta[n] for n in order3])
    r = np.asarray(pj3.residuals(th))
    sims = pj3.get_simulations()
    J = np.array(pj3.calc_project_jacobian(th).values, copy=True)
    G['project_mm_two_measures_nosf'] = {
        'project_order': order3, 'theta': th.tolist(), 'residuals': r.tolist(),
        'index': [list(x) for x in sims.index.tolist()], 'sims': sims.values[:, 0].tolist(),
        'jacobian': J.tolist()}

    with open(OUT, 'w') as fh:
        json.dump(G, fh, indent=1)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
