"""Shared test fixtures: the reference's own test projects (data from reference tests/test_Project.py:29-63
and :283-306), built twice - through the product API (CUDA) and through the CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

import sysbio_modeling_b200 as sb                                    # noqa: E402
from sysbio_modeling_b200 import model_files                          # noqa: E402
from sysbio_modeling_b200.symbolic import symbolic_model_from_file, make_ode_model   # noqa: E402
import sbm_oracle                                                     # noqa: E402

SETTINGS_1STATE = {'Global': ['k_synt'], 'Shared': {'Group_1': {'k_deg': ('Deg_Rate',)}}}


def data_1state():
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


def experiments_1state():
    t1, low, t2, high = data_1state()
    e_low = sb.Experiment('Low_Deg_Exp', sb.TimecourseMeasurement('Variable_1', low, t1),
                          experiment_settings={'Deg_Rate': 'Low'})
    e_high = sb.Experiment('High_Deg_Exp', sb.TimecourseMeasurement('Variable_1', high, t2),
                           experiment_settings={'Deg_Rate': 'High'})
    return [e_low, e_high]


def python_rhs(name, fixed=None):
    sm = symbolic_model_from_file(model_files.path(name), name=name, fixed_params=fixed)
    return sm, make_ode_model(sm, False), make_ode_model(sm, True)


def project_1state(loss=None, **kw):
    """(product Project, OracleProject) for reference TestProject.setUpClass."""
    import copy
    om = sb.OdeModel.from_model_file(model_files.path('simple_synth_deg'), model_name='simple')
    mapping = {'Variable_1': ('direct', 0)}
    groups = [frozenset(['Variable_1'])]
    args = {} if loss is None else {'loss_function': loss}
    proj = sb.Project(om, experiments_1state(), copy.deepcopy(SETTINGS_1STATE), mapping, sf_groups=groups,
                      **args, **kw)
    sm, m, s = python_rhs('simple_synth_deg')
    kind = getattr(loss, 'kind', 'square') if loss is not None else 'square'
    oracle = sbm_oracle.OracleProject(m, s, 1, sm.param_order, experiments_1state(), copy.deepcopy(SETTINGS_1STATE),
                                      mapping, sf_groups=groups, loss=kind)
    return proj, oracle


def mm_data(golden):
    g = golden['project_mm_sum']
    return np.array(g['t']), np.array(g['data'])


def project_mm_sum(golden, **kw):
    """Reference tests/test_Project.py:283-306: MM model, 'sum' mapping, no scale factors."""
    t, total = mm_data(golden)
    exps = lambda: [sb.Experiment('Standard', sb.TimecourseMeasurement('Total', total, t))]
    om = sb.OdeModel.from_model_file(model_files.path('michaelis_menten'), model_name='mm')
    mapping = {'Total': ('sum', [0, 1])}
    proj = sb.Project(om, exps(), {}, mapping, **kw)
    sm, m, s = python_rhs('michaelis_menten')
    oracle = sbm_oracle.OracleProject(m, s, 2, sm.param_order, exps(), {}, mapping)
    return proj, oracle


def project_mm_two_measures(golden, with_sf=True, with_priors=True, reference_jac_rows=False, **kw):
    """Reference tests/test_Project.py:359-375: two measurements in one experiment, shared sf group, priors."""
    g = golden['project_mm_two_measures']
    t = np.array(g['t'])

    def exps():
        return [sb.Experiment('Substrate Experiment',
                              [sb.TimecourseMeasurement('Substrate', np.array(g['substrate']), t),
                               sb.TimecourseMeasurement('Product', np.array(g['product']), t)])]
    mapping = {'Substrate': ('direct', 0), 'Product': ('direct', 1)}
    groups = [frozenset(['Substrate', 'Product'])] if with_sf else None
    om = sb.OdeModel.from_model_file(model_files.path('michaelis_menten'), model_name='mm')
    proj = sb.Project(om, exps(), {}, mapping, sf_groups=groups, reference_jac_rows=reference_jac_rows, **kw)
    sm, m, s = python_rhs('michaelis_menten')
    oracle = sbm_oracle.OracleProject(m, s, 2, sm.param_order, exps(), {}, mapping, sf_groups=groups,
                                      reference_jac_rows=reference_jac_rows)
    if with_priors:
        for name in ('k_synt_s', 'k_deg_s', 'k_deg_p'):
            proj.set_parameter_log_prior(name, 'Global', np.log(0.01), 0.5)
            oracle.set_parameter_log_prior(name, 'Global', np.log(0.01), 0.5)
    return proj, oracle


def theta_by_name(order_to, names_from, values):
    """Re-order a parameter vector given in ``names_from`` order into ``order_to`` order."""
    d = dict(zip(names_from, values))
    return np.array([d[n] for n in order_to])
