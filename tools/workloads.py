"""The BASELINE.json configurations as synthetic projects, shared by ``bench.py`` and the GPU
parity tests so that what is timed is exactly what is checked.

The reference ships none of these models (only the 1- and 2-state fixtures of its tests); the model files live in
``sysbio_modeling_b200/model_files`` and go through the same symbolic front end as the reference's
(``symbolic/sympy_tools.py:272-397``).  Data are generated with the product's own integrator at the nominal parameters
and a seeded multiplicative noise; the CPU oracle receives the same arrays.

Every builder returns ``(project, make_oracle)``: ``make_oracle()`` builds the matching ``OracleProject`` (CPU
restatement of the reference, test / baseline infrastructure only - never on the product path).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

C4_OBS = {'MAPK_PP': 7, 'MKK_PP': 4}
C4_T = np.linspace(150.0, 3000.0, 20)
C3_T = np.logspace(-2, 3, 12)
C3_OBS = {'y1': 0, 'y3': 2}
C2_T = np.linspace(0.25, 4, 16)
C5_T = np.linspace(50, 600, 12)
C5_OBS = {'ERKPP': 24, 'RasT': 16}


def _symbolic(name):
    from sysbio_modeling_b200 import model_files
    from sysbio_modeling_b200.symbolic import symbolic_model_from_file
    return symbolic_model_from_file(model_files.path(name), name=name)


def nominal_of(name):
    from sysbio_modeling_b200 import model_files
    mod = model_files.module(name)
    sm = _symbolic(name)
    return sm, np.array([mod.nominal[p] for p in sm.param_order]), np.array(mod.initial_conditions, dtype=float)


def make_thetas(nominal, B, seed, spread):
    """theta = log(nominal) + spread * N(0, 1), ``np.random.default_rng(seed)``."""
    rng = np.random.default_rng(seed)
    return np.log(nominal)[None, :] + spread * rng.standard_normal((B, len(nominal)))


def _oracle_factory(name, n_vars, exps_fn, settings, mapping, **kw):
    def make():
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import copy
        import sbm_oracle
        from sysbio_modeling_b200.symbolic import make_ode_model
        sm = _symbolic(name)
        return sbm_oracle.OracleProject(make_ode_model(sm, False), make_ode_model(sm, True), n_vars, sm.param_order,
                                        exps_fn(), copy.deepcopy(settings), mapping, **kw)
    return make


# ---------------------------------------------------------------------------------------------------------
# C4: Kholodenko MAPK cascade, 1 experiment x 2 observables x 20 times, two scale-factor groups, full Jacobian
# ---------------------------------------------------------------------------------------------------------
def c4_experiments(data):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('mapk_kholodenko')
    meas = [sb.TimecourseMeasurement(k, data[k], C4_T, 0.05 * np.abs(data[k]) + 1e-3) for k in C4_OBS]
    return [sb.Experiment('MAPK_cascade', meas, initial_conditions=mod.initial_conditions)]


def _c4_project(tol, data, shape):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), model_name='mapk', group_lanes=shape[0],
                                     cols_per_lane=shape[1], sens_layout=shape[2])
    return sb.Project(om, c4_experiments(data), {}, {k: ('direct', v) for k, v in C4_OBS.items()},
                      sf_groups=list(C4_OBS.keys()), rtol=tol, atol=tol)


def c4_project(tol=1e-10, shape=(0, 0, 0)):
    """The headline project of ``bench.py``: data = the model at the nominal parameters x (1 + 0.05 N(0,1)), seed 40."""
    sm, nominal, _ = nominal_of('mapk_kholodenko')
    proj0 = _c4_project(tol, {k: np.ones_like(C4_T) for k in C4_OBS}, shape)
    sims = proj0.residuals_batch(np.log(nominal)[None, :])['sims'][0]
    rng = np.random.default_rng(40)
    labels = [l[1] for l in proj0._labels]
    data = {k: np.array([sims[i] for i, l in enumerate(labels) if l == k]) * (1 + 0.05 * rng.standard_normal(20))
            for k in C4_OBS}
    proj = _c4_project(tol, data, shape)
    mapping = {k: ('direct', v) for k, v in C4_OBS.items()}
    return proj, _oracle_factory('mapk_kholodenko', 8, lambda: c4_experiments(data), {}, mapping,
                                 sf_groups=list(C4_OBS.keys()), rtol=tol, atol=tol)


# ---------------------------------------------------------------------------------------------------------
# C3: Robertson, residuals only through the Rosenbrock kernel (rtol 1e-6, atol 1e-10), 12 log-spaced times
# ---------------------------------------------------------------------------------------------------------
def c3_experiments(data):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('robertson')
    meas = [sb.TimecourseMeasurement(k, data[k], C3_T, 0.05 * np.abs(data[k]) + 1e-6) for k in C3_OBS]
    return [sb.Experiment('Robertson', meas, initial_conditions=mod.initial_conditions)]


def c3_project(rtol=1e-6, atol=1e-10, solver='rodas4', data=None):
    """``data`` (dict observable -> values): reuse another C3 project's data (``project.workload_data``)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    sm, nominal, y0 = nominal_of('robertson')
    om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver=solver, rtol=rtol, atol=atol)
    if data is None:
        ynom = om.simulate(nominal, np.concatenate([[0.0], C3_T]), y0)[1:]
        rng = np.random.default_rng(30)
        data = {k: ynom[:, v] * (1 + 0.05 * rng.standard_normal(len(C3_T))) for k, v in C3_OBS.items()}
    mapping = {k: ('direct', v) for k, v in C3_OBS.items()}
    proj = sb.Project(om, c3_experiments(data), {}, mapping, solver=solver, rtol=rtol, atol=atol)
    proj.workload_data = data
    # the CPU oracle integrates with LSODA at a tolerance well below the kernel's
    return proj, _oracle_factory('robertson', 3, lambda: c3_experiments(data), {}, mapping, rtol=1e-10, atol=1e-13)


# ---------------------------------------------------------------------------------------------------------
# C2: reversible Michaelis-Menten, E experiments with experiment-local parameters, sensitivities
# ---------------------------------------------------------------------------------------------------------
def c2_experiments(E, seed=22):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('reversible_mm')
    r = np.random.default_rng(seed)
    out = []
    for k in range(E):
        data = 0.5 + 0.4 * r.random(len(C2_T))
        out.append(sb.Experiment('exp%05d' % k, sb.TimecourseMeasurement('P', data, C2_T, 0.05 * data + 1e-3),
                                 initial_conditions=mod.initial_conditions))
    return out


def c2_project(E=4096, shared_scale_factor=False, **kw):
    """C2: every experiment has its own four parameters (4 E project columns).  ``shared_scale_factor`` puts all
    experiments' 'P' rows into one scale-factor group - the coupling the experiment shard has to respect
    (abstract_loss_function.py:62-91)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    sm, nominal, _ = nominal_of('reversible_mm')
    settings = {'Local': list(sm.param_order)}
    om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
    groups = ['P'] if shared_scale_factor else None
    proj = sb.Project(om, c2_experiments(E), dict(settings), {'P': ('direct', 3)}, sf_groups=groups, **kw)
    return proj, _oracle_factory('reversible_mm', 4, lambda: c2_experiments(E), settings, {'P': ('direct', 3)},
                                 sf_groups=groups)


def c2_theta(E, n_sets=1, seed=2):
    _, nominal, _ = nominal_of('reversible_mm')
    rng = np.random.default_rng(seed)
    return np.tile(np.log(nominal), E)[None, :] + 0.3 * rng.standard_normal((n_sets, 4 * E))


# ---------------------------------------------------------------------------------------------------------
# C5: EGFR-scale network (25 species, 52 parameters), SSR / free energy of many walkers
# ---------------------------------------------------------------------------------------------------------
def c5_experiments(ynom):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('egfr_signalling')
    meas = [sb.TimecourseMeasurement(k, ynom[:, v], C5_T, 0.05 * ynom[:, v] + 0.1) for k, v in C5_OBS.items()]
    return [sb.Experiment('EGF_stimulation', meas, initial_conditions=mod.initial_conditions)]


def c5_project(**kw):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    sm, nominal, y0 = nominal_of('egfr_signalling')
    om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'))
    ynom = om.simulate(nominal, np.concatenate([[0.0], C5_T]), y0)[1:]
    mapping = {k: ('direct', v) for k, v in C5_OBS.items()}
    proj = sb.Project(om, c5_experiments(ynom), {}, mapping, **kw)
    return proj, _oracle_factory('egfr_signalling', 25, lambda: c5_experiments(ynom), {}, mapping)
