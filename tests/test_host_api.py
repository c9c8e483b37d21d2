"""CPU: host-side mirror of the reference's data classes and Project bookkeeping.  These follow the
reference's own unit tests (tests/test_Experiment.py, test_Measurements.py, test_Project_Utils.py and the
set-up assertions of test_Project.py:78-94, 215-252)."""
import numpy as np
import pytest

import sysbio_modeling_b200 as sb
from sysbio_modeling_b200.project.utils import OrderedHashDict
from fixtures import project_1state, project_mm_two_measures, experiments_1state, SETTINGS_1STATE


# ---- measurements / experiments --------------------------------------------------------------
def test_measurement_validation():
    t = np.linspace(0, 10, 5)
    with pytest.raises(ValueError):
        sb.TimecourseMeasurement('a', np.ones(5), t, np.array([1, 1, 0, 1, 1.0]))
    with pytest.raises(ValueError):
        sb.TimecourseMeasurement('a', np.ones(5), t[:4])
    with pytest.raises(ValueError):
        sb.TimecourseMeasurement('a', np.ones(5), t, np.ones(4))
    m = sb.TimecourseMeasurement('a', np.arange(5.0), t)
    assert np.all(m.std == 1)
    v, s, tt = m.get_nonzero_measurements()
    assert len(v) == 4 and tt[0] == 2.5
    m.drop_timepoint_zero()
    assert len(m.values) == len(m.std) == len(m.timepoints) == 4


def test_experiment_sorting_and_timepoints():
    t1, t2 = np.array([0., 1, 2]), np.array([0.5, 2, 3])
    mb = sb.TimecourseMeasurement('b', np.ones(3), t1)
    ma = sb.TimecourseMeasurement('a', np.ones(3), t2)
    e = sb.Experiment('Exp1', [mb, ma], experiment_settings={'x': 1})
    assert [m.variable_name for m in e.measurements] == ['a', 'b']
    assert np.array_equal(e.get_unique_timepoints(), [0.5, 1, 2, 3])
    assert np.array_equal(e.get_unique_timepoints(include_zero=True), [0, 0.5, 1, 2, 3])
    with pytest.raises(KeyError):
        e.add_measurement(sb.TimecourseMeasurement('a', np.ones(3), t1))
    with pytest.raises(KeyError):
        e.get_variable_measurements('zzz')
    with pytest.raises(ValueError):
        sb.Experiment('_bad', mb)
    e.drop_timepoint_zero('b')
    assert len(e.get_variable_measurements('b').timepoints) == 2


# ---- OrderedHashDict (reference tests/test_Project_Utils.py) ----------------------------------
def test_hashdict_semantics():
    d = OrderedHashDict()
    with pytest.raises(TypeError):
        d[5] = 'x'
    with pytest.raises(TypeError):
        d[('a', 'b')] = 'x'
    with pytest.raises(TypeError):
        d[frozenset(['hi', 3])] = 'x'
    d['hi'] = 'hello'
    d[frozenset(['a', 'b'])] = 'set'
    assert d['a'] == 'set' and d['b'] == 'set' and 'a' in d and 'zz' not in d
    with pytest.raises(KeyError):
        d[frozenset(['c', 'b'])] = 'other'
    with pytest.raises(KeyError):
        d['a'] = 'other'
    with pytest.raises(KeyError):
        d[frozenset(['hi', 'q'])] = 'other'
    d['hi'] = 'again'
    assert d['hi'] == 'again'
    d[frozenset(['a', 'b'])] = 'set2'
    assert d['b'] == 'set2'
    keys = [chr(65 + i) * 3 for i in range(20)]
    e = OrderedHashDict()
    for k in keys:
        e[k] = 1
    assert list(e.keys()) == keys
    with pytest.raises(KeyError):
        e['nope']


# ---- Project bookkeeping ------------------------------------------------------------------------
def test_project_initialisation_and_plan():
    proj, oracle = project_1state()
    assert list(proj.get_param_index('k_synt').keys()) == ['Global']
    assert len(proj.get_param_index('Group_1')) == 2
    idx = proj.project_param_idx
    for e in proj.experiments:
        assert idx['Group_1'][(e.settings['Deg_Rate'],)] == e.param_global_vector_idx['k_deg']
    assert proj.n_project_residuals == 35 and proj.n_project_params == 3
    assert proj.project_param_idx == oracle.project_param_idx
    assert proj.get_param_index('Group_1', ('Low',)) == 2
    t = proj.plan_tables
    assert (t.E, t.n_rows, t.n_proj, t.n_groups, t.jac_atomic) == (2, 35, 3, 1, 0)
    # High_Deg_Exp sorts first; its rows 0..25, Low_Deg_Exp rows 26..34 (t = 0 dropped)
    assert list(t.tout_off) == [0, 26, 35]
    assert np.array_equal(t.pmap, [[1, 0], [2, 0]])
    # snapped times are nodes of linspace(0, t_end, 1000) at or after the measurement time
    high = [e for e in proj.experiments if e.name == 'High_Deg_Exp'][0]
    tm = high.get_variable_measurements('Variable_1').timepoints
    grid = np.linspace(0, tm[-1], 1000)
    assert np.all(np.isin(t.tout[:26], grid)) and np.all(t.tout[:26] >= tm) and np.all(t.tout[:26] - tm <= grid[1])
    assert np.array_equal(oracle.meas_mean, t.data) and np.array_equal(oracle.meas_std, t.sigma)
    df = proj.measurements_df
    assert list(df.columns) == ['mean', 'std', 'timepoints'] and len(df) == 35
    with pytest.raises(ValueError):
        proj.get_simulations()


def test_project_rows_with_priors_match_oracle(golden):
    proj, oracle = project_mm_two_measures(golden)
    assert [tuple(l) for l in proj._labels] == [tuple(l) for l in oracle.labels]
    assert [list(l) for l in proj._labels] == golden['project_mm_two_measures']['index']
    t = proj.plan_tables
    assert t.n_rows == 41 and t.n_groups == 1
    assert list(t.row_kind[-3:]) == [1, 1, 1]
    # prior rows: data sorted by label, simulated value (theta column) in insertion order (reference quirk)
    order = [n.split(' ')[0] for n in proj.get_ordered_project_params()]
    assert [order[c] for c in t.row_aux[-3:]] == ['k_synt_s', 'k_deg_s', 'k_deg_p']
    assert sorted(t.grp_rows.tolist()) == list(range(38))
    # reference-compat Jacobian rows fold the second measurement onto the first and need atomics
    proj2, _ = project_mm_two_measures(golden, with_sf=False, with_priors=False, reference_jac_rows=True)
    t2 = proj2.plan_tables
    assert t2.jac_atomic == 1 and set(t2.row_jac.tolist()) == set(range(19))
    assert sorted(t2.row_sim.tolist()) == list(range(38))


def test_add_and_remove_experiment():
    proj, _ = project_1state()
    t = np.linspace(0, 100, 10)
    e3 = sb.Experiment('Simple_Experiment', sb.TimecourseMeasurement('Variable_1', np.linspace(1, 2, 10), t),
                       experiment_settings={'Deg_Rate': 'Very High'})
    proj.add_experiment(e3)
    assert 'Simple_Experiment' in [e.name for e in proj.experiments]
    assert proj.n_project_params == 4 and proj.n_project_residuals == 44
    with pytest.raises(KeyError):
        proj.add_experiment(e3)
    proj.remove_experiments_by_settings({'Deg_Rate': 'Very High'})
    assert 'Simple_Experiment' not in [e.name for e in proj.experiments]
    assert proj.n_project_params == 3
    with pytest.raises(KeyError):
        proj.remove_experiments_by_settings({'Absent': 5})
    with pytest.raises(KeyError):
        proj.remove_experiments_by_settings({'Deg_Rate': 'Nope'})
    assert proj.get_experiment_index('Low_Deg_Exp') == 1
    with pytest.raises(KeyError):
        proj.get_experiment_index('zzz')


def test_param_dict_vector_round_trip_and_errors():
    proj, _ = project_1state()
    d = {'k_synt': {'Global': 0.5}, 'Group_1': {('High',): 1.5, ('Low',): 2.5}}
    v = proj.project_param_dict_to_vect(d)
    assert list(v) == [0.5, 1.5, 2.5]
    assert proj.project_param_vect_to_dict(v) == d
    assert list(proj.project_param_dict_to_vect({}, default_value=7.0)) == [7.0] * 3
    with pytest.raises(KeyError):
        proj.set_parameter_log_prior('nope', 'Global', 0.0, 1.0)
    om = proj._model
    with pytest.raises(ValueError):
        sb.Project(om, experiments_1state(), dict(SETTINGS_1STATE), {'Variable_1': ('direct', 3)})
    with pytest.raises(NotImplementedError):
        sb.Project(om, experiments_1state(), dict(SETTINGS_1STATE), {'Variable_1': ('custom', (None, None, None))})
    with pytest.raises(ValueError):
        sb.Project(om, experiments_1state(), dict(SETTINGS_1STATE), {'Variable_1': ('weird', 0)})
    g = proj.group_experiments(['Deg_Rate'])
    assert set(g.keys()) == {('High',), ('Low',)}


def test_local_and_fixed_parameters():
    """'Local' parameters get one column per experiment; experiment-fixed parameters must match the model's
    fixed set (one sensitivity model per OdeModel, reference sympy_tools.py:137-139)."""
    from sysbio_modeling_b200 import model_files
    t = np.array([1.0, 2.0, 3.0])
    mk = lambda n, fixed=None: sb.Experiment(n, sb.TimecourseMeasurement('P', np.ones(3), t), fixed_parameters=fixed,
                                             initial_conditions=[1.0, 0.1, 0.0, 0.0])
    om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
    proj = sb.Project(om, [mk('E1'), mk('E0')], {'Local': ['k1f', 'k1r', 'k2f', 'k2r']}, {'P': ('direct', 3)})
    assert proj.n_project_params == 8
    assert proj.project_param_idx['k1f_E0'] == {'Local': 0} and proj.project_param_idx['k2r_E1'] == {'Local': 7}
    assert np.array_equal(proj.plan_tables.pmap, [[0, 1, 2, 3], [4, 5, 6, 7]])
    assert np.array_equal(proj.plan_tables.y0[0], [1.0, 0.1, 0.0, 0.0])
    omf = sb.OdeModel.from_model_file(model_files.path('reversible_mm'), fixed_params=['k2r'])
    projf = sb.Project(omf, [mk('E0', {'k2r': 0.1})], {'Fixed': ['k2r']}, {'P': ('direct', 3)})
    assert projf.n_project_params == 3 and projf.plan_tables.pmap[0, 3] == -1 and projf.plan_tables.pfix[0, 3] == 0.1
    with pytest.raises(ValueError):
        sb.Project(om, [mk('E0', {'k2r': 0.1})], {'Fixed': ['k2r']}, {'P': ('direct', 3)})   # model has a k2r column
    with pytest.raises(ValueError):
        sb.Project(omf, [mk('E0')], {'Fixed': ['k2r']}, {'P': ('direct', 3)})                # no value provided
