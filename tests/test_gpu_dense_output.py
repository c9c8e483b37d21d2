"""Dense output of the states-only Dormand-Prince kernel (sbm_dp5_plain_dense): on the reference's own call shape -
``OdeModel.simulate(p, np.linspace(0, t_end, 1000))`` (base_project.py:419-420) - output times are read off the 4th-order
continuous extension of the step that passes them instead of clipping a step onto every one of the 1,000 nodes."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _models():
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    out = {}
    for mode in (1, 2):
        om = sb.OdeModel.from_model_file(model_files.path('michaelis_menten'), model_name='mm', rtol=1e-10, atol=1e-10)
        om.dense_output = mode
        out[mode] = om
    return out


def test_dense_grid_matches_step_landing_and_takes_fewer_steps():
    oms = _models()
    rng = np.random.default_rng(8)
    P = np.array([1e-3, 0.001, 0.01, 0.01, 0.001]) * np.exp(0.3 * rng.standard_normal((64, 5)))
    t = np.linspace(0, 500.0, 1000)
    y0 = np.array([1.0, 0.0])
    yd = oms[1].simulate_batch(P, t, y0)
    steps_dense = oms[1].last_counters['steps']
    assert (oms[1].last_status == 0).all()
    yc = oms[2].simulate_batch(P, t, y0)
    steps_clip = oms[2].last_counters['steps']
    assert (oms[2].last_status == 0).all()
    ratio = np.abs(yd - yc) / (1e-10 + 1e-10 * np.abs(yc))
    print('steps: dense %d, landing on every node %d; worst difference %.2f x (atol + rtol |y|)' % (steps_dense, steps_clip, ratio.max()))
    assert steps_dense * 3 < steps_clip                      # 1,000 nodes no longer dictate the step count
    assert ratio.max() <= 10.0
    # the automatic choice: dense from 256 output times per trajectory on, landing below
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    om = sb.OdeModel.from_model_file(model_files.path('michaelis_menten'), model_name='mm', rtol=1e-10, atol=1e-10)
    assert np.array_equal(om.simulate_batch(P[:4], t, y0), yd[:4])
    t_few = np.linspace(0, 500.0, 40)
    a = om.simulate_batch(P[:4], t_few, y0)
    assert np.array_equal(a, oms[2].simulate_batch(P[:4], t_few, y0))


def test_dense_output_against_the_closed_form():
    """y' = k_synt - k_deg y: y(t) = k_synt / k_deg (1 - exp(-k_deg t)), at 1,000 nodes within 10 x tolerance."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    om = sb.OdeModel.from_model_file(model_files.path('simple_synth_deg'), model_name='simple', rtol=1e-10, atol=1e-10)
    om.dense_output = 1
    order = list(om.param_order)
    ks, kd = 0.8, 0.05
    p = np.array([ks if n == 'k_synt' else kd for n in order])
    t = np.linspace(0, 100.0, 1000)
    y = om.simulate(p, t, np.zeros(1))[:, 0]
    exact = ks / kd * (1 - np.exp(-kd * t))
    assert om.last_counters['steps'] < 400
    assert np.all(np.abs(y - exact) <= 10 * (1e-10 + 1e-10 * np.abs(exact))), np.abs(y - exact).max()


def test_rodas4_dense_output_on_the_robertson_model():
    """The Rosenbrock kernel's continuous extension (3rd order) on 1,000 nodes: far fewer steps than landing on every node,
    values within the method's accuracy of the landing kernel and of LSODA at a tight tolerance."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    import sbm_oracle as O
    from fixtures import python_rhs
    mod = model_files.module('robertson')
    sm, m, s = python_rhs('robertson')
    nominal = np.array([mod.nominal[k] for k in sm.param_order])
    P = nominal * np.exp(0.3 * np.random.default_rng(4).standard_normal((16, 3)))
    t = np.linspace(0, 1000.0, 1000)
    y0 = np.array(mod.initial_conditions, dtype=float)
    out = {}
    for mode in (1, 2):
        om = sb.OdeModel.from_model_file(model_files.path('robertson'), solver='rodas4', rtol=1e-7, atol=1e-11)
        om.dense_output = mode
        out[mode] = (om.simulate_batch(P, t, y0), om.last_counters['steps'])
        assert (om.last_status == 0).all()
    (yd, nd), (yc, nc) = out[1], out[2]
    rel = np.abs(yd - yc) / (1e-11 + 1e-7 * np.abs(yc))
    print('steps: dense %d, landing %d; worst difference %.1f x (atol + rtol |y|)' % (nd, nc, rel.max()))
    assert nd * 4 < nc
    assert rel.max() <= 50.0
    yr = O.simulate(m, 3, P[0], t, y0, rtol=1e-12, atol=1e-14)
    assert np.all(np.abs(yd[0] - yr) <= 1e-5 * np.abs(yr) + 1e-9)
    assert np.abs(yd.sum(axis=2) - 1.0).max() < 1e-9          # the interpolant keeps the linear invariant
