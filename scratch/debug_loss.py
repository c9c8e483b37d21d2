import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle'))
import numpy as np, pandas as pd
import sbm_oracle as O
from sysbio_modeling_b200 import SquareLossFunction, LogSquareLossFunction
np.set_printoptions(precision=6, linewidth=200)
def _model3(p, t):
    a, b, c = p
    return a * np.sin(b * t) - c * t ** 2.0
def _frames3(vals, t):
    sim = pd.DataFrame({'mean': vals, 'timecourse': t})
    sim.index = pd.MultiIndex.from_tuples([('Experiment_1', 'Val')] * len(t))
    return sim
t = np.linspace(0, 10, 11)
p1 = np.array([0.3, 0.5, 1.3])
sim = _frames3(_model3(p1, t), t)
jac = pd.DataFrame(np.stack([np.sin(p1[1] * t), p1[0] * t * np.cos(p1[1] * t), -t ** 2], axis=1),
                   columns=['a', 'b', 'c'], index=sim.index)
measures = sim.copy()
measures.insert(1, 'std', np.ones(11))
lf = SquareLossFunction(sf_groups=['Val'])
measures['mean'] = measures['mean'] * 5 - np.random.default_rng(1).standard_normal(11)
def sf_of(p):
    lf.update_scale_factors(_frames3(_model3(p, t), t), measures)
    return np.array([lf.scale_factors['Val'].sf])
out = lf._run_frames(sim, measures, jac)
print('sf', out['sf'], 'sfgrad', out['sfgrad'])
s = sim['mean'].values; d = measures['mean'].values
print('expected sf', (s*d).sum()/(s*s).sum())
print('FD', O.central_difference(sf_of, p1))
print('tables', vars(lf._frame_tables(sim, measures, 3)))
