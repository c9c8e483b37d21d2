"""Golden vectors for the sampler-side pieces, from the REAL reference run in this container:

  * ``project/Ensembles.py``: ``_sampling_matrix`` (:226-264), ``_quadratic_cost`` (:272-282), ``autocorrelation`` (:14-29);
  * ``project/loss_functions/squared_loss/linear_scale_factor.py``: ``_entropy_integrand`` (:13-18) and
    ``LinearScaleFactor.calc_scale_factor_entropy`` (:63-80).

Ensembles.py is Python 2 and written against the numpy aliases SciPy has since removed (``scipy.transpose``, ``scipy.sqrt``,
``scipy.compress``, ``scipy.inf`` ...).  The function sources are taken from the files as they lie and executed with a ``scipy``
name that falls back to numpy for those aliases - no numerical change.  ``/root/reference`` does not exist on the GPU box: the
numbers are committed as ``tests/golden/reference_golden_ensembles.json``.

Run:  python tests/golden/make_golden_ensembles.py        (only works where /root/reference exists)
"""
import json
import os
import re
import types

import numpy as np
import scipy
import scipy.fftpack
import scipy.integrate
import scipy.linalg

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_golden_ensembles.json')


class ScipyShim(types.ModuleType):
    """``scipy`` as Ensembles.py expects it: real sub-packages, numpy for the removed aliases."""
    def __getattr__(self, name):
        if hasattr(scipy, name):
            return getattr(scipy, name)
        return getattr(np, name)


def function_source(text, name):
    m = re.search(r'^def %s\(.*?(?=^def |^class |\Z)' % re.escape(name), text, flags=re.S | re.M)
    if not m:
        raise RuntimeError('no function %s' % name)
    return m.group(0)


def main():
    shim = ScipyShim('scipy')
    ns = {'scipy': shim}
    ens_src = open(os.path.join(REF, 'project', 'Ensembles.py')).read()
    for fn in ('autocorrelation', '_sampling_matrix', '_quadratic_cost'):
        exec(function_source(ens_src, fn), ns)
    lsf_src = open(os.path.join(REF, 'project', 'loss_functions', 'squared_loss', 'linear_scale_factor.py')).read()
    integrand_src = function_source(lsf_src, '_entropy_integrand')
    ns2 = {'np': np, 'scipy': shim}
    exec(integrand_src, ns2)
    # the method body of calc_scale_factor_entropy, as a function of (self, ...)
    m = re.search(r'    def calc_scale_factor_entropy\(self.*?(?=^    @property|^    def |\Z)', lsf_src, flags=re.S | re.M)
    body = '\n'.join(line[4:] if line.startswith('    ') else line for line in m.group(0).split('\n'))
    exec(body, ns2)

    golden = {'sampling_matrix': [], 'entropy': [], 'quadratic_cost': [], 'autocorrelation': {}}
    rng = np.random.default_rng(2024)
    for n, cutoff, T, scale in ((3, 0, 1.0, 1.0), (5, 1e-3, 2.5, 0.5), (8, 0.05, 1e-3, 0.7), (22, 1e-4, 1.0, 1.0)):
        A = rng.standard_normal((2 * n, n)) * np.exp(1.5 * rng.standard_normal(n))[None, :]
        H = A.T.dot(A)
        S = ns['_sampling_matrix'](H, cutoff, T, scale)
        step = rng.standard_normal(n)
        golden['sampling_matrix'].append({'hessian': H.tolist(), 'cutoff': cutoff, 'temperature': T, 'step_scale': scale,
                                          'samp_mat': np.asarray(S).tolist()})
        golden['quadratic_cost'].append({'step': step.tolist(), 'hessian': H.tolist(),
                                         'cost': float(ns['_quadratic_cost'](step, H))})
    series = np.sin(np.arange(64) / 5.0) + 0.1 * rng.standard_normal(64)
    golden['autocorrelation'] = {'series': series.tolist(), 'ac': np.asarray(ns['autocorrelation'](series)).tolist()}

    class Holder(object):
        pass
    for mu, sg, T, amp in ((np.log(3.5), 0.4, 1.0, 1.0), (np.log(3.5), 0.4, 2e-3, 1.0), (0.0, 2.0, 0.5, 0.3),
                           (1.0, 0.05, 1.0, 4.0), (-2.0, 1.0, 10.0, 1.0)):
        sim = amp * (1.0 - np.exp(-np.linspace(0.2, 3.0, 12)))
        data = 3.3 * sim * (1 + 0.05 * rng.standard_normal(12))
        std = 0.1 + 0.05 * np.abs(data)
        h = Holder()
        h._sf, h.log_prior, h.log_prior_sigma = 1.0, mu, sg
        ent = ns2['calc_scale_factor_entropy'](h, sim, data, std, T)
        golden['entropy'].append({'sim': sim.tolist(), 'data': data.tolist(), 'std': std.tolist(), 'log_prior': mu,
                                  'log_prior_sigma': sg, 'temperature': T, 'entropy': float(ent), 'sf': float(h._sf)})
    json.dump(golden, open(OUT, 'w'), indent=0)
    print('wrote', OUT, {k: len(v) for k, v in golden.items()})


if __name__ == '__main__':
    main()
