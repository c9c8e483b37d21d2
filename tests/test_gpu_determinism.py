"""Race detector for the CTA-transposed sensitivity kernel (compute-sanitizer is not available on this pool).

The kernel hands J and G from the model warp to the column warps through double-buffered shared memory ordered by one
CTA barrier per stage, keeps two vectors per column in tensor memory (tcgen05.ld / st ordered by tcgen05.wait), and lets
lanes fetch trajectories from a global queue.  A missing barrier or fence between a producer and a consumer would make
the results depend on timing.  So: the same launch, over and over, must give bitwise identical results

  * under random delays injected before every CTA barrier, every block of tensor-memory transfers and every share of
    the model (-DSBM_CT_JITTER: up to 2 us per site, a different pattern in every launch),
  * with the persistent grid forced to 1 / 37 / 148 CTAs (which trajectories share a CTA, and when, changes),
  * with the stage's buffer index passed into sbm_ct_primal as an argument instead of being derived from the stage number
    (-DSBM_CT_BUF_ARG=1: the variant DESIGN.md section 6 of round 1 reported as changing results),

and all of them must equal the plain build bit for bit."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _solve(defines, max_ctas, nb, reps, tol=1e-8):
    import torch
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files, engine
    from sysbio_modeling_b200.symbolic import symbolic_model_from_file
    from tools import workloads as wl
    old = {k: os.environ.get(k) for k in ('SBM_KERNEL_DEFINES', 'SBM_CT_MAX_CTAS')}
    try:
        if defines:
            os.environ['SBM_KERNEL_DEFINES'] = defines
        else:
            os.environ.pop('SBM_KERNEL_DEFINES', None)
        sm = symbolic_model_from_file(model_files.path('mapk_kholodenko'), name='mapk')
        cm = engine.CompiledModel(sm)                       # not the process-wide cache: the defines are part of the build
        om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), model_name='mapk', rtol=tol, atol=tol)
        om._compiled = cm
        assert cm.info()['sens_layout'] == 2
        if max_ctas:
            os.environ['SBM_CT_MAX_CTAS'] = str(max_ctas)
        else:
            os.environ.pop('SBM_CT_MAX_CTAS', None)
        sm2, nominal, y0 = wl.nominal_of('mapk_kholodenko')
        P = np.exp(wl.make_thetas(nominal, nb, 4, 0.25))
        t = np.concatenate([[0.0], wl.C4_T])
        outs = []
        for _ in range(reps):
            y, S = om.calc_jacobian_batch(P, t, y0)
            assert (om.last_status == 0).all()
            outs.append((y.copy(), S.copy(), dict(om.last_counters)))
        return outs
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_results_do_not_depend_on_timing():
    nb = 6000                                               # more than one wave of 148 x 32 lanes: the queue refills
    ref = _solve('', 0, nb, 1)[0]
    assert np.isfinite(ref[1]).all()
    # jitter: 10 launches, every one with its own delay pattern
    for k, (y, S, cnt) in enumerate(_solve('-DSBM_CT_JITTER=64', 0, nb, 10)):
        assert np.array_equal(y, ref[0]) and np.array_equal(S, ref[1]), 'launch %d under jitter differs' % k
        assert cnt['steps'] == ref[2]['steps'] and cnt['rejects'] == ref[2]['rejects']
    # grid size: which trajectories share a CTA changes, the arithmetic of a trajectory must not
    for ctas, n in ((1, 96), (37, 3000), (148, nb)):
        base = _solve('', 0, n, 1)[0]
        for y, S, cnt in _solve('-DSBM_CT_JITTER=64', ctas, n, 2):
            assert np.array_equal(y, base[0]) and np.array_equal(S, base[1]), 'grid of %d CTAs differs' % ctas
    # the buffer index as an argument (round 1's code-generation hazard): same results, same step counts
    for y, S, cnt in _solve('-DSBM_CT_BUF_ARG=1', 0, nb, 2) + _solve('-DSBM_CT_BUF_ARG=1 -DSBM_CT_JITTER=64', 0, nb, 2):
        assert np.array_equal(y, ref[0]) and np.array_equal(S, ref[1])
        assert cnt['steps'] == ref[2]['steps']
