"""Thin Python layer over the C ABI: compiled models and device plans.

``CompiledModel`` = emitted CUDA source + NVRTC-compiled integrator kernels (``sbm_model``).
``DevicePlan``    = flat numeric form of a project on the device (``sbm_plan``) with the evaluation
                    entry points: host-buffer calls (numpy in / numpy out, what the reference-facing
                    ``Project`` methods use) and device-pointer calls (torch tensors, what the batch
                    API and the benchmark use).

Nothing here computes on the CPU; if the native library or a CUDA device is missing the calls raise.
"""
import ctypes
import os
import weakref

import numpy as np

from . import _lib
from ._lib import SbmError, SolverOpts, PlanDesc, ModelDesc, ModelInfo
from .symbolic import emit_cuda, SymbolicModel

SOLVERS = {'dp5': _lib.SBM_SOLVER_DP5, 'rodas4': _lib.SBM_SOLVER_RODAS4, 'rodas': _lib.SBM_SOLVER_RODAS4}
LOSSES = {'square': _lib.SBM_LOSS_SQUARE, 'log': _lib.SBM_LOSS_LOG, 'normalized': _lib.SBM_LOSS_NORMALIZED}


def _ptr(arr):
    return arr.ctypes.data_as(ctypes.c_void_p)


class CompiledModel(object):
    """Emit + NVRTC-compile the model (reference analogue: numba ``enable_jit``, ode_model.py:46-74)."""

    def __init__(self, symbolic, group_lanes=0, cols_per_lane=0, block_threads=0, sens_layout=0,
                 cache_dir=_lib.CACHE_DIR):
        if not isinstance(symbolic, SymbolicModel):
            raise TypeError('symbolic must be a SymbolicModel')
        self.symbolic = symbolic
        if not sens_layout:
            sens_layout = int(os.environ.get('SBM_SENS_LAYOUT', '0'))
        self.cuda = emit_cuda(symbolic, ct_cols=cols_per_lane if sens_layout == 2 else 0)
        self.n_vars, self.n_params, self.n_free = self.cuda.n_vars, self.cuda.n_params, self.cuda.n_free
        L = _lib.lib()
        cache = cache_dir.encode() if cache_dir else None
        desc = ModelDesc(self.n_vars, self.n_params, self.n_free, group_lanes, cols_per_lane, block_threads,
                         sens_layout, None, cache)
        h = ctypes.c_void_p()
        _lib.check(L.sbm_model_compile(self.cuda.source.encode(), ctypes.byref(desc), ctypes.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, L.sbm_model_destroy, h)

    @property
    def handle(self):
        return self._h

    def info(self):
        out = ModelInfo()
        _lib.check(_lib.lib().sbm_model_info(self._h, ctypes.byref(out)))
        return out.as_dict()

    @property
    def log(self):
        return _lib.lib().sbm_model_log(self._h).decode(errors='replace')


_MODEL_CACHE = {}


def compiled_model_for(symbolic, **kw):
    key = (symbolic.fingerprint(), tuple(sorted(kw.items())))
    m = _MODEL_CACHE.get(key)
    if m is None:
        m = CompiledModel(symbolic, **kw)
        _MODEL_CACHE[key] = m
    return m


class PlanTables(object):
    """Host arrays of a plan (field names = ``sbm_plan_desc_t``); built by Project / OdeModel."""

    INT = ('pmap', 'tout_off', 'orow_off', 'row_sim', 'row_jac', 'row_group', 'row_kind', 'row_aux',
           'grp_off', 'grp_rows')
    DBL = ('pfix', 'y0', 'tout', 'data', 'sigma')

    def __init__(self, **kw):
        self.E = int(kw.pop('E'))
        self.n_rows = int(kw.pop('n_rows'))
        self.n_proj = int(kw.pop('n_proj'))
        self.n_groups = int(kw.pop('n_groups', 0))
        self.jac_atomic = int(kw.pop('jac_atomic', 0))
        for name in self.INT:
            setattr(self, name, np.ascontiguousarray(kw.pop(name, np.zeros(0)), dtype=np.int32))
        for name in self.DBL:
            setattr(self, name, np.ascontiguousarray(kw.pop(name, np.zeros(0)), dtype=np.float64))
        self.row_mask = np.ascontiguousarray(kw.pop('row_mask', np.zeros(0)), dtype=np.uint64)
        if kw:
            raise TypeError('unknown plan fields: %s' % ', '.join(kw))
        if len(self.grp_off) == 0:
            self.grp_off = np.zeros(1, dtype=np.int32)
        self.n_tout = len(self.tout)
        self.n_orow = len(self.row_sim)

    def desc(self):
        d = PlanDesc()
        d.E, d.n_rows, d.n_proj, d.n_groups = self.E, self.n_rows, self.n_proj, self.n_groups
        d.n_tout, d.n_orow, d.jac_atomic = self.n_tout, self.n_orow, self.jac_atomic
        for name in self.INT:
            setattr(d, name, getattr(self, name).ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
        for name in self.DBL:
            setattr(d, name, getattr(self, name).ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
        d.row_mask = self.row_mask.ctypes.data_as(ctypes.POINTER(ctypes.c_ulonglong))
        return d


class DevicePlan(object):
    def __init__(self, compiled_model, tables):
        self.model = compiled_model
        self.t = tables
        L = _lib.lib()
        h = ctypes.c_void_p()
        d = tables.desc()
        mh = compiled_model.handle if compiled_model is not None else None
        _lib.check(L.sbm_plan_create(mh, ctypes.byref(d), ctypes.byref(h)))
        self._h = h
        self._finalizer = weakref.finalize(self, L.sbm_plan_destroy, h)

    @property
    def handle(self):
        return self._h

    @staticmethod
    def opts(solver='dp5', rtol=1e-10, atol=1e-10, max_steps=0, h0=0.0, jac_raw=False, dense_output=0):
        o = SolverOpts()
        o.solver = SOLVERS[solver] if isinstance(solver, str) else int(solver)
        o.max_steps, o.jac_raw, o.dense_output = int(max_steps), int(bool(jac_raw)), int(dense_output)
        o.rtol, o.atol, o.h0 = float(rtol), float(atol), float(h0)
        return o

    # ---- host-buffer entry points -------------------------------------------------------------
    def residuals_host(self, theta, opts, loss='square', want_sims=True, stream=0):
        t = self.t
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        B = theta.shape[0]
        if theta.shape[1] != t.n_proj:
            raise ValueError('theta has %d columns, project has %d parameters' % (theta.shape[1], t.n_proj))
        out = {'resid': np.empty((B, t.n_rows)), 'ssr': np.empty((B,)), 'sf': np.empty((B, max(t.n_groups, 1))),
               'sims': np.empty((B, t.n_rows)), 'status': np.empty((B, t.E), dtype=np.int32),
               'counters': np.zeros(_lib.SBM_N_COUNTERS, dtype=np.uint64)}
        _lib.check(_lib.lib().sbm_residuals_host(
            self._h, ctypes.byref(opts), LOSSES[loss], B, _ptr(theta), _ptr(out['resid']), _ptr(out['ssr']),
            _ptr(out['sf']), _ptr(out['sims']), _ptr(out['status']), _ptr(out['counters']), ctypes.c_void_p(stream)))
        out['sf'] = out['sf'][:, :t.n_groups]
        return out

    def jacobian_host(self, theta, opts, loss='square', flags=0, stream=0):
        t = self.t
        theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
        B = theta.shape[0]
        if theta.shape[1] != t.n_proj:
            raise ValueError('theta has %d columns, project has %d parameters' % (theta.shape[1], t.n_proj))
        out = {'jac': np.empty((B, t.n_rows, t.n_proj)), 'model_jac': np.empty((B, t.n_rows, t.n_proj)),
               'resid': np.empty((B, t.n_rows)), 'ssr': np.empty((B,)), 'grad': np.empty((B, t.n_proj)),
               'sf': np.empty((B, max(t.n_groups, 1))), 'sims': np.empty((B, t.n_rows)),
               'status': np.empty((B, t.E), dtype=np.int32),
               'counters': np.zeros(_lib.SBM_N_COUNTERS, dtype=np.uint64)}
        _lib.check(_lib.lib().sbm_jacobian_host(
            self._h, ctypes.byref(opts), LOSSES[loss], int(flags), B, _ptr(theta), _ptr(out['jac']),
            _ptr(out['model_jac']), _ptr(out['resid']), _ptr(out['ssr']), _ptr(out['grad']), _ptr(out['sf']),
            _ptr(out['sims']), _ptr(out['status']), _ptr(out['counters']), ctypes.c_void_p(stream)))
        out['sf'] = out['sf'][:, :t.n_groups]
        return out

    # ---- device-pointer entry points (torch tensors are only the memory hand-off) ----------------
    def solve_device(self, theta, opts, sims, jac, status, counters, theta_sb=None, theta_sp=None, stream=None):
        """Launch the integrator on device tensors.  ``theta`` is [B][n_proj] unless strides are given."""
        import torch
        B = theta.shape[0]
        sb = theta.stride(0) if theta_sb is None else theta_sb
        sp = theta.stride(1) if theta_sp is None else theta_sp
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream
        _lib.check(_lib.lib().sbm_solve(
            self._h, ctypes.byref(opts), B, ctypes.c_void_p(theta.data_ptr()), sb, sp,
            ctypes.c_void_p(sims.data_ptr()), ctypes.c_void_p(jac.data_ptr() if jac is not None else None),
            ctypes.c_void_p(status.data_ptr()), ctypes.c_void_p(counters.data_ptr()), ctypes.c_void_p(st)))

    def loss_device(self, theta, sims, jac, resid, ssr, sf, sfgrad, grad, loss='square', flags=0, stream=None):
        import torch
        B = theta.shape[0]
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream

        def dp(x):
            return ctypes.c_void_p(x.data_ptr() if x is not None else None)
        _lib.check(_lib.lib().sbm_loss(
            self._h, LOSSES[loss], int(flags), B, dp(theta), theta.stride(0), theta.stride(1),
            dp(sims), dp(jac), dp(resid), dp(ssr), dp(sf), dp(sfgrad), dp(grad), ctypes.c_void_p(st)))


    def loss_sums_device(self, theta, sims, jac, sums, loss='square', flags=0, stream=None):
        """First half of the loss epilogue (C ABI ``sbm_loss_sums``): the scale-factor group sums of THIS plan's rows into
        ``sums`` [B, n_groups, 3 + 2 n_proj]; the experiment shard all-reduces them over the ranks."""
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream

        def dp(x):
            return ctypes.c_void_p(x.data_ptr() if x is not None else None)
        _lib.check(_lib.lib().sbm_loss_sums(
            self._h, LOSSES[loss], int(flags), theta.shape[0], dp(theta), theta.stride(0), theta.stride(1),
            dp(sims), dp(jac), dp(sums), ctypes.c_void_p(st)))

    def loss_finish_device(self, theta, sims, jac, sums, resid, ssr, sf, sfgrad, grad, loss='square', flags=0, stream=None):
        """Second half (C ABI ``sbm_loss_finish``): scale factors from the given (reduced) sums, residuals, SSR, scaled
        Jacobian and J^T r of this plan's rows."""
        import torch
        st = torch.cuda.current_stream().cuda_stream if stream is None else stream

        def dp(x):
            return ctypes.c_void_p(x.data_ptr() if x is not None else None)
        _lib.check(_lib.lib().sbm_loss_finish(
            self._h, LOSSES[loss], int(flags), theta.shape[0], dp(theta), theta.stride(0), theta.stride(1),
            dp(sims), dp(jac), dp(sums), dp(resid), dp(ssr), dp(sf), dp(sfgrad), dp(grad), ctypes.c_void_p(st)))


def normal_equations_device(jac, resid, out=None, stream=None):
    """Per-set normal equations on device tensors (C ABI ``sbm_normal_equations``): jac [B, n_rows, n_proj], resid
    [B, n_rows] -> out [B, n_proj (n_proj + 1) / 2 + n_proj + 1] = lower triangle of J^T J (row-major) | J^T r | r^T r."""
    import torch
    B, R, P = jac.shape
    n = P * (P + 1) // 2 + P + 1
    if out is None:
        out = torch.empty((B, n), dtype=torch.float64, device=jac.device)
    if not (jac.is_contiguous() and resid.is_contiguous() and out.is_contiguous()):
        raise ValueError('normal_equations_device needs contiguous tensors')
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    vp = ctypes.c_void_p
    _lib.check(_lib.lib().sbm_normal_equations(B, R, P, vp(jac.data_ptr()), vp(resid.data_ptr()), vp(out.data_ptr()), vp(st)))
    return out


def unpack_normal_equations(packed, n_proj):
    """(J^T J [B, P, P] symmetric, J^T r [B, P], r^T r [B]) from the packed rows of ``normal_equations_device``."""
    import torch
    P = n_proj
    T = P * (P + 1) // 2
    B = packed.shape[0]
    A = torch.zeros((B, P, P), dtype=packed.dtype, device=packed.device)
    ii, jj = torch.tril_indices(P, P, device=packed.device)
    A[:, ii, jj] = packed[:, :T]
    A[:, jj, ii] = packed[:, :T]
    return A, packed[:, T:T + P], packed[:, T + P]


def lm_step_device(jac, resid, lam, dscale=None, delta=None, pred=None, gmax=None, info=None, stream=None):
    """Batched Levenberg-Marquardt step on device tensors (C ABI ``sbm_lm_step``, csrc/sbm_lm.cuh):
    ``(J^T J + lam D) delta = -J^T r`` for every parameter set, ``D = max(dscale, diag(J^T J))`` written back to
    ``dscale`` [B, n_proj] (zero before the first step; None: this step's diagonal only).  jac [B, n_rows, n_proj],
    resid [B, n_rows], lam [B] (CUDA float64, contiguous).  Returns (delta [B, n_proj], pred [B], gmax [B], info [B] int32)."""
    import torch
    B, R, P = jac.shape
    f64 = dict(dtype=torch.float64, device=jac.device)
    delta = torch.empty((B, P), **f64) if delta is None else delta
    pred = torch.empty((B,), **f64) if pred is None else pred
    gmax = torch.empty((B,), **f64) if gmax is None else gmax
    info = torch.empty((B,), dtype=torch.int32, device=jac.device) if info is None else info
    for x in (jac, resid, lam, delta, pred, gmax, info) + ((dscale,) if dscale is not None else ()):
        if not x.is_contiguous():
            raise ValueError('lm_step_device needs contiguous tensors')
    st = torch.cuda.current_stream().cuda_stream if stream is None else stream
    vp = ctypes.c_void_p
    _lib.check(_lib.lib().sbm_lm_step(B, R, P, vp(jac.data_ptr()), vp(resid.data_ptr()), vp(lam.data_ptr()),
                                      vp(dscale.data_ptr() if dscale is not None else None),
                                      vp(delta.data_ptr()), vp(pred.data_ptr()), vp(gmax.data_ptr()),
                                      vp(info.data_ptr()), vp(st)))
    return delta, pred, gmax, info

