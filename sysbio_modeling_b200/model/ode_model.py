"""``OdeModel`` with the B200 backend - drop-in for reference ``model/ode_model.py:7-169``.

The reference integrates one experiment at a time with SciPy's LSODA through a Python callback.  Here
``simulate`` / ``calc_jacobian`` keep their signatures and return shapes but run the NVRTC-compiled
Dormand-Prince / Rosenbrock kernels (``csrc/kernels/sbm_ode_kernels.cuh``); ``*_batch`` variants take
many parameter vectors at once.  The constructor still accepts the reference's Python RHS callables
``model(y, t, yout, p)``: they are traced symbolically (``symbolic.trace_rhs_callable``) to obtain the
expressions the CUDA emitter needs, unless a ``SymbolicModel`` is passed in.

There is no CPU integration path in this class: without the native library and a CUDA device the
compute methods raise ``SbmError``.
"""
import numpy as np

from .abstract_model import ModelABC
from ..symbolic import (SymbolicModel, symbolic_model_from_file, trace_rhs_callable, make_ode_model)
from .. import engine
from .._lib import SbmError, STATUS_NAMES


class OdeModel(ModelABC):
    """Differential-equation model integrated on the GPU.

    Parameters (reference ``ode_model.py:27-29`` first, additive keyword-only ones after ``*``)
    ----------
    model, sens_model : callables ``f(y, t, yout, p)`` (may be None when ``symbolic`` is given)
    n_vars : int
    param_order : list of str
    model_name, use_jit, model_jac, sens_model_jac, jit_type : kept for signature compatibility; the
        CUDA backend always compiles (NVRTC) and derives the analytic Jacobian itself.
    symbolic : SymbolicModel, optional - skip tracing.
    fixed_params : parameters without sensitivity columns (reference ``sympy_tools.py:137-139``).
    rtol, atol : tolerances; defaults are the reference's hard-coded 1e-10 (``ode_model.py:122-123``).
    solver : 'dp5' (non-stiff) or 'rodas4' (stiff; with sensitivities for autonomous models of at most 12 states).
    """

    def __init__(self, model, sens_model, n_vars, param_order, model_name='Model', use_jit=True,
                 model_jac=None, sens_model_jac=None, jit_type='numba', *, symbolic=None, fixed_params=None,
                 rtol=1e-10, atol=1e-10, solver='dp5', max_steps=0, group_lanes=0, cols_per_lane=0, sens_layout=0):
        super(OdeModel, self).__init__(model, n_vars, param_order, model_name)
        self.sens_model = sens_model
        self.model_jac = model_jac
        self.sens_model_jac = sens_model_jac
        self.use_jac = True
        self.jit_type = jit_type
        self.rtol, self.atol, self.solver, self.max_steps = rtol, atol, solver, max_steps
        # states-only Dormand-Prince on a dense time grid: 0 = automatic (continuous extension from 256 output times per
        # trajectory on: the reference's 1000-node grids), 1 = always, 2 = every step lands on its output time
        self.dense_output = 0
        if symbolic is None:
            if model is None:
                raise ValueError('either a model callable or a SymbolicModel is required')
            symbolic = trace_rhs_callable(model, n_vars, param_order, name=model_name, fixed_params=fixed_params)
        elif fixed_params is not None:
            symbolic = symbolic.with_fixed(fixed_params)
        if symbolic.n_vars != n_vars or list(symbolic.param_order) != list(param_order):
            raise ValueError('symbolic model does not match n_vars / param_order')
        self.symbolic = symbolic
        self._group = (group_lanes, cols_per_lane, sens_layout)
        self._compiled = None
        self._jit_enabled = bool(use_jit)

    # ---- construction helpers ---------------------------------------------------------------------
    @classmethod
    def from_model_file(cls, model_file, model_name=None, fixed_params=None, **kw):
        """Build from a ``#*!`` model file (the reference's symbolic front end, sympy_tools.py:272-397)."""
        sm = symbolic_model_from_file(model_file, name=model_name or 'Model', fixed_params=fixed_params)
        return cls(None, None, sm.n_vars, sm.param_order, model_name=sm.name, symbolic=sm, **kw)

    def python_callables(self):
        """(model, sens_model) in the reference ABI, generated from the same expressions (for oracles)."""
        return make_ode_model(self.symbolic, False), make_ode_model(self.symbolic, True)

    def enable_jit(self):
        self._jit_enabled = True

    def disable_jit(self):
        # kept for API compatibility: the GPU backend has no interpreted mode
        self._jit_enabled = False

    @property
    def compiled(self):
        if self._compiled is None:
            self._compiled = engine.compiled_model_for(self.symbolic, group_lanes=self._group[0],
                                                       cols_per_lane=self._group[1], sens_layout=self._group[2])
        return self._compiled

    @property
    def n_free(self):
        return self.symbolic.n_free

    # ---- integration ------------------------------------------------------------------------------
    def _grid_plan(self, params, t_sim, y0, with_sens):
        """A plan with one experiment per parameter vector and one 'direct' row per (time, state)."""
        P = np.ascontiguousarray(np.atleast_2d(params), dtype=np.float64)
        nb, NP, N = P.shape[0], len(self.param_order), self._n_vars
        if P.shape[1] != NP:
            raise ValueError('expected %d parameters, got %d' % (NP, P.shape[1]))
        t_sim = np.asarray(t_sim, dtype=np.float64)
        if np.any(np.diff(t_sim) <= 0):
            raise ValueError('t_sim must be strictly increasing')
        # the reference hands t_sim straight to odeint, which takes t_sim[0] as the INITIAL time: y(t_sim[0]) = init_conditions
        # (ode_model.py:122-123, 167-168).  The kernels integrate from 0, so the grid is shifted; that is exact for an
        # autonomous right-hand side (every model the symbolic front end has produced so far) and refused otherwise.
        if t_sim[0] != 0.0:
            if not self.symbolic.autonomous:
                raise NotImplementedError('t_sim[0] != 0 with a time-dependent right-hand side is not supported by the CUDA backend')
            t_sim = t_sim - t_sim[0]
        tpos = t_sim[t_sim > 0]
        T = len(tpos)
        y0 = np.broadcast_to(np.asarray(y0, dtype=np.float64), (nb, N))
        rows_per = T * N
        tout_off = np.arange(nb + 1, dtype=np.int32) * T
        tout = np.tile(tpos, nb)
        orow_off = np.arange(nb * T + 1, dtype=np.int32) * N
        base = (np.arange(nb)[:, None, None] * rows_per + np.arange(T)[None, :, None] * N + np.arange(N)[None, None, :])
        row_sim = base.reshape(-1)
        row_mask = np.tile((np.uint64(1) << np.arange(N, dtype=np.uint64)), nb * T)
        n_rows = nb * rows_per
        tables = engine.PlanTables(
            E=nb, n_rows=n_rows, n_proj=self.n_free if with_sens else 0, n_groups=0,
            pmap=np.full((nb, NP), -1), pfix=P, y0=y0, tout_off=tout_off, tout=tout, orow_off=orow_off,
            row_sim=row_sim, row_jac=row_sim, row_mask=row_mask, data=np.zeros(n_rows), sigma=np.ones(n_rows),
            row_group=np.full(n_rows, -1), row_kind=np.zeros(n_rows), row_aux=np.zeros(n_rows))
        return engine.DevicePlan(self.compiled, tables), T, t_sim

    def _run(self, params, t_sim, y0, with_sens):
        import torch
        from .. import _lib
        _lib.require_gpu()
        plan, T, t_sim = self._grid_plan(params, t_sim, y0, with_sens)
        nb, N, NF = plan.t.E, self._n_vars, self.n_free
        dev = torch.device('cuda')
        theta = torch.zeros((1, max(plan.t.n_proj, 1)), dtype=torch.float64, device=dev)
        sims = torch.zeros((1, plan.t.n_rows), dtype=torch.float64, device=dev)
        jac = torch.zeros((1, plan.t.n_rows, NF), dtype=torch.float64, device=dev) if with_sens else None
        status = torch.zeros((1, nb), dtype=torch.int32, device=dev)
        counters = torch.zeros(_lib.SBM_N_COUNTERS, dtype=torch.int64, device=dev)
        # sensitivities: Dormand-Prince, or RODAS4 with forward sensitivities when the model was declared stiff
        solver = self.solver if (not with_sens or str(self.solver).lower() in ('rodas4', 'rodas')) else 'dp5'
        opts = plan.opts(solver, self.rtol, self.atol, self.max_steps, jac_raw=True, dense_output=self.dense_output)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        plan.solve_device(theta, opts, sims, jac, status, counters)
        ev1.record()
        torch.cuda.synchronize()
        self.last_solve_ms = ev0.elapsed_time(ev1)       # device time of the integrator launch (CUDA events)
        st = status.cpu().numpy().reshape(-1)
        self.last_status = st
        self.last_counters = dict(zip(_lib.COUNTER_NAMES, counters.cpu().numpy().tolist()))
        y = sims.cpu().numpy().reshape(nb, T, N)
        S = jac.cpu().numpy().reshape(nb, T, N, NF) if with_sens else None
        # t = 0 rows come straight from the initial conditions
        has0 = len(t_sim) > T
        if has0:
            y0b = np.broadcast_to(np.asarray(y0, dtype=np.float64), (nb, N))
            y = np.concatenate([y0b[:, None, :], y], axis=1)
            if with_sens:
                S = np.concatenate([np.zeros((nb, 1, N, NF)), S], axis=1)
        return y, S

    def simulate(self, experiment_params, t_sim, init_conditions=None):
        """Reference ``ode_model.py:128-169``: returns f8[len(t_sim), n_vars]."""
        y0 = np.zeros((self._n_vars,)) if init_conditions is None else np.asarray(init_conditions)[:self._n_vars]
        y, _ = self._run(np.asarray(experiment_params)[None, :], t_sim, y0, False)
        return y[0]

    def calc_jacobian(self, experiment_params, t_sim, init_conditions):
        """Reference ``ode_model.py:83-126``: returns f8[len(t_sim), n_vars*n_free], column i*n_free+j =
        d x_i / d p_j.  Sensitivities start from zero (the reference always passes zeros, base_project.py:512);
        only the first n_vars entries of ``init_conditions`` (the state) are used."""
        init_conditions = np.asarray(init_conditions, dtype=np.float64)
        y0 = init_conditions[:self._n_vars]
        if init_conditions.size > self._n_vars and np.any(init_conditions[self._n_vars:] != 0):
            raise NotImplementedError('non-zero initial sensitivities are not supported by the CUDA backend')
        _, S = self._run(np.asarray(experiment_params)[None, :], t_sim, y0, True)
        return S[0].reshape(S.shape[1], -1)

    # ---- additive batch API -----------------------------------------------------------------------
    def simulate_batch(self, params, t_sim, init_conditions=None):
        """params f8[B, n_params] -> f8[B, len(t_sim), n_vars]."""
        y0 = np.zeros((self._n_vars,)) if init_conditions is None else np.asarray(init_conditions)
        return self._run(params, t_sim, y0, False)[0]

    def calc_jacobian_batch(self, params, t_sim, init_conditions=None):
        """params f8[B, n_params] -> (y f8[B,T,N], S f8[B,T,N,n_free])."""
        y0 = np.zeros((self._n_vars,)) if init_conditions is None else np.asarray(init_conditions)
        return self._run(params, t_sim, y0, True)
