"""Loss-function front ends (reference ``project/loss_functions/abstract_loss_function.py:9-147``).

The classes keep the reference's DataFrame-facing API (``residuals``, ``jacobian``,
``update_scale_factors``, ``scale_sim_values``...) but every number is produced by the CUDA loss
epilogue ``sbm_loss`` (``csrc/sbm_epilogue.cuh``): the frames are flattened into a loss-only device
plan (row tables + scale-factor groups), the kernel runs, and the results are wrapped back into
pandas objects.  ``Project`` bypasses the frames and calls the same kernel on its own plan.
"""
from copy import deepcopy

import numpy as np

from ..utils import OrderedHashDict
from ... import _lib, engine

PRIOR_LABEL = '~Prior'
SF_PRIOR_LABEL = '~~SF_Prior'


class LossFunctionABC(object):
    def __init__(self):
        pass

    def evaluate(self, simulations, experiment_measures):
        pass

    def residuals(self, simulations, experiment_measures):
        pass


class DifferentiableLossFunctionABC(LossFunctionABC):
    def gradient(self, sim_values, exp_values, exp_std):
        pass

    def jacobian(self, sim_values, exp_values, exp_std):
        pass


class LossFunctionWithScaleFactors(LossFunctionABC):
    kind = 'square'

    def __init__(self, sf_groups, SF):
        super(LossFunctionWithScaleFactors, self).__init__()
        self._scale_factors = OrderedHashDict()
        if sf_groups is not None:
            if isinstance(sf_groups, (str, frozenset)):
                sf_groups = [sf_groups]
            for measure_group in sf_groups:
                self._scale_factors[measure_group] = SF()

    # ---- reference accessors ----------------------------------------------------------------------
    @property
    def scale_factors(self):
        return deepcopy(self._scale_factors)

    def set_scale_factor_priors(self, measure_name, log_scale_factor_prior, log_sigma_scale_factor):
        try:
            self._scale_factors[measure_name].log_prior = log_scale_factor_prior
            self._scale_factors[measure_name].log_prior_sigma = log_sigma_scale_factor
        except KeyError:
            raise KeyError("%s not present as a scale factor" % (measure_name,))

    # ---- group bookkeeping shared with Project ----------------------------------------------------
    def group_keys(self):
        return list(self._scale_factors.keys())

    def group_members(self, key):
        return [key] if isinstance(key, str) else list(key)

    def group_index(self, measure_name):
        for g, key in enumerate(self._scale_factors.keys()):
            if measure_name == key or (not isinstance(key, str) and measure_name in key):
                return g
        return -1

    def store_results(self, sf_values, sf_gradients=None):
        """Record the scale factors (and gradients) of the last evaluation in the holder objects."""
        for g, key in enumerate(self._scale_factors.keys()):
            holder = OrderedHashDictGet(self._scale_factors, key)
            holder._sf = float(sf_values[g])
            if sf_gradients is not None:
                holder._sf_gradient = np.array(sf_gradients[g], dtype=np.float64)

    # ---- frames -> loss-only device plan ----------------------------------------------------------
    def _frame_tables(self, simulations, experiment_measures, n_proj=0):
        labels = list(simulations.index)
        n = len(labels)
        data = np.asarray(experiment_measures['mean'].values, dtype=np.float64).copy()
        sigma = np.asarray(experiment_measures['std'].values, dtype=np.float64).copy()
        kind = np.zeros(n, dtype=np.int32)
        group = np.full(n, -1, dtype=np.int32)
        aux = np.zeros(n, dtype=np.int32)
        for r, lab in enumerate(labels):
            l0, l1 = lab[0], lab[1]
            if l0 == SF_PRIOR_LABEL:
                kind[r] = _lib.SBM_ROW_SF_PRIOR
                aux[r] = self.group_index(l1[1:])
            elif l0 == PRIOR_LABEL:
                kind[r] = _lib.SBM_ROW_PLAIN
            else:
                group[r] = self.group_index(l1)
        if self.kind == 'normalized':
            meas = kind == _lib.SBM_ROW_MEASUREMENT
            sigma[meas] *= data[meas]           # normalized_squared_loss_function.py:38-46
        if self.kind == 'log' and np.any(data[kind == _lib.SBM_ROW_MEASUREMENT] <= 0):
            raise ValueError("LogSquare loss cannot handle measurements smaller or equal to zero")
        off, rows = [0], []
        for key in self._scale_factors.keys():
            for name in self.group_members(key):      # measure by measure (reference :72-76)
                rows += [r for r, lab in enumerate(labels) if lab[0] not in (PRIOR_LABEL, SF_PRIOR_LABEL)
                         and lab[1] == name]
            off.append(len(rows))
        # a scale factor with a prior needs its "~~SF_Prior" row (reference raises KeyError, :104-105)
        for g, key in enumerate(self._scale_factors.keys()):
            sf = OrderedHashDictGet(self._scale_factors, key)
            if sf.log_prior is not None and not np.any((kind == _lib.SBM_ROW_SF_PRIOR) & (aux == g)):
                raise KeyError("No prior in simulations for %s" % (key,))
        return engine.PlanTables(E=0, n_rows=n, n_proj=n_proj, n_groups=len(self._scale_factors),
                                 data=data, sigma=sigma, row_group=group, row_kind=kind, row_aux=aux,
                                 grp_off=np.array(off), grp_rows=np.array(rows))

    def _run_frames(self, simulations, experiment_measures, simulations_jacobian=None):
        import torch
        _lib.require_gpu()
        n_proj = 0 if simulations_jacobian is None else simulations_jacobian.shape[1]
        tables = self._frame_tables(simulations, experiment_measures, n_proj)
        plan = engine.DevicePlan(None, tables)
        dev = torch.device('cuda')
        f64 = dict(dtype=torch.float64, device=dev)
        n, G = tables.n_rows, tables.n_groups
        sims = torch.from_numpy(np.ascontiguousarray(simulations['mean'].values, dtype=np.float64)[None, :].copy()).to(dev)
        theta = torch.zeros((1, max(n_proj, 1)), **f64)
        jac = None
        if simulations_jacobian is not None:
            # C-contiguous [1][n_rows][n_proj]: DataFrame.values is often column-major and torch keeps strides
            jac = torch.from_numpy(np.ascontiguousarray(simulations_jacobian.values, dtype=np.float64)[None, :, :]
                                   .copy()).to(dev).contiguous()
        resid = torch.zeros((1, n), **f64)
        ssr = torch.zeros((1,), **f64)
        sf = torch.zeros((1, max(G, 1)), **f64)
        sfgrad = torch.zeros((1, max(G, 1), max(n_proj, 1)), **f64) if jac is not None else None
        grad = torch.zeros((1, max(n_proj, 1)), **f64) if jac is not None else None
        kind = 'square' if self.kind == 'normalized' else self.kind   # sigma already rescaled above
        plan.loss_device(theta, sims, jac, resid, ssr, sf, sfgrad, grad, loss=kind)
        torch.cuda.synchronize()
        out = {'resid': resid.cpu().numpy()[0], 'ssr': float(ssr.cpu().numpy()[0]),
               'sf': sf.cpu().numpy()[0][:G], 'sims': sims.cpu().numpy()[0]}
        if jac is not None:
            out['jac'] = jac.cpu().numpy()[0]
            out['sfgrad'] = sfgrad.cpu().numpy()[0][:G]
            out['grad'] = grad.cpu().numpy()[0]
        self.store_results(out['sf'], out.get('sfgrad'))
        return out

    # ---- reference DataFrame API ------------------------------------------------------------------
    def update_scale_factors(self, simulations, experiment_measures):
        """Reference :87-91."""
        self._run_frames(simulations, experiment_measures)

    def scale_sim_values(self, simulations):
        """Reference :46-59: multiply each measure's rows by its scale factor (host-side view helper)."""
        scaled = simulations.copy()
        vals = np.array(scaled['mean'].values, dtype=np.float64, copy=True)
        for r, lab in enumerate(scaled.index):
            if lab[0] in (PRIOR_LABEL, SF_PRIOR_LABEL):
                continue
            if lab[1] in self._scale_factors:
                vals[r] *= self._scale_factors[lab[1]].sf
        scaled['mean'] = vals
        return scaled

    def total_sf_entropy(self, simulations, experiment_measures, t=1.0):
        """Reference :122-134."""
        sims = np.asarray(simulations['mean'].values, dtype=np.float64)
        tables = self._frame_tables(simulations, experiment_measures)
        entropy = 0.0
        for g, key in enumerate(self._scale_factors.keys()):
            rows = tables.grp_rows[tables.grp_off[g]:tables.grp_off[g + 1]]
            s, d, sg = sims[rows], tables.data[rows], tables.sigma[rows]
            sds = np.sum(s * s / sg ** 2)
            sde = np.sum(s * d / sg ** 2)
            holder = OrderedHashDictGet(self._scale_factors, key)
            entropy += t * holder.calc_scale_factor_entropy(sds, sde, t)
        return entropy


def OrderedHashDictGet(d, key):
    """Fetch by exact key (str or frozenset) without the member-alias lookup."""
    from collections import OrderedDict
    return OrderedDict.__getitem__(d, key)
