"""``Project`` - the objective assembled from a model, experiments and a loss function, evaluated on B200.

Drop-in for reference ``project/base_project.py`` (class ``Project``): same constructor, same public
methods and return types (``residuals`` -> Series, ``calc_project_jacobian`` -> DataFrame, ...), same
row and column conventions.  What changes is where the numbers come from:

* the parameter-index bookkeeping (``_set_local_param_idx``, reference :164-276) stays host Python;
* everything the reference recomputes per call in Python/pandas - parameter gather, ``linspace`` /
  ``searchsorted`` time snapping, mapping to measurements, chain rule, column scatter, scale
  factors, residuals, scaled Jacobian - is compiled ONCE into a flat numeric plan (``PlanTables``)
  and evaluated by the CUDA integrators + loss epilogue for any number of parameter vectors.

Additive API: ``residuals_batch``, ``ssr_batch``, ``jacobian_batch``, ``rss_gradient_batch`` (numpy in,
numpy out) and ``evaluate_device`` (torch tensors resident in HBM), plus solver options.

Reference defects and how they are handled:
* Jacobian row bookkeeping for experiments with several measurements (reference :467, :488, :516):
  default here is the intended one row block per measurement; ``reference_jac_rows=True`` reproduces
  the reference's rows exactly.
* parameter-prior Jacobian rows stay zero in the reference (``_update_parameter_prior_jacobian`` is
  never called); default here matches the reference, ``prior_jacobian_unit=True`` gives unit rows.
* default-Global parameters are ordered by ``param_order`` (the reference iterates a set).
* 'custom' (arbitrary Python) mappings cannot run inside a kernel and raise NotImplementedError.
"""
import copy
import warnings
from collections import OrderedDict, defaultdict

import numpy as np
import pandas as pd

from .loss_functions.squared_loss import SquareLossFunction
from .loss_functions.abstract_loss_function import PRIOR_LABEL, SF_PRIOR_LABEL
from .. import _lib, engine

N_GRID = 1000        # reference base_project.py:419


def snapped_times(t_end, t_meas):
    """Measurement times snapped UP to the reference's 1000-node grid (base_project.py:418-420,
    utils.py:18-21).  Bit-identical to the reference because it is the same numpy expression."""
    t_sim = np.linspace(0, t_end, N_GRID)
    idx = np.searchsorted(t_sim, t_meas)
    return t_sim[idx]


class Project(object):
    def __init__(self, model, experiments, model_parameter_settings, measurement_to_model_map,
                 sf_groups=None, loss_function=SquareLossFunction, *, solver='dp5', rtol=None, atol=None,
                 max_steps=0, reference_jac_rows=False, prior_jacobian_unit=False):
        self.project_description = ""
        self._model = model
        self._model_parameter_settings = model_parameter_settings
        if hasattr(loss_function, 'scale_factors'):
            self._loss_function = loss_function(sf_groups)
        elif sf_groups is not None:
            raise ValueError("Loss Function %s does not support scale factors" % type(loss_function))
        else:
            self._loss_function = loss_function
        # solver options (reference hard-codes rtol = atol = 1e-10, ode_model.py:122-123).  solver: 'dp5' (explicit), 'rodas4'
        # (stiff: states, and sensitivities for autonomous models of at most 12 states) or 'auto' (_concrete_solver)
        self.solver = solver
        self.rtol = rtol if rtol is not None else getattr(model, 'rtol', 1e-10)
        self.atol = atol if atol is not None else getattr(model, 'atol', 1e-10)
        self.max_steps = max_steps
        self.reference_jac_rows = bool(reference_jac_rows)
        self._auto_choice = {}                 # solver='auto': decision per path (False: states only, True: sensitivities)
        self.auto_solver_report = {}
        self.prior_jacobian_unit = bool(prior_jacobian_unit)

        self._parameter_priors = OrderedDict()
        self._scale_factor_priors = []
        self._measurement_to_model_map = {}
        self._make_mapping(measurement_to_model_map)

        self._experiments = []
        self._project_param_idx = None
        self._n_project_params = None
        self._residuals_per_param = None
        self._n_residuals = None
        self._plan_tables = None
        self._device_plan = None
        self._last = None
        self.add_experiment(experiments)
        self._project_param_vector = np.zeros((self.n_project_params,))

    # ------------------------------------------------------------------------------------------
    # mapping (reference :109-136)
    # ------------------------------------------------------------------------------------------
    def _make_mapping(self, measurement_to_model_map):
        n_vars = self._model.n_vars
        for measure_name, (mapping_type, mapping_args) in measurement_to_model_map.items():
            if mapping_type == 'direct':
                assert type(mapping_args) is int
                if mapping_args >= n_vars:
                    raise ValueError('Index (%d) has to be smaller than %d' % (mapping_args, n_vars))
                mask = 1 << mapping_args
            elif mapping_type == 'sum':
                assert type(mapping_args) is list
                mask = 0
                for v in mapping_args:
                    if v >= n_vars:
                        raise ValueError('Index (%d) has to be smaller than %d' % (v, n_vars))
                    if mask & (1 << v):
                        raise NotImplementedError("'sum' mapping with a repeated variable index")
                    mask |= (1 << v)
            elif mapping_type == 'custom':
                raise NotImplementedError("'custom' measurement mappings are arbitrary Python and cannot run "
                                          "inside the CUDA integrator; use 'direct' or 'sum'")
            else:
                raise ValueError('Invalid mapping type')
            self._measurement_to_model_map[measure_name] = {'type': mapping_type, 'parameters': mapping_args,
                                                            'mask': mask}

    # ------------------------------------------------------------------------------------------
    # parameter index bookkeeping (reference :164-276)
    # ------------------------------------------------------------------------------------------
    def _set_local_param_idx(self):
        settings = self._model_parameter_settings
        param_order = list(self._model.param_order)
        local_pars = list(settings.get('Local', []))
        project_fixed = list(settings.get('Fixed', []))
        global_pars = list(settings.get('Global', []))
        shared_groups = settings.get('Shared', {})
        shared_pars = set(p for g in shared_groups for p in shared_groups[g])
        named = set(local_pars) | set(project_fixed) | shared_pars | set(global_pars)
        unspecified = [p for p in param_order if p not in named]       # deterministic order
        global_pars = global_pars + unspecified

        idx = OrderedDict()
        per_param = defaultdict(lambda: defaultdict(lambda: 0))
        n = 0
        for p in global_pars:
            idx[p] = {'Global': n}
            n += 1
        for experiment in self._experiments:
            exp_idx = OrderedDict()
            exp_fixed = list(experiment.fixed_parameters.keys()) if experiment.fixed_parameters else []
            all_fixed = project_fixed + exp_fixed
            for p in project_fixed:
                if p not in exp_fixed:
                    raise ValueError('%s was declared as a fixed parameter, but in experiment %s no value provided'
                                     % (p, experiment.name))
            n_pts = len(experiment.get_unique_timepoints())
            for p in global_pars:
                if p not in all_fixed:
                    exp_idx[p] = idx[p]['Global']
                    per_param[p]['Global'] += n_pts
            for g in shared_groups:
                for p, dep in shared_groups[g].items():
                    if p in exp_fixed:
                        continue
                    key = 'None' if dep is None else tuple(experiment.settings[s] for s in tuple(dep))
                    idx.setdefault(g, {})
                    if key not in idx[g]:
                        idx[g][key] = n
                        n += 1
                    exp_idx[p] = idx[g][key]
                    per_param[g][key] += n_pts
            for p in local_pars:
                if p in exp_fixed:
                    continue
                name = '%s_%s' % (p, experiment.name)
                idx[name] = {'Local': n}
                exp_idx[p] = n
                n += 1
                per_param[name]['Local'] += n_pts
            experiment.param_global_vector_idx = exp_idx
        return idx, n, per_param

    def _update_n_residuals(self, include_zero=False):
        n = 0
        for experiment in self._experiments:
            for m in experiment.measurements:
                t = m.timepoints
                n += len(t) if include_zero else int(np.count_nonzero(t != 0))
        return n

    def _update_project_settings(self):
        self._project_param_idx, self._n_project_params, self._residuals_per_param = self._set_local_param_idx()
        self._n_residuals = self._update_n_residuals()
        self._build_plan_tables()
        self._device_plan = None
        self._last = None
        self._project_param_vector = np.zeros((self._n_project_params,))

    # ------------------------------------------------------------------------------------------
    # the flat plan
    # ------------------------------------------------------------------------------------------
    def get_ordered_project_params(self):
        """Column labels of the Jacobian: project parameters by index (reference :1064-1079)."""
        names = [None] * self._n_project_params
        for group, d in self._project_param_idx.items():
            for setting, i in d.items():
                names[i] = group + ' ' + ''.join(str(x) for x in setting)
        return names

    def _build_plan_tables(self):
        model = self._model
        NP, N = len(model.param_order), model.n_vars
        free = list(getattr(model, 'symbolic').free)
        lf = self._loss_function
        has_sf = hasattr(lf, 'group_keys')

        labels, data, sigma, meas_t = [], [], [], []
        row_group, row_kind, row_aux = [], [], []
        pmap, pfix, y0 = [], [], []
        tout_off, tout, orow_off, row_sim, row_jac, row_mask = [0], [], [0], [], [], []
        jac_atomic = 1 if self.reference_jac_rows else 0
        sim_cursor = 0
        jac_cursor = 0
        for experiment in self._experiments:
            ep = experiment.param_global_vector_idx
            # ---- parameter gather map (reference :343-363)
            pm, pf = [], []
            for name in model.param_order:
                if name in ep:
                    pm.append(ep[name])
                    pf.append(0.0)
                else:
                    if not experiment.fixed_parameters or name not in experiment.fixed_parameters:
                        raise KeyError('%s not in %s fixed parameters.' % (name, experiment.name))
                    pm.append(-1)
                    pf.append(float(experiment.fixed_parameters[name]))
            exp_free = [name for name in model.param_order if name in ep]
            if exp_free != free:
                raise ValueError('experiment %s optimises parameters %s but the model was built with sensitivity '
                                 'columns for %s (every experiment must share the model\'s fixed-parameter set)'
                                 % (experiment.name, exp_free, free))
            cols = [ep[name] for name in exp_free]
            if len(set(cols)) != len(cols):
                jac_atomic = 1          # two model parameters share one project column: accumulate
            pmap.append(pm)
            pfix.append(pf)
            ic = getattr(experiment, 'initial_conditions', None)
            y0.append(np.zeros(N) if ic is None or (isinstance(ic, dict) and not ic) else np.asarray(ic, dtype=np.float64))

            # ---- rows and snapped output times (reference :365-391, :418-420; utils.py:18-21)
            t_end = experiment.get_unique_timepoints()[-1]
            per_time = OrderedDict()
            exp_jac_start = jac_cursor
            last_len = 0
            for m in experiment.measurements:
                if m.variable_name not in self._measurement_to_model_map:
                    raise KeyError('no mapping for measurement %s' % m.variable_name)
                mask = self._measurement_to_model_map[m.variable_name]['mask']
                vals, stds, tms = m.get_nonzero_measurements()
                snapped = snapped_times(t_end, tms)
                g = lf.group_index(m.variable_name) if has_sf else -1
                jac_base = exp_jac_start if self.reference_jac_rows else jac_cursor
                for q in range(len(vals)):
                    labels.append((experiment.name, m.variable_name))
                    data.append(vals[q]), sigma.append(stds[q]), meas_t.append(snapped[q])
                    row_group.append(g), row_kind.append(_lib.SBM_ROW_MEASUREMENT), row_aux.append(0)
                    per_time.setdefault(float(snapped[q]), []).append((sim_cursor, jac_base + q, mask))
                    sim_cursor += 1
                last_len = len(vals)
                if not self.reference_jac_rows:
                    jac_cursor += last_len
            if self.reference_jac_rows:
                jac_cursor = exp_jac_start + last_len           # reference :488, :516
            for tk in sorted(per_time.keys()):
                tout.append(tk)
                for (rs, rj, mk) in per_time[tk]:
                    row_sim.append(rs), row_jac.append(rj), row_mask.append(mk)
                orow_off.append(len(row_sim))
            tout_off.append(len(tout))
        n_meas_rows = sim_cursor

        # ---- parameter-prior rows: data sorted by label (measurement frame is lexsorted, :340) but the
        #      simulated value (theta) is written in insertion order (:427-437)
        prior_items = list(self._iter_parameter_priors())
        prior_labels = sorted((PRIOR_LABEL, g + ' ' + ''.join(str(s) for s in st)) for g, st, _, _ in prior_items)
        by_label = {(PRIOR_LABEL, g + ' ' + ''.join(str(s) for s in st)): (mu, sg) for g, st, mu, sg in prior_items}
        for k, lab in enumerate(prior_labels):
            mu, sg = by_label[lab]
            g_ins, st_ins = prior_items[k][0], prior_items[k][1]
            labels.append(lab), data.append(mu), sigma.append(sg), meas_t.append(np.nan)
            row_group.append(-1), row_kind.append(_lib.SBM_ROW_PARAM_PRIOR)
            row_aux.append(self._project_param_idx[g_ins][st_ins])
        # ---- scale-factor prior rows (:327-335)
        sfp = []
        for measure in self._scale_factor_priors:
            sf = lf.scale_factors[measure]
            name = measure if isinstance(measure, str) else next(iter(measure))
            sfp.append(((SF_PRIOR_LABEL, '~%s' % name), sf.log_prior, sf.log_prior_sigma, lf.group_index(name)))
        for lab, mu, sg, g in sorted(sfp, key=lambda r: r[0]):
            labels.append(lab), data.append(mu), sigma.append(sg), meas_t.append(np.nan)
            row_group.append(-1), row_kind.append(_lib.SBM_ROW_SF_PRIOR), row_aux.append(g)

        n_rows = len(labels)
        data = np.array(data, dtype=np.float64)
        sigma = np.array(sigma, dtype=np.float64)
        kind = np.array(row_kind, dtype=np.int32)
        loss_kind = getattr(lf, 'kind', 'square')
        sigma_eff = sigma.copy()
        if loss_kind == 'normalized':
            sigma_eff[kind == _lib.SBM_ROW_MEASUREMENT] *= data[kind == _lib.SBM_ROW_MEASUREMENT]
        if loss_kind == 'log' and np.any(data[kind == _lib.SBM_ROW_MEASUREMENT] <= 0):
            raise ValueError("LogSquare loss cannot handle measurements smaller or equal to zero")

        grp_off, grp_rows = [0], []
        n_groups = 0
        if has_sf:
            for key in lf.group_keys():
                for name in lf.group_members(key):
                    grp_rows += [r for r in range(n_meas_rows) if labels[r][1] == name]
                grp_off.append(len(grp_rows))
                n_groups += 1

        self._labels = labels
        self._index = pd.MultiIndex.from_tuples(labels) if labels else pd.MultiIndex.from_arrays([[], []])
        self._n_meas_rows = n_meas_rows
        self._meas_sigma_raw = sigma
        self._meas_t = np.array(meas_t, dtype=np.float64)
        self._loss_kind = loss_kind
        self._plan_tables = engine.PlanTables(
            E=len(self._experiments), n_rows=n_rows, n_proj=self._n_project_params, n_groups=n_groups,
            jac_atomic=jac_atomic, pmap=np.array(pmap).reshape(-1, NP) if pmap else np.zeros((0, NP)),
            pfix=np.array(pfix).reshape(-1, NP) if pfix else np.zeros((0, NP)),
            y0=np.array(y0).reshape(-1, N) if y0 else np.zeros((0, N)),
            tout_off=tout_off, tout=tout, orow_off=orow_off, row_sim=row_sim, row_jac=row_jac,
            row_mask=np.array(row_mask, dtype=np.uint64), data=data, sigma=sigma_eff,
            row_group=row_group, row_kind=kind, row_aux=row_aux, grp_off=grp_off, grp_rows=grp_rows)

    def _iter_parameter_priors(self):
        for g in self._parameter_priors:
            for st in self._parameter_priors[g]:
                mu, sg = self._parameter_priors[g][st]
                yield g, st, mu, sg

    @property
    def plan_tables(self):
        return self._plan_tables

    @property
    def device_plan(self):
        """Device copy of the plan; created on first use (needs the native library and a GPU)."""
        if self._device_plan is None:
            if len(self._experiments) == 0:
                raise ValueError('Project has no experiments')
            self._device_plan = engine.DevicePlan(self._model.compiled, self._plan_tables)
        return self._device_plan

    def _jac_solver(self, thetas=None):
        """Integrator of the sensitivity path: the project's stiff choice carries over to the Jacobian (RODAS4 with forward
        sensitivities, autonomous models of at most 12 states - the C ABI refuses other models), otherwise Dormand-Prince."""
        return self._concrete_solver(True, thetas)

    def _concrete_solver(self, with_jac, thetas=None, requested=None):
        """'dp5' / 'rodas4' for this call.  ``solver='auto'`` (the reference's LSODA switches between Adams and BDF by itself,
        ode_model.py:122-123) is decided ONCE per project and path by measurement: the first parameter sets that come in (at
        most 64) are integrated with both kernels under a bounded step budget, and the faster one that fails nowhere is kept
        (``auto_solver_report``).  A stiff model exhausts the explicit kernel's budget and selects the Rosenbrock kernel."""
        s = str(requested or self.solver).lower()
        if s in ('rodas', 'rodas4'):
            return 'rodas4'
        if s != 'auto':
            return 'dp5' if with_jac else s
        if with_jac in self._auto_choice:
            return self._auto_choice[with_jac]
        if thetas is None:
            return 'dp5'
        import torch
        _lib.require_gpu()
        plan, t = self.device_plan, self._plan_tables
        th = thetas if isinstance(thetas, torch.Tensor) else torch.from_numpy(
            np.ascontiguousarray(np.atleast_2d(np.asarray(thetas, dtype=np.float64))))
        th = th[:64].to('cuda').contiguous()
        n = th.shape[0]
        f64 = dict(dtype=torch.float64, device=th.device)
        report = {}
        for cand in ('dp5', 'rodas4'):
            sims = torch.zeros((n, t.n_rows), **f64)
            jac = torch.empty((n, t.n_rows, t.n_proj), **f64) if with_jac else None
            status = torch.zeros((n, t.E), dtype=torch.int32, device=th.device)
            counters = torch.zeros(_lib.SBM_N_COUNTERS, dtype=torch.int64, device=th.device)
            opts = engine.DevicePlan.opts(cand, self.rtol, self.atol, 50000)
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            try:
                plan.solve_device(th, opts, sims, jac, status, counters)           # (first call: module load)
                sims.zero_()
                ev0.record()
                plan.solve_device(th, opts, sims, jac, status, counters)
                ev1.record()
                torch.cuda.synchronize()
            except _lib.SbmError as exc:                                            # kernel not available for this model
                report[cand] = {'available': False, 'why': str(exc)}
                continue
            report[cand] = {'available': True, 'ms': ev0.elapsed_time(ev1), 'failed': int((status != 0).sum().item())}
        ok = [c for c in report if report[c]['available'] and report[c]['failed'] == 0]
        if ok:
            choice = min(ok, key=lambda c: report[c]['ms'])
        else:
            avail = [c for c in report if report[c]['available']] or ['dp5']
            choice = min(avail, key=lambda c: report[c]['failed'])
        self._auto_choice[with_jac] = choice
        self.auto_solver_report[with_jac] = dict(report, choice=choice, sets=n)
        return choice

    def _opts(self, solver=None):
        return engine.DevicePlan.opts(solver or self.solver, self.rtol, self.atol, self.max_steps)

    # ------------------------------------------------------------------------------------------
    # setters / getters (reference :536-697)
    # ------------------------------------------------------------------------------------------
    @property
    def project_param_idx(self):
        return copy.deepcopy(self._project_param_idx)

    @property
    def n_project_params(self):
        return self._n_project_params

    @property
    def n_project_residuals(self):
        return self._n_residuals

    @property
    def scale_factors(self):
        try:
            return self._loss_function.scale_factors
        except AttributeError:
            raise AttributeError("%s type doesn't support scale factors" % type(self._loss_function))

    @property
    def experiments(self):
        return iter(self._experiments)

    @property
    def project_param_vector(self):
        return np.copy(self._project_param_vector)

    @property
    def parameter_priors(self):
        return copy.deepcopy(self._parameter_priors)

    @property
    def measurements_df(self):
        t = self._plan_tables
        return pd.DataFrame({'mean': t.data.copy(), 'std': self._meas_sigma_raw.copy(),
                             'timepoints': self._meas_t_raw()}, index=self._index)

    def _meas_t_raw(self):
        out = []
        for experiment in self._experiments:
            for m in experiment.measurements:
                out.extend(m.get_nonzero_measurements()[2].tolist())
        out += [np.nan] * (self._plan_tables.n_rows - len(out))
        return np.array(out, dtype=np.float64)

    def _require_last(self, what):
        if self._last is None or what not in self._last:
            raise ValueError("No %s executed yet" % ('simulations' if what == 'sims' else 'jacobian calculations'))

    def get_simulations(self, scaled=False, include_priors=False):
        """Reference :568-579: frame with columns ('mean', 'timepoints') of the last evaluation."""
        self._require_last('sims')
        sims = self._last['sims'].copy()
        if scaled:
            t = self._plan_tables
            for r in range(self._n_meas_rows):
                if t.row_group[r] >= 0:
                    sims[r] *= self._last['sf'][t.row_group[r]]
        out = pd.DataFrame({'mean': sims, 'timepoints': self._meas_t.copy()}, index=self._index)
        if not include_priors:
            out = out.iloc[:self._n_meas_rows]
        return out

    def get_model_jacobian_df(self, include_priors=False):
        """Reference :585-593: the unscaled model Jacobian of the last ``calc_project_jacobian``."""
        self._require_last('model_jac')
        out = pd.DataFrame(self._last['model_jac'].copy(), index=self._index,
                           columns=self.get_ordered_project_params())
        if not include_priors:
            out = out.iloc[:self._n_meas_rows]
        return out

    def add_experiment(self, experiment):
        if type(experiment) is not list:
            experiment = [experiment]
        names = {e.name for e in self._experiments}
        for e in experiment:
            if e.name in names:
                raise KeyError("An Experiment with name %s is already present" % e.name)
            self._experiments.append(e)
            names.add(e.name)
        self._experiments.sort(key=lambda x: x.name)       # reference :624
        self._update_project_settings()

    def remove_experiments_by_settings(self, settings):
        removed = []
        for i, experiment in enumerate(self._experiments):
            all_equal = True
            for s in settings:
                if s not in experiment.settings:
                    raise KeyError('%s is not a setting in experiment %s' % (s, experiment.name))
                if not (experiment.settings[s] == settings[s]):
                    all_equal = False
            if all_equal:
                removed.append(i)
        if not removed:
            raise KeyError('None of the settings chosen were in the experiments in the project')
        deleted = [self._experiments.pop(i) for i in reversed(removed)]
        if len(self._experiments) == 0:
            warnings.warn('Project has no more experiments')
            self._project_param_idx, self._n_project_params, self._residuals_per_param = self._set_local_param_idx()
            self._n_residuals = 0
            self._plan_tables, self._device_plan, self._last = None, None, None
        else:
            self._update_project_settings()
        return deleted

    def get_experiment(self, exp_idx):
        return copy.deepcopy(self._experiments[exp_idx])

    def get_experiment_index(self, exp_name):
        for i, experiment in enumerate(self._experiments):
            if exp_name == experiment.name:
                return i
        raise KeyError('%s not in experiments' % exp_name)

    def set_scale_factor_log_prior(self, measure_name, log_scale_factor_prior, log_sigma_scale_factor):
        self._loss_function.set_scale_factor_priors(measure_name, log_scale_factor_prior, log_sigma_scale_factor)
        self._scale_factor_priors.append(measure_name)
        self._update_project_settings()

    def set_parameter_log_prior(self, p_group, settings, log_scale_parameter_prior, log_sigma_parameter):
        try:
            self._project_param_idx[p_group][settings]
        except KeyError:
            raise KeyError('%s with settings %s not in the project parameters' % (p_group, settings))
        self._parameter_priors.setdefault(p_group, OrderedDict())
        self._parameter_priors[p_group][settings] = (log_scale_parameter_prior, log_sigma_parameter)
        self._update_project_settings()

    def get_param_index(self, parameter_group, settings='all'):
        if isinstance(settings, str) and settings == 'all':
            return self._project_param_idx[parameter_group]
        return self._project_param_idx[parameter_group][settings]

    def get_experiment_parameters(self, experiment):
        """Reference :343-363 (host copy, for inspection; the kernels do this gather themselves)."""
        out = np.zeros((len(self._model.param_order),))
        for m, name in enumerate(self._model.param_order):
            try:
                out[m] = np.exp(self._project_param_vector[experiment.param_global_vector_idx[name]])
            except KeyError:
                out[m] = experiment.fixed_parameters[name]
        return out

    def reset_calcs(self):
        self._last = None
        self._project_param_vector = np.zeros((self.n_project_params,))

    # ------------------------------------------------------------------------------------------
    # evaluation
    # ------------------------------------------------------------------------------------------
    def _flags(self):
        return (_lib.SBM_FLAG_PRIOR_JAC_UNIT if self.prior_jacobian_unit else 0) | \
            (_lib.SBM_FLAG_LOSS_GENERIC if getattr(self, 'loss_generic_kernel', False) else 0)

    def _loss_name(self):
        # 'normalized' is folded into the plan's sigma, the kernel then runs the square loss
        return 'square' if self._loss_kind == 'normalized' else self._loss_kind

    def _warn_failed(self, status):
        """Numerical failure is not an exception in the reference - it is inf residuals / Jacobian - but
        a silent inf hides WHY: surface the integrator's per-trajectory status (max_steps: typically a stiff region the
        explicit integrator cannot cross - solver='rodas4' / 'auto' select the Rosenbrock kernels)."""
        status = np.asarray(status)
        if status.size and np.any(status != 0):
            bad = status[status != 0]
            names = sorted(set(_lib.STATUS_NAMES.get(int(x), str(int(x))) for x in bad))
            warnings.warn('%d of %d trajectories failed in the integrator (%s): their parameter sets get inf residuals; '
                          'with the explicit integrator (solver="dp5") a stiff model exhausts its step budget - use '
                          'solver="rodas4" or "auto" (sensitivities: autonomous models of at most 12 states)'
                          % (bad.size, status.size, ', '.join(names)), RuntimeWarning, stacklevel=3)

    def _store_last(self, theta, out, with_jac):
        self._warn_failed(out['status'])
        self._project_param_vector = np.copy(theta)
        last = {'sims': out['sims'][0], 'resid': out['resid'][0], 'sf': out['sf'][0], 'status': out['status'][0],
                'counters': out['counters']}
        if with_jac:
            last['model_jac'] = out['model_jac'][0]
            last['jac'] = out['jac'][0]
            last['grad'] = out['grad'][0]
        if self._last is not None and not with_jac and 'model_jac' in self._last and \
                np.array_equal(self._last.get('theta'), theta):
            last['model_jac'] = self._last['model_jac']
        last['theta'] = np.copy(theta)
        self._last = last
        if hasattr(self._loss_function, 'store_results') and self._plan_tables.n_groups:
            self._loss_function.store_results(out['sf'][0])

    def residuals(self, project_param_vector):
        """Reference :708-729.  One plain solve per experiment + loss epilogue, all on the GPU."""
        theta = np.asarray(project_param_vector, dtype=np.float64)
        out = self.device_plan.residuals_host(theta[None, :], self._opts(self._concrete_solver(False, theta)), self._loss_name())
        self._store_last(theta, out, False)
        res = out['resid'][0]
        if np.isinf(res).all() and len(res):
            return res                              # reference returns a bare inf array on NaN simulations
        return pd.Series(res, index=self._index)

    def calc_project_jacobian(self, project_param_vector):
        """Reference :731-771.  One sensitivity-augmented solve per experiment (its state column doubles as
        the plain simulation, so the reference's extra ``_sim_experiments`` pass is not repeated)."""
        theta = np.asarray(project_param_vector, dtype=np.float64)
        out = self.device_plan.jacobian_host(theta[None, :], self._opts(self._jac_solver(theta)), self._loss_name(), self._flags())
        self._store_last(theta, out, True)
        return pd.DataFrame(out['jac'][0], index=self._index, columns=self.get_ordered_project_params())

    def calc_rss_gradient(self, project_param_vector, *args):
        """Reference :773-805: (J^T r) - formed in the loss epilogue from the same solve."""
        theta = np.asarray(project_param_vector, dtype=np.float64)
        out = self.device_plan.jacobian_host(theta[None, :], self._opts(self._jac_solver(theta)), self._loss_name(), self._flags())
        self._store_last(theta, out, True)
        return pd.Series(out['grad'][0], index=self.get_ordered_project_params())

    def calc_sum_square_residuals(self, project_param_vector):
        """Reference :807-827: 0.5 * sum(residuals^2)."""
        theta = np.asarray(project_param_vector, dtype=np.float64)
        out = self.device_plan.residuals_host(theta[None, :], self._opts(self._concrete_solver(False, theta)), self._loss_name())
        self._store_last(theta, out, False)
        return float(out['ssr'][0])

    def nlopt_fcn(self, project_param_vector, grad):
        """Reference :829-852."""
        if grad.size > 0:
            grad[:] = np.asarray(self.calc_rss_gradient(project_param_vector))
        return self.calc_sum_square_residuals(project_param_vector)

    def calc_scale_factors_entropy(self, temperature=1.0):
        """Reference :854-872 -> abstract_loss_function.py:122-134 (host quadrature, sampler-only path)."""
        self._require_last('sims')
        lf, t = self._loss_function, self._plan_tables
        entropy = 0.0
        for g, key in enumerate(lf.group_keys()):
            rows = t.grp_rows[t.grp_off[g]:t.grp_off[g + 1]]
            s, d, sg = self._last['sims'][rows], t.data[rows], t.sigma[rows]
            holder = lf._scale_factors[lf.group_members(key)[0]]
            entropy += temperature * holder.calc_scale_factor_entropy(np.sum(s * s / sg ** 2),
                                                                      np.sum(s * d / sg ** 2), temperature)
        return entropy

    def free_energy(self, project_param_vector, temperature=1):
        """Reference :874-892."""
        rss = self.calc_sum_square_residuals(project_param_vector)
        return rss - self.calc_scale_factors_entropy(temperature)

    # ---- additive batch API -------------------------------------------------------------------
    def residuals_batch(self, thetas, solver=None):
        """thetas f8[B, n_proj] -> dict(resid [B, n_rows], ssr [B], sf [B, G], sims, status [B, E], counters)."""
        out = self.device_plan.residuals_host(np.asarray(thetas, dtype=np.float64),
                                              self._opts(self._concrete_solver(False, thetas, solver)), self._loss_name())
        self._warn_failed(out['status'])
        return out

    def ssr_batch(self, thetas, solver=None):
        return self.residuals_batch(thetas, solver)['ssr']

    def jacobian_batch(self, thetas):
        """thetas f8[B, n_proj] -> dict(jac [B, n_rows, n_proj], model_jac, resid, ssr, grad [B, n_proj], ...)."""
        out = self.device_plan.jacobian_host(np.asarray(thetas, dtype=np.float64), self._opts(self._jac_solver(thetas)),
                                             self._loss_name(), self._flags())
        self._warn_failed(out['status'])
        return out

    def rss_gradient_batch(self, thetas):
        return self.jacobian_batch(thetas)['grad']

    def evaluate_device(self, theta, with_jacobian=True, solver=None, out=None):
        """Device-resident evaluation: ``theta`` is a CUDA float64 tensor [B, n_proj]; returns a dict of
        CUDA tensors (no host copies, no synchronisation).  ``out`` lets the caller reuse buffers."""
        import torch
        plan, t = self.device_plan, self._plan_tables
        B = theta.shape[0]
        f64 = dict(dtype=torch.float64, device=theta.device)
        o = out if out is not None else {}

        def buf(name, shape, dtype=None, zero=False):
            x = o.get(name)
            if x is None or tuple(x.shape) != tuple(shape):
                x = torch.empty(shape, dtype=dtype or torch.float64, device=theta.device)
                o[name] = x
            if zero:
                x.zero_()
            return x
        sims = buf('sims', (B, t.n_rows), zero=True)
        resid = buf('resid', (B, t.n_rows))
        ssr = buf('ssr', (B,))
        sf = buf('sf', (B, max(t.n_groups, 1)))
        status = buf('status', (B, t.E), torch.int32)
        counters = buf('counters', (_lib.SBM_N_COUNTERS,), torch.int64, zero=True)
        jac = sfgrad = grad = None
        if with_jacobian:
            jac = buf('jac', (B, t.n_rows, t.n_proj))
            sfgrad = buf('sfgrad', (B, max(t.n_groups, 1), t.n_proj))
            grad = buf('grad', (B, t.n_proj))
        opts = self._opts(self._jac_solver(theta) if with_jacobian else self._concrete_solver(False, theta, solver))
        plan.solve_device(theta, opts, sims, jac, status, counters)
        plan.loss_device(theta, sims, jac, resid, ssr, sf, sfgrad, grad, loss=self._loss_name(), flags=self._flags())
        return o

    def lm_fit_batch(self, thetas, max_iter=100, lam0=1e-3, ftol=1e-12, xtol=1e-10, gtol=0.0, max_log_step=3.0):
        """Batched multistart Levenberg-Marquardt on the device: the consumer the reference drives through
        ``scipy.optimize.leastsq(proj.residuals, x0, Dfun=proj.calc_project_jacobian)`` (tests/test_Project.py:202-213,
        351-357), for B start vectors at once.  Every iteration is one fused evaluation (integrator + loss epilogue) of
        all trial points and one ``sbm_lm_step`` launch; parameters, residuals and Jacobians never leave the GPU.

        The step solves ``(J^T J + lam D) delta = -J^T r`` with Marquardt's ``D = diag(J^T J)`` kept at its running
        maximum over the iterations (MINPACK's column scaling: a parameter drifting into a flat direction keeps its
        damping); ``lam`` follows
        Nielsen's gain-ratio rule (accepted: ``lam *= max(1/3, 1 - (2 rho - 1)^3)``; rejected: ``lam *= nu, nu *= 2``).
        A set stops when an accepted step reduces SSR by less than ``ftol`` (relative), its step is below ``xtol``
        (relative, per parameter), ``max |J^T r| <= gtol`` or ``lam`` overflows.  A step that would change a log-parameter
        by more than ``max_log_step`` is not tried (its damping is raised instead): a nearly undamped first step can land
        where the model is extremely stiff, and the whole batch would wait for those trajectories.  J and r are exactly what the reference's
        callbacks return (the scaled Jacobian is not divided by sigma), so the fixed points are scipy's.

        Returns a dict of numpy arrays: theta [B, n_proj], ssr [B], n_iter [B], converged [B], lam [B]."""
        import torch
        _lib.require_gpu()
        dev = torch.device('cuda')
        th = torch.from_numpy(np.ascontiguousarray(np.atleast_2d(thetas), dtype=np.float64)).to(dev)
        B, P = th.shape
        if P != self.n_project_params:
            raise ValueError('expected %d project parameters, got %d' % (self.n_project_params, P))
        f64 = dict(dtype=torch.float64, device=dev)
        cur, tri = {}, {}
        self.evaluate_device(th, True, out=cur)
        lam = torch.full((B,), float(lam0), **f64)
        nu = torch.full((B,), 2.0, **f64)
        dscale = torch.zeros((B, P), **f64)
        active = torch.isfinite(cur['ssr'])
        converged = torch.zeros((B,), dtype=torch.bool, device=dev)
        n_iter = torch.zeros((B,), dtype=torch.int64, device=dev)
        step = None
        keys = ('sims', 'resid', 'ssr', 'sf', 'jac', 'sfgrad', 'grad', 'status')
        for it in range(int(max_iter)):
            step = engine.lm_step_device(cur['jac'], cur['resid'], lam, dscale, *(step or ()))
            delta, pred, gmax, info = step
            step_ok = active & (info == 0) & (delta.abs().amax(dim=1) <= max_log_step)
            th_try = torch.where(step_ok[:, None], th + delta, th)
            self.evaluate_device(th_try, True, out=tri)
            ssr0, ssr1 = cur['ssr'], tri['ssr']
            accept = step_ok & torch.isfinite(ssr1) & (ssr1 < ssr0) & (pred > 0)
            rho = torch.where(accept, (ssr0 - ssr1) / torch.where(pred > 0, pred, torch.ones_like(pred)), torch.zeros_like(pred))
            small = (delta.abs() <= xtol * (th.abs() + xtol)).all(dim=1)
            done = (accept & ((ssr0 - ssr1) <= ftol * ssr0)) | (step_ok & small) | (active & (gmax <= gtol))
            lam_new = torch.where(accept, lam * torch.clamp(1.0 - (2.0 * rho - 1.0) ** 3, min=1.0 / 3.0), lam * nu)
            nu = torch.where(accept, torch.full_like(nu, 2.0), nu * 2.0)
            lam = torch.where(active, lam_new, lam)
            th = torch.where(accept[:, None], th_try, th)
            for k in keys:
                if k in cur and k in tri:
                    m = accept.reshape((B,) + (1,) * (cur[k].dim() - 1))
                    cur[k].copy_(torch.where(m, tri[k], cur[k]))
            n_iter += active.to(torch.int64)
            converged |= active & done
            active = active & ~done & (lam < 1e16)
            if (it % 4 == 3 or it == max_iter - 1) and not bool(active.any()):
                break
        self._last_lm = {'theta': th.cpu().numpy(), 'ssr': cur['ssr'].cpu().numpy(), 'n_iter': n_iter.cpu().numpy(),
                         'converged': converged.cpu().numpy(), 'lam': lam.cpu().numpy()}
        return self._last_lm

    # ------------------------------------------------------------------------------------------
    # dict <-> vector helpers (reference :898-972)
    # ------------------------------------------------------------------------------------------
    def project_param_dict_to_vect(self, param_dict, default_value=0.0):
        vec = np.ones((self.n_project_params,)) * default_value
        for g in self._project_param_idx:
            for st, i in self._project_param_idx[g].items():
                try:
                    vec[i] = param_dict[g][st]
                except KeyError:
                    pass
        return vec

    def project_param_vect_to_dict(self, param_vector):
        out = {}
        for g in self._project_param_idx:
            out[g] = {}
            for st, i in self._project_param_idx[g].items():
                out[g][st] = param_vector[i]
        return out

    def group_experiments(self, settings_groups):
        grouped = defaultdict(list)
        for experiment in self._experiments:
            key = []
            for s in settings_groups:
                if s not in experiment.settings:
                    raise KeyError('%s is not a setting in experiment %s' % (s, experiment.name))
                key.append(experiment.settings[s])
            grouped[tuple(key)].append(experiment)
        return grouped
