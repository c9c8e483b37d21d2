"""CPU ORACLE - TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/SciPy restatement of the reference's hot path (FedericoV/SysBio_Modeling), used only
as the checker by ``tests/``, ``__graft_entry__.smoke()`` and as the ``cpu_baseline`` / ``--impl
reference`` leg of ``bench.py``.  Nothing under ``sysbio_modeling_b200/`` imports this module and the
product path raises if its CUDA library is missing - there is no CPU fallback.

Parity status: PINNED.  ``tests/golden/make_golden.py`` runs the real reference (a py3-patched
throw-away copy of /root/reference made in /tmp, patch list in that script) on the reference's own
test fixtures and commits the numbers under ``tests/golden/``; ``tests/test_oracle.py`` checks this
restatement against them, against the golden vector of reference ``tests/test_OdeModel.py:25-26`` and
against the closed forms of reference ``tests/test_Project.py:19-23``.

The numerical integrator of the reference is a third-party dependency that is not vendored in
/root/reference: SciPy's ``scipy.integrate.odeint`` = ODEPACK LSODA (version unpinned by the
reference; 1.18.1 in this image).  The oracle calls the same entry point with the reference's
arguments (``ode_model.py:122-123,167-168``: ``rtol=atol=1e-10, col_deriv=True``).

Every function cites the reference file:line it follows.
"""
from collections import OrderedDict

import numpy as np
from scipy.integrate import odeint

REF_RTOL = 1e-10
REF_ATOL = 1e-10
N_GRID = 1000   # base_project.py:419


# --------------------------------------------------------------------------------------------
# L1/L2: integrate (model/ode_model.py)
# --------------------------------------------------------------------------------------------
def simulate(model, n_vars, experiment_params, t_sim, init_conditions=None, rtol=REF_RTOL, atol=REF_ATOL):
    """ode_model.py:128-169 - plain solve, y(0) = 0 unless given."""
    if init_conditions is None:
        init_conditions = np.zeros((n_vars,))
    p = np.asarray(experiment_params, dtype=np.float64)
    yout = np.zeros_like(np.asarray(init_conditions, dtype=np.float64))

    def func_wrapper(y, t):
        model(y, t, yout, p)
        return yout
    return odeint(func_wrapper, np.asarray(init_conditions, dtype=np.float64), t_sim,
                  rtol=rtol, atol=atol)


def calc_jacobian(sens_model, n_vars, experiment_params, t_sim, init_conditions, rtol=REF_RTOL,
                  atol=REF_ATOL, return_states=False):
    """ode_model.py:83-126 - augmented solve; returns sol[:, n_vars:] (column i*P_free + j)."""
    p = np.asarray(experiment_params, dtype=np.float64)
    y0 = np.asarray(init_conditions, dtype=np.float64)
    yout = np.zeros_like(y0)

    def func_wrapper(y, t):
        sens_model(y, t, yout, p)
        return yout
    sol = odeint(func_wrapper, y0, t_sim, rtol=rtol, atol=atol)
    if return_states:
        return sol[:, :n_vars], sol[:, n_vars:]
    return sol[:, n_vars:]


# --------------------------------------------------------------------------------------------
# Mapping functions (project/utils.py)
# --------------------------------------------------------------------------------------------
def snapped_index(t_sim, t_meas_nonzero):
    """utils.py:18-20 - measurement times are snapped UP to the next node of the simulation grid."""
    return np.searchsorted(t_sim, t_meas_nonzero)


def map_sim(kind, args, model_sim, t_sim, t_meas):
    """utils.py:10-26 (direct) and :48-68 (sum).  Returns (mapped values, snapped times)."""
    idx = snapped_index(t_sim, t_meas)
    if kind == 'direct':
        return model_sim[idx, args], t_sim[idx]
    out = np.zeros((len(idx),))
    for v in args:
        out += model_sim[idx, v]
    return out, t_sim[idx]


def map_jac(kind, args, model_jac, t_sim, t_meas, n_exp_params):
    """utils.py:29-45 (direct) and :71-89 (sum).  Returns (k, n_exp_params)."""
    idx = snapped_index(t_sim, t_meas)
    if kind == 'direct':
        start = args * n_exp_params
        return model_jac[idx, start:start + n_exp_params]
    out = np.zeros((len(idx), n_exp_params))
    for v in args:
        out += model_jac[idx, v * n_exp_params:(v + 1) * n_exp_params]
    return out


# --------------------------------------------------------------------------------------------
# Scale factors (project/loss_functions/squared_loss/*.py)
# --------------------------------------------------------------------------------------------
def linear_sf(sim, exp, std):
    """linear_scale_factor.py:27-31."""
    return np.sum(sim * exp / std ** 2) / np.sum(sim * sim / std ** 2)


def linear_sf_gradient(sim, exp, std, jac):
    """linear_scale_factor.py:33-42 (note the factor 2)."""
    sde = np.sum(sim * exp / std ** 2)
    sds = np.sum(sim * sim / std ** 2)
    jde = np.sum(jac.T * exp / std ** 2, axis=1)
    jds = np.sum(jac.T * sim / std ** 2, axis=1)
    return jde / sds - 2 * sde * jds / sds ** 2


def log_sf(sim, exp, std):
    """log_scale_factor.py:17-27 (sigma is rescaled by the data: std/exp)."""
    s = std / exp
    inv = np.sum(1.0 / s ** 2)
    return np.exp(np.sum(np.log(exp) / s ** 2) / inv) * np.exp(-np.sum(np.log(sim) / s ** 2) / inv)


def log_sf_gradient(sim, exp, std, jac, sf):
    """log_scale_factor.py:29-36."""
    s = std / exp
    inv = np.sum(1.0 / s ** 2)
    return sf * (-1.0 / inv) * np.sum(jac.T / (sim * s ** 2), axis=1)


# --------------------------------------------------------------------------------------------
# Project (project/base_project.py) restated on flat arrays
# --------------------------------------------------------------------------------------------
class OracleProject(object):
    """Restatement of ``Project`` (base_project.py) + ``SquareLossFunction`` and its variants.

    ``model`` / ``sens_model`` are reference-ABI callables ``f(y, t, yout, p)``.
    ``experiments`` are duck-typed: ``.name``, ``.measurements`` (each ``.variable_name .values .std
    .timepoints``), ``.fixed_parameters``, ``.settings``, optional ``.initial_conditions``.

    ``reference_jac_rows=True`` reproduces the reference's Jacobian row bookkeeping exactly
    (base_project.py:467,488,516: the row cursor is not advanced between the measurements of one
    experiment, and the experiment cursor advances by the LAST measurement's length); ``False``
    gives the mathematically intended rows.  Both agree for single-measurement experiments.
    """

    def __init__(self, model, sens_model, n_vars, param_order, experiments, model_parameter_settings,
                 measurement_to_model_map, sf_groups=None, loss='square', fixed_params=(),
                 reference_jac_rows=False, rtol=REF_RTOL, atol=REF_ATOL):
        self.model, self.sens_model = model, sens_model
        self.n_vars, self.param_order = n_vars, list(param_order)
        self.experiments = sorted(experiments, key=lambda e: e.name)        # :624
        self.settings = model_parameter_settings
        self.mapping = measurement_to_model_map
        self.loss = loss
        self.sf_groups = [g if isinstance(g, str) else frozenset(g) for g in (sf_groups or [])]
        self.reference_jac_rows = reference_jac_rows
        self.rtol, self.atol = rtol, atol
        self.param_priors = OrderedDict()       # (group, settings) -> (mu, sigma)   :681-691
        self.sf_priors = OrderedDict()          # group key -> (mu, sigma)            :675-679
        self._set_local_param_idx()
        self.sf = {}
        self.sf_grad = {}

    # -- :164-276 -----------------------------------------------------------------------------
    def _set_local_param_idx(self):
        s = self.settings
        local_pars = list(s.get('Local', []))
        fixed_pars = list(s.get('Fixed', []))
        global_pars = list(s.get('Global', []))
        shared = s.get('Shared', {})
        shared_names = set(p for g in shared for p in shared[g])
        named = set(local_pars) | set(fixed_pars) | shared_names | set(global_pars)
        # :189-197 - unspecified parameters default to Global (deterministic param_order order here)
        global_pars = global_pars + [p for p in self.param_order if p not in named]
        idx = OrderedDict()
        n = 0
        for p in global_pars:
            idx[p] = {'Global': n}
            n += 1
        self.exp_param_idx = []
        for e in self.experiments:
            ep = OrderedDict()
            exp_fixed = list(e.fixed_parameters.keys()) if e.fixed_parameters else []
            all_fixed = fixed_pars + exp_fixed
            for p in fixed_pars:
                if p not in exp_fixed:
                    raise ValueError('%s was declared as a fixed parameter, but in experiment %s no value provided'
                                     % (p, e.name))
            for p in global_pars:
                if p not in all_fixed:
                    ep[p] = idx[p]['Global']
            for g in shared:
                for p, settings in shared[g].items():
                    if p in exp_fixed:
                        continue
                    key = 'None' if settings is None else tuple(e.settings[k] for k in tuple(settings))
                    idx.setdefault(g, {})
                    if key not in idx[g]:
                        idx[g][key] = n
                        n += 1
                    ep[p] = idx[g][key]
            for p in local_pars:
                if p in exp_fixed:
                    continue
                idx['%s_%s' % (p, e.name)] = {'Local': n}
                ep[p] = n
                n += 1
            self.exp_param_idx.append(ep)
        self.project_param_idx = idx
        self.n_project_params = n
        self._layout_rows()

    # -- :296-341 -----------------------------------------------------------------------------
    def _layout_rows(self):
        labels, mean, std, tp = [], [], [], []
        for e in self.experiments:
            for m in e.measurements:
                keep = m.timepoints != 0
                for v, s_, t_ in zip(m.values[keep], m.std[keep], m.timepoints[keep]):
                    labels.append((e.name, m.variable_name))
                    mean.append(v), std.append(s_), tp.append(t_)
        self.n_meas_rows = len(labels)
        prior_rows = []
        for (g, st), (mu, sg) in self.param_priors.items():
            name = g + ' ' + ''.join(str(x) for x in st)
            prior_rows.append((('~Prior', name), mu, sg))
        prior_rows.sort(key=lambda r: r[0])      # the measurement frame is lexsorted (:340)
        for lab, mu, sg in prior_rows:
            labels.append(lab), mean.append(mu), std.append(sg), tp.append(np.nan)
        sfp = []
        for g, (mu, sg) in self.sf_priors.items():
            name = g if isinstance(g, str) else next(iter(g))
            sfp.append((('~~SF_Prior', '~%s' % name), mu, sg, g))
        sfp.sort(key=lambda r: r[0])
        self.sf_prior_rows = {}
        for lab, mu, sg, g in sfp:
            self.sf_prior_rows[g] = len(labels)
            labels.append(lab), mean.append(mu), std.append(sg), tp.append(np.nan)
        self.labels = labels
        self.meas_mean = np.array(mean, dtype=np.float64)
        self.meas_std = np.array(std, dtype=np.float64)
        self.meas_t = np.array(tp, dtype=np.float64)
        self.n_rows = len(labels)
        # rows of each scale-factor group
        self.group_rows = {}
        for g in self.sf_groups:
            names = [g] if isinstance(g, str) else list(g)
            rows = []
            for name in names:     # measure by measure, as abstract_loss_function.py:72-76
                rows += [r for r in range(self.n_meas_rows) if labels[r][1] == name]
            self.group_rows[g] = np.array(rows, dtype=np.int64)

    def set_parameter_log_prior(self, p_group, settings, mu, sigma):
        self.project_param_idx[p_group][settings]
        self.param_priors[(p_group, settings)] = (mu, sigma)
        self._layout_rows()

    def set_scale_factor_log_prior(self, group, mu, sigma):
        key = None
        for g in self.sf_groups:
            if g == group or (not isinstance(g, str) and group in g):
                key = g
        if key is None:
            raise KeyError('%s not present as a scale factor' % (group,))
        self.sf_priors[key] = (mu, sigma)
        self._layout_rows()

    # -- :343-363 -----------------------------------------------------------------------------
    def experiment_parameters(self, k, theta):
        e, ep = self.experiments[k], self.exp_param_idx[k]
        p = np.zeros((len(self.param_order),))
        for m, name in enumerate(self.param_order):
            if name in ep:
                p[m] = np.exp(theta[ep[name]])
            else:
                p[m] = e.fixed_parameters[name]
        return p

    def _grid(self, e):
        t_end = np.unique(np.concatenate([m.timepoints for m in e.measurements]))
        t_end = t_end[t_end != 0][-1]                      # experiments.py:71-96
        return np.linspace(0, t_end, N_GRID)               # :418-419

    def _y0(self, e):
        y0 = getattr(e, 'initial_conditions', None)
        if y0 is None or (isinstance(y0, dict) and not y0):
            return np.zeros((self.n_vars,))
        return np.asarray(y0, dtype=np.float64)

    # -- :393-425 + :365-391 ------------------------------------------------------------------
    def sim_experiments(self, theta):
        sims = np.zeros((self.n_rows,))
        times = np.zeros((self.n_rows,))
        r = 0
        for k, e in enumerate(self.experiments):
            p = self.experiment_parameters(k, theta)
            t_sim = self._grid(e)
            y = simulate(self.model, self.n_vars, p, t_sim, self._y0(e), self.rtol, self.atol)
            for m in e.measurements:
                kind, args = self.mapping[m.variable_name]
                tm = m.timepoints[m.timepoints != 0]
                v, ts = map_sim(kind, args, y, t_sim, tm)
                sims[r:r + len(v)] = v
                times[r:r + len(v)] = ts
                r += len(v)
        # :427-437 - parameter-prior rows carry theta itself, written in insertion order
        for (g, st) in self.param_priors:
            sims[r] = theta[self.project_param_idx[g][st]]
            r += 1
        return sims, times

    # -- :490-517 + :443-488 ------------------------------------------------------------------
    def model_jacobian(self, theta):
        J = np.zeros((self.n_rows, self.n_project_params))
        dexp = np.exp(theta)                                # :450
        res_idx = 0
        for k, e in enumerate(self.experiments):
            ep = self.exp_param_idx[k]
            p = self.experiment_parameters(k, theta)
            t_sim = self._grid(e)
            free = [name for name in self.param_order if name in ep]
            n_free = len(free)
            y0 = np.concatenate([self._y0(e), np.zeros((n_free * self.n_vars,))])
            S = calc_jacobian(self.sens_model, self.n_vars, p, t_sim, y0, self.rtol, self.atol)
            row = res_idx
            last = 0
            for m in e.measurements:
                kind, args = self.mapping[m.variable_name]
                tm = m.timepoints[m.timepoints != 0]
                mj = map_jac(kind, args, S, t_sim, tm, n_free)
                last = len(mj)
                for j, name in enumerate(free):
                    c = ep[name]
                    J[row:row + last, c] += mj[:, j] * dexp[c]          # :482-485
                if not self.reference_jac_rows:
                    row += last
            res_idx = res_idx + last if self.reference_jac_rows else row
        return J

    # -- loss functions -----------------------------------------------------------------------
    def _measure_std(self):
        """normalized_squared_loss_function.py:38-46: sigma <- sigma * mean on measurement rows."""
        std = self.meas_std.copy()
        if self.loss == 'normalized':
            std[:self.n_meas_rows] *= self.meas_mean[:self.n_meas_rows]
        return std

    def _update_sf(self, sims, std):
        for g, rows in self.group_rows.items():
            if self.loss == 'log':
                self.sf[g] = log_sf(sims[rows], self.meas_mean[rows], std[rows])
            else:
                self.sf[g] = linear_sf(sims[rows], self.meas_mean[rows], std[rows])

    def residuals(self, theta, return_state=False):
        """base_project.py:708-729 -> squared_loss_function.py:27-42 (and log/normalized variants)."""
        theta = np.asarray(theta, dtype=np.float64)
        sims, times = self.sim_experiments(theta)
        std = self._measure_std()
        mean = self.meas_mean.copy()
        if self.loss != 'log' and np.isnan(sims).any():
            return np.full((self.n_rows,), np.inf)
        scaled = sims.copy()
        if self.sf_groups:
            self._update_sf(sims, std)
            for g, r in self.sf_prior_rows.items():          # abstract_loss_function.py:93-105
                sims[r] = np.log(self.sf[g])
                scaled[r] = sims[r]
            for g, rows in self.group_rows.items():          # :46-59
                scaled[rows] *= self.sf[g]
        if self.loss == 'log':                               # log_squared_loss_function.py:23-64
            if np.isnan(scaled).any():
                return np.full((self.n_rows,), np.inf)
            if mean[:self.n_meas_rows].min() <= 0:
                raise ValueError("LogSquare loss cannot handle measurements smaller or equal to zero")
            scaled[:self.n_meas_rows] = np.log(scaled[:self.n_meas_rows])
            mean[:self.n_meas_rows] = np.log(mean[:self.n_meas_rows])
        res = (scaled - mean) / std
        self.last_sims, self.last_scaled, self.last_times = sims, scaled, times
        return res

    def jacobian(self, theta):
        """base_project.py:731-771 -> squared_loss_function.py:44-80 / log variant :66-98."""
        theta = np.asarray(theta, dtype=np.float64)
        sims, _ = self.sim_experiments(theta)
        J = self.model_jacobian(theta)
        std = self._measure_std()
        if self.loss != 'log' and (np.isnan(sims).any() or np.isnan(J).any()):
            return np.full_like(J, np.inf)
        if not self.sf_groups:
            if self.loss == 'log':
                return J / sims[:, None]
            return J
        self._update_sf(sims, std)
        for g, rows in self.group_rows.items():
            if self.loss == 'log':
                self.sf_grad[g] = log_sf_gradient(sims[rows], self.meas_mean[rows], std[rows],
                                                  J[rows], self.sf[g])
            else:
                self.sf_grad[g] = linear_sf_gradient(sims[rows], self.meas_mean[rows], std[rows], J[rows])
        out = J.copy()
        for g, r in self.sf_prior_rows.items():              # abstract_loss_function.py:107-120
            out[r, :] = self.sf_grad[g] / self.sf[g]
        for g, rows in self.group_rows.items():
            if self.loss == 'log':
                out[rows] = J[rows] / sims[rows][:, None] + self.sf_grad[g] / self.sf[g]
            else:
                out[rows] = J[rows] * self.sf[g] + sims[rows][:, None] * self.sf_grad[g]   # :76
        return out

    def sum_square_residuals(self, theta):
        r = self.residuals(theta)                            # :807-827
        return 0.5 * np.sum(r ** 2)

    def rss_gradient(self, theta):
        r = self.residuals(theta)                            # :773-805
        J = self.jacobian(theta)
        return (J.T * r).sum(axis=1)

    def scale_factors_entropy(self, theta, temperature=1.0):
        """base_project.py:854-872 -> abstract_loss_function.py:122-134 -> linear_scale_factor.py:13-18, 63-80: for every
        scale-factor group T * log of the SloppyCell integral over the log scale factor (QUADPACK over the whole line in
        the reference; the finite window below holds the integrand down to exp(-72) of its peaks)."""
        from scipy.integrate import quad
        sims, _ = self.sim_experiments(theta)
        entropy = 0.0
        for g in self.sf_groups:
            rows = self.group_rows[g]
            s, d, sg_ = sims[rows], self.meas_mean[rows], self.meas_std[rows]
            sde, sds = np.sum(s * d / sg_ ** 2), np.sum(s * s / sg_ ** 2)
            B = sde / sds
            lB = np.log(B)
            if g not in self.sf_priors:
                raise TypeError('scale-factor entropy needs a log prior (as in the reference)')
            mu, sig = self.sf_priors[g]

            def integrand(u):                                # linear_scale_factor.py:13-18
                return np.exp(-sds / (2 * temperature) * (np.exp(u) * B - B) ** 2 - (u + lB - mu) ** 2 / (2 * sig ** 2))
            w = max(np.sqrt(temperature / (sds * B * B)), sig)
            pts = sorted([0.0, mu - lB])
            ans, _ = quad(integrand, pts[0] - 12 * w, pts[1] + 12 * w, points=pts, limit=1000, epsabs=0, epsrel=1e-12)
            entropy += temperature * np.log(ans)             # abstract_loss_function.py:132
        return entropy

    def free_energy(self, theta, temperature=1.0):
        return self.sum_square_residuals(theta) - self.scale_factors_entropy(theta, temperature)    # :874-892


def central_difference(fun, x, rel=None):
    """Central finite differences with the step rule of statsmodels' ``approx_fprime(centered=True)``
    (the checker the reference's tests use, tests/test_Project.py:164,175): h = eps**(1/3)*max(|x|,0.1)."""
    x = np.asarray(x, dtype=np.float64)
    f0 = np.atleast_1d(fun(x))
    out = np.zeros((len(f0), len(x)))
    eps = np.finfo(float).eps ** (1.0 / 3)
    for j in range(len(x)):
        h = eps * max(abs(x[j]), 0.1)
        xp, xm = x.copy(), x.copy()
        xp[j] += h
        xm[j] -= h
        out[:, j] = (np.atleast_1d(fun(xp)) - np.atleast_1d(fun(xm))) / (2 * h)
    return out


# ---------------------------------------------------------------------------------------------------------
# Levenberg-Marquardt.  The reference has no LM of its own: its tests hand
# Project.residuals / calc_project_jacobian to scipy.optimize.leastsq (MINPACK lmder; tests/test_Project.py:202-213,
# 351-357).  MINPACK is a third-party dependency of SciPy (unpinned by the reference); the restatement below is the
# textbook Marquardt-scaled step with Nielsen's damping update (Madsen, Nielsen, Tingleff, "Methods for non-linear
# least squares problems", 2004, algorithm 3.16), the one csrc/sbm_lm.cuh and Project.lm_fit_batch implement.  It is pinned
# by identities (tests/test_lm.py: Gauss-Newton limit = lstsq, predicted = actual reduction on linear problems) and,
# for whole fits, by agreeing with scipy.optimize.leastsq on the reference's own fixture projects.
# ---------------------------------------------------------------------------------------------------------
def lm_step(J, r, lam, dscale=None):
    """(delta, pred, gmax, info) of one damped step: (J^T J + lam D) delta = -J^T r with D_j = max(dscale_j, (J^T J)_jj)
    (1 where still 0); ``dscale`` (None: no memory) is updated in place - MINPACK's running column scaling."""
    J = np.asarray(J, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    A = J.T.dot(J)
    g = J.T.dot(r)
    D = np.diag(A).copy()
    if dscale is not None:
        D = np.where(dscale > D, dscale, D)
        ok = np.isfinite(D) & (D < 1e300)
        dscale[ok] = D[ok]
    D[~(D > 0)] = 1.0
    gmax = np.max(np.abs(g)) if g.size else 0.0
    if not (np.all(np.isfinite(A)) and np.all(np.isfinite(g)) and np.isfinite(lam)):
        return np.zeros(J.shape[1]), 0.0, gmax, 1
    try:
        L = np.linalg.cholesky(A + lam * np.diag(D))
    except np.linalg.LinAlgError:
        return np.zeros(J.shape[1]), 0.0, gmax, 1
    y = np.linalg.solve(L, -g)
    delta = np.linalg.solve(L.T, y)
    pred = 0.5 * delta.dot(lam * D * delta - g)
    return delta, pred, gmax, 0


def lm_fit(residuals, jacobian, theta0, max_iter=100, lam0=1e-3, ftol=1e-12, xtol=1e-10, gtol=0.0, max_log_step=3.0):
    """One start vector of Project.lm_fit_batch, same rules, on the CPU.  Returns (theta, ssr, n_iter, converged)."""
    th = np.asarray(theta0, dtype=np.float64).copy()
    r, J = np.asarray(residuals(th)), np.asarray(jacobian(th))
    ssr = 0.5 * np.sum(r ** 2)
    lam, nu = float(lam0), 2.0
    dscale = np.zeros(th.shape)
    if not np.isfinite(ssr):
        return th, ssr, 0, False
    for it in range(1, int(max_iter) + 1):
        delta, pred, gmax, info = lm_step(J, r, lam, dscale)
        step_ok = info == 0 and bool(np.max(np.abs(delta)) <= max_log_step)
        accept = False
        if step_ok:
            r1 = np.asarray(residuals(th + delta))
            ssr1 = 0.5 * np.sum(r1 ** 2)
            accept = bool(np.isfinite(ssr1) and ssr1 < ssr and pred > 0)
        small = step_ok and bool(np.all(np.abs(delta) <= xtol * (np.abs(th) + xtol)))
        done = (accept and (ssr - ssr1) <= ftol * ssr) or small or gmax <= gtol
        if accept:
            rho = (ssr - ssr1) / pred
            lam *= max(1.0 / 3.0, 1.0 - (2.0 * rho - 1.0) ** 3)
            nu = 2.0
            th = th + delta
            r, ssr = r1, ssr1
            J = np.asarray(jacobian(th))
        else:
            lam *= nu
            nu *= 2.0
        if done:
            return th, ssr, it, True
        if not lam < 1e16:
            return th, ssr, it, False
    return th, ssr, int(max_iter), False


# ---------------------------------------------------------------------------------------------------------
# Metropolis ensemble sampler: restatement of project/Ensembles.py:55-178 with
# _sampling_matrix (:226-264), _trial_move (:266-270), _accept_move (:193-198) for ONE walker fed with given random
# numbers, with the two defects SURVEY.md lists fixed as in the product (curr_F follows accepted moves; Hessian = J^T J).
# ---------------------------------------------------------------------------------------------------------
def sampling_matrix(hessian, cutoff=0, temperature=1, step_scale=1):
    u, sing_vals, vh = np.linalg.svd(0.5 * np.asarray(hessian, dtype=np.float64))          # :229
    cutoff_sing_val = cutoff * max(sing_vals)                                                # :236
    D = 1.0 / np.maximum(sing_vals, cutoff_sing_val)                                         # :238
    samp_mat = np.transpose(vh) * np.sqrt(D)                                                 # :243
    cutoff_vals = np.compress(sing_vals < cutoff_sing_val, sing_vals)                        # :247
    if len(cutoff_vals):
        scale = np.sqrt(len(sing_vals) - len(cutoff_vals) + sum(cutoff_vals) / cutoff_sing_val)
    else:
        scale = np.sqrt(len(sing_vals))
    samp_mat = samp_mat / scale                                                              # :254
    samp_mat = samp_mat * step_scale                                                         # :255
    samp_mat = samp_mat * np.sqrt(temperature)                                               # :256
    return samp_mat


def ensemble_walker(free_energy, params, samp_mat, normals, uniforms, temperature=1.0, skip_elems=0):
    """One walker: normals f8[steps, n], uniforms f8[steps].  Returns (ens, ens_fes, ratio)."""
    cur = np.asarray(params, dtype=np.float64).copy()
    curF = free_energy(cur)
    ens, fes, accepted = [cur.copy()], [curF], 0
    for it in range(len(normals)):
        nxt = cur + samp_mat.dot(normals[it])                                                # :266-270, :134
        nxtF = free_energy(nxt)
        if not np.isfinite(nxtF):
            nxtF = np.inf
        if uniforms[it] < np.exp(-(nxtF - curF) / temperature):                              # :193-198
            cur, curF = nxt, nxtF
            accepted += 1
        if (it + 1) % (skip_elems + 1) == 0:
            ens.append(cur.copy())
            fes.append(curF)
    return np.array(ens), np.array(fes), accepted / max(len(normals), 1)
