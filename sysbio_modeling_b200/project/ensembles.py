"""Batched Metropolis ensemble sampler - the caller that generates SURVEY.md's C5 workload (section 8f rank 3).
The Monte-Carlo loop runs on the device (csrc/sbm_sampler.cuh: ``sbm_mc_propose`` / ``sbm_mc_accept`` around the
integrator and the loss epilogue); ``_entropy_device`` below is the torch form of the entropy quadrature, kept as a
cross-check of the kernel's.

Reference: ``project/Ensembles.py:55-178`` (``ensemble_log_params``: ONE walker, one ``Project.free_energy`` call per
Monte-Carlo step) with ``_sampling_matrix`` (``:226-264``), ``_trial_move`` (``:266-270``) and ``_accept_move``
(``:193-198``).  Here W independent walkers advance in lock-step: every step is one fused evaluation of W candidate
points on the GPU (integrator + loss epilogue, residual path), the walkers stay resident in HBM, and accept / reject is a few
element-wise operations.  Same proposal density (Gaussian with the axes of the Gauss-Newton Hessian, truncated at
``sing_val_cutoff``), same Metropolis rule.

Two defects of the reference are not reproduced: ``curr_F`` is never updated after an accepted move
(``:151-154``; every candidate is compared with the STARTING point's free energy), and the default ``recalc_func =
p.hessian`` (``:118``) does not exist on ``Project`` - the Hessian here is ``J^T J`` of the scaled Jacobian at the starting
point (the quadratic approximation the sampling matrix is built for).
"""
import numpy as np

from .. import _lib


def autocorrelation(series):
    """Normalised autocorrelation of a series through the FFT (reference ``Ensembles.py:14-29``, same calls: FFTPACK's
    packed real transform, whose real and imaginary parts the reference squares SEPARATELY before transforming back -
    kept, so that the numbers are the reference's)."""
    import scipy.fftpack
    x = np.asarray(series, dtype=np.float64)
    n = len(x)
    f = scipy.fftpack.rfft(x - np.mean(x), n=2 * n)
    ac = scipy.fftpack.irfft(abs(f) ** 2)[:n]
    ac *= n / (n - 1.0 * np.arange(n))
    return ac / ac[0]


def sampling_matrix(hessian, cutoff=0, temperature=1, step_scale=1):
    """Reference ``_sampling_matrix`` (``Ensembles.py:226-264``): "square root" of the inverse of ``0.5 * hessian`` with
    singular values below ``cutoff * max`` raised to it, scaled so that a move raises the quadratic cost by about 1."""
    u, sing_vals, vh = np.linalg.svd(0.5 * np.asarray(hessian, dtype=np.float64))
    cutoff_sing_val = cutoff * np.max(sing_vals)
    D = 1.0 / np.maximum(sing_vals, cutoff_sing_val)
    samp_mat = vh.T * np.sqrt(D)
    cutoff_vals = sing_vals[sing_vals < cutoff_sing_val]
    if len(cutoff_vals):
        scale = np.sqrt(len(sing_vals) - len(cutoff_vals) + np.sum(cutoff_vals) / cutoff_sing_val)
    else:
        scale = np.sqrt(len(sing_vals))
    return samp_mat / scale * step_scale * np.sqrt(temperature)


def _entropy_device(proj, sims, temperature):
    """Scale-factor entropy of every walker (reference ``base_project.py:854-872`` ->
    ``abstract_loss_function.py:122-134`` -> ``linear_scale_factor.py:13-18, 63-80``) on the device: the SloppyCell integral
    over the log scale factor by composite 16-point Gauss-Legendre panels, the same rule as the host version
    (``LinearScaleFactor.calc_scale_factor_entropy``) with a fixed number of panels.  Zero when no group has a prior-less
    scale factor to integrate (no scale-factor groups)."""
    import torch
    lf, t = proj._loss_function, proj._plan_tables
    W = sims.shape[0]
    ent = torch.zeros((W,), dtype=torch.float64, device=sims.device)
    x, w = np.polynomial.legendre.leggauss(16)
    xg = torch.from_numpy(x).to(sims.device)
    wg = torch.from_numpy(w).to(sims.device)
    for g, key in enumerate(lf.group_keys()):
        holder = lf._scale_factors[lf.group_members(key)[0]]
        mu, sg = holder.log_prior, holder.log_prior_sigma
        if mu is None or sg is None:
            raise TypeError('scale-factor entropy needs a log prior (as in the reference)')
        rows = torch.from_numpy(np.asarray(t.grp_rows[t.grp_off[g]:t.grp_off[g + 1]], dtype=np.int64)).to(sims.device)
        d = torch.from_numpy(np.asarray(t.data)[rows.cpu().numpy()]).to(sims.device)
        sgm = torch.from_numpy(np.asarray(t.sigma)[rows.cpu().numpy()]).to(sims.device)
        s = sims[:, rows]
        sds = (s * s / sgm ** 2).sum(dim=1)
        sde = (s * d / sgm ** 2).sum(dim=1)
        Bf = sde / sds
        lB = torch.log(Bf)
        w_like = torch.sqrt(temperature / (sds * Bf * Bf))
        wide = torch.clamp(w_like, min=sg)
        zero = torch.zeros_like(lB)
        lo = torch.minimum(zero, mu - lB) - 12.0 * wide
        hi = torch.maximum(zero, mu - lB) + 12.0 * wide
        # panels no wider than the narrower of the two peaks (likelihood, prior), as the host rule; one panel count for
        # the whole batch, walkers in chunks so that the [chunk, panels, 16] node array stays small
        narrow = torch.clamp(torch.minimum(w_like, torch.full_like(w_like, sg)), min=1e-9)
        ratio = ((hi - lo) / narrow)
        ratio = ratio[torch.isfinite(ratio)]
        NPAN = int(min(max(float(ratio.max().ceil()) if ratio.numel() else 8.0, 8.0), 4096.0))
        frac = torch.linspace(0, 1, NPAN + 1, dtype=torch.float64, device=sims.device)
        chunk = max(1, int(4e6 // (NPAN * 16)))
        for c0 in range(0, W, chunk):
            sl = slice(c0, min(c0 + chunk, W))
            edges = lo[sl, None] + (hi - lo)[sl, None] * frac[None, :]
            half = 0.5 * (edges[:, 1:] - edges[:, :-1])
            mid = 0.5 * (edges[:, 1:] + edges[:, :-1])
            u = mid[:, :, None] + half[:, :, None] * xg[None, None, :]
            f = torch.exp(-sds[sl, None, None] / (2 * temperature) * (Bf[sl, None, None] * torch.exp(u) - Bf[sl, None, None]) ** 2
                          - (u + lB[sl, None, None] - mu) ** 2 / (2 * sg ** 2))
            ent[sl] = ent[sl] + temperature * torch.log((half[:, :, None] * wg[None, None, :] * f).sum(dim=(1, 2)))
    return ent


def gauss_newton_hessian(p, theta):
    """Gauss-Newton Hessian of the cost ``0.5 sum r^2`` at ``theta``: ``(J / sigma)^T (J / sigma)``.  The project's scaled
    Jacobian is NOT divided by sigma (reference ``squared_loss_function.py:76``), the residuals are: the rows
    are divided here.  (The reference's default ``recalc_func = p.hessian``, ``Ensembles.py:118``, does not exist.)"""
    out = p.jacobian_batch(np.asarray(theta, dtype=np.float64)[None, :])
    J = out['jac'][0] / np.asarray(p.plan_tables.sigma)[:, None]
    return J.T.dot(J)


def _sf_prior_arrays(p, dev):
    """(mu [G], sigma [G]) device tensors of the scale factors' log priors (NaN where a group has none), or (None, None)
    when no group has a prior: the free energy is then the plain SSR (reference base_project.py:874-892 with an empty sum)."""
    import torch
    t = p._plan_tables
    if t.n_groups == 0:
        return None, None
    lf = p._loss_function
    mu, sg = [], []
    for key in lf.group_keys():
        holder = lf._scale_factors[lf.group_members(key)[0]]
        has = holder.log_prior is not None and holder.log_prior_sigma is not None
        mu.append(float(holder.log_prior) if has else float('nan'))
        sg.append(float(holder.log_prior_sigma) if has else float('nan'))
    if all(np.isnan(x) for x in mu):
        return None, None
    return (torch.tensor(mu, dtype=torch.float64, device=dev), torch.tensor(sg, dtype=torch.float64, device=dev))


def ensemble_log_params_batch(p, params, hess=None, steps=100, temperature=1.0, step_scale=1.0, sing_val_cutoff=0,
                              seeds=None, skip_elems=0, n_walkers=None, normals=None, uniforms=None, solver=None):
    """W Metropolis walkers in the logarithm of the parameters, in lock-step on the GPU.

    Every Monte-Carlo step is four launches and no host synchronisation: ``sbm_mc_propose`` (proposal ``cur + samp_mat . z``,
    Philox random numbers generated in the kernel), the integrator and the loss epilogue on the W candidates (the hot
    path), ``sbm_mc_accept`` (free energy incl. the scale-factor entropy quadrature, Metropolis rule, bookkeeping, recording).
    Walkers, free energies and the recorded ensemble stay in HBM; one copy to the host when the run ends.

    Parameters (reference names, ``Ensembles.py:55-63``; additive: ``n_walkers``, ``normals``, ``uniforms``)
    ----------
    p : Project (CUDA backend)
    params : f8[n_proj] (every walker starts there) or f8[W, n_proj]
    hess : f8[n_proj, n_proj], optional - default the Gauss-Newton Hessian at the (first) starting point (``gauss_newton_hessian``)
    steps : Monte-Carlo steps of every walker
    temperature, step_scale, sing_val_cutoff, skip_elems : as in the reference
    seeds : int - seed of the device random generator
    normals f8[steps, W, n_proj], uniforms f8[steps, W] : optional host random numbers instead of the device generator
        (tests: the CPU restatement consumes the same numbers)

    Returns ``(ens f8[n_kept, W, n_proj], ens_fes f8[n_kept, W], ratio f8[W])``: the walkers' positions and free energies
    (starting point first, then every ``skip_elems + 1``-th step) and each walker's acceptance ratio.
    """
    import ctypes
    import torch
    _lib.require_gpu()
    L = _lib.lib()
    vp = ctypes.c_void_p
    dev = torch.device('cuda')
    params = np.asarray(params, dtype=np.float64)
    if params.ndim == 1:
        params = np.tile(params[None, :], (int(n_walkers or 1), 1))
    W, P = params.shape
    if hess is None:
        hess = gauss_newton_hessian(p, params[0])
    f64 = dict(dtype=torch.float64, device=dev)
    samp = torch.from_numpy(np.ascontiguousarray(sampling_matrix(hess, sing_val_cutoff, temperature, step_scale))).to(dev)
    seed = int(seeds) if seeds is not None else 0
    sf_mu, sf_sig = _sf_prior_arrays(p, dev)
    plan = p.device_plan
    st = torch.cuda.current_stream().cuda_stream
    steps, every = int(steps), int(skip_elems) + 1
    n_kept = 1 + steps // every
    cur = torch.from_numpy(np.ascontiguousarray(params)).to(dev)
    nxt = torch.empty_like(cur)
    curF = torch.empty((W,), **f64)
    accepted = torch.zeros((W,), **f64)
    ens = torch.empty((n_kept, W, P), **f64)
    fes = torch.empty((n_kept, W), **f64)
    if normals is not None:
        normals_d = torch.from_numpy(np.ascontiguousarray(normals, dtype=np.float64)).to(dev)
        uniforms_d = torch.from_numpy(np.ascontiguousarray(uniforms, dtype=np.float64)).to(dev)
    bufs = {}

    def ptr(x):
        return vp(x.data_ptr()) if x is not None else None

    def accept(theta_eval, init, it, slot):
        o = p.evaluate_device(theta_eval, with_jacobian=False, solver=solver, out=bufs)
        _lib.check(L.sbm_mc_accept(plan.handle, W, int(init), float(temperature), ptr(cur), ptr(nxt), ptr(curF),
                                   ptr(o['ssr']), ptr(o['sims']), ptr(sf_mu), ptr(sf_sig),
                                   ptr(uniforms_d[it]) if (normals is not None and not init) else None, seed, it,
                                   ptr(accepted), ptr(ens[slot]) if slot is not None else None,
                                   ptr(fes[slot]) if slot is not None else None, vp(st)))
    accept(cur, True, 0, 0)
    kept = 1
    for it in range(steps):
        _lib.check(L.sbm_mc_propose(W, P, ptr(samp), ptr(cur), ptr(nxt),
                                    ptr(normals_d[it]) if normals is not None else None, seed, it, vp(st)))
        rec = (it + 1) % every == 0
        accept(nxt, False, it, kept if rec else None)
        kept += 1 if rec else 0
    return ens.cpu().numpy(), fes.cpu().numpy(), (accepted / max(steps, 1)).cpu().numpy()
