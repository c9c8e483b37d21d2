"""Multi-GPU evaluation over the ranks of one box (one process per GPU, ``torch.distributed``; NCCL on GPUs, gloo in
the CPU tests).  Two shards, as SURVEY.md section 8e lays them out:

``ShardedProject`` - PRIMARY shard, by parameter set.  Every (parameter set, experiment) initial-value problem is
    independent and all experiments of a parameter set stay on one rank, so scale factors, residuals, SSR and Jacobian
    rows of a set are complete locally.  No collective inside the integration; afterwards one all-gather of what the
    consumer needs: the scalar SSR (8 B per set), SSR + J^T r, the per-set normal equations J^T J | J^T r | r^T r
    (22 x 22 + 22 + 1 packed = 277 doubles per set on C4 instead of the 40 x 22 Jacobian block), or the blocks themselves.

``ExperimentShardedProject`` - SECONDARY shard, by experiment (few parameter sets, many experiments: C2's 4,096
    experiments).  Every rank integrates a contiguous block of experiments (its slice of the project's flat plan).  A scale factor couples the rows of its
    group across ALL experiments (reference ``abstract_loss_function.py:62-91``: the group's rows are gathered over every
    experiment; ``linear_scale_factor.py:27-42``: ``B = sum(s d / sigma^2) / sum(s^2 / sigma^2)`` and its gradient from
    ``J^T d / sigma^2`` and ``J^T s / sigma^2``), so the loss epilogue runs in two halves with ONE all-reduce of the
    group sums ``[p1, p2, p3, j1[n_proj], j2[n_proj]]`` in between (C ABI ``sbm_loss_sums`` / ``sbm_loss_finish``), then
    an all-reduce of SSR and ``J^T r``.  Jacobian / residual rows stay partitioned by experiment (the reference's row
    order is experiment-major, ``base_project.py:307-314, 624``), gathered on request.

The CPU tests drive the same classes over gloo with injected evaluators (there is no CPU compute path in the product).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced block [lo, hi) of ``n`` items for ``rank`` (first n % world blocks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group), dist.get_rank(group)
    return 1, 0


def gather_rows(local, n_total, world, group=None):
    """all-gather variable-sized row blocks (``shard_bounds`` sizes: they differ by at most one row, and a rank may hold
    none): pad to the largest block, one ``all_gather_into_tensor``, strip the padding."""
    if world == 1:
        return local
    per = max(-(-n_total // world), 1)
    padded = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[:local.shape[0]] = local
    out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, padded, group=group)
    if n_total == world * per:
        return out
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, world, r)
        pieces.append(out[r * per:r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


class ShardedProject(object):
    """Primary shard: contiguous blocks of parameter sets per rank."""

    def __init__(self, project, group=None, local_eval=None, local_normal=None):
        self.project = project
        self.group = group
        self.world, self.rank = _world(group)
        self._local_eval = local_eval or self._cuda_eval
        self._local_normal = local_normal or self._cuda_normal
        self._bufs = {}

    # ---- local evaluation (CUDA) --------------------------------------------------------------------------
    def _cuda_eval(self, thetas_local, with_jacobian):
        return self.project.evaluate_device(thetas_local, with_jacobian=with_jacobian, out=self._bufs)

    def _cuda_normal(self, out):
        from . import engine
        packed = self._bufs.get('normal')
        B, R, P = out['jac'].shape
        n = P * (P + 1) // 2 + P + 1
        if packed is None or tuple(packed.shape) != (B, n):
            packed = torch.empty((B, n), dtype=torch.float64, device=out['jac'].device)
            self._bufs['normal'] = packed
        return engine.normal_equations_device(out['jac'], out['resid'], out=packed)

    def _empty_like_eval(self, thetas_local, with_jacobian):
        """Correctly shaped zero-row results for a rank whose block is empty (fewer parameter sets than ranks): the rank
        must still enter every collective."""
        t = self.project.plan_tables
        f64 = dict(dtype=torch.float64, device=thetas_local.device)
        out = {'ssr': torch.zeros((0,), **f64), 'resid': torch.zeros((0, t.n_rows), **f64),
               'sf': torch.zeros((0, max(t.n_groups, 1)), **f64)}
        if with_jacobian:
            out['jac'] = torch.zeros((0, t.n_rows, t.n_proj), **f64)
            out['grad'] = torch.zeros((0, t.n_proj), **f64)
        return out

    def local_block(self, thetas_global):
        lo, hi = shard_bounds(thetas_global.shape[0], self.world, self.rank)
        return lo, hi, thetas_global[lo:hi]

    def evaluate_local(self, thetas_local, with_jacobian=True):
        if thetas_local.shape[0] == 0:
            return self._empty_like_eval(thetas_local, with_jacobian)
        return self._local_eval(thetas_local, with_jacobian)

    # ---- global views: ``thetas_global`` is the same [B, n_proj] tensor on every rank ------------------------
    def ssr(self, thetas_global):
        """SSR of every parameter set, on every rank."""
        lo, hi, th = self.local_block(thetas_global)
        out = self.evaluate_local(th, False)
        return gather_rows(out['ssr'].reshape(-1, 1), thetas_global.shape[0], self.world, self.group).reshape(-1)

    def ssr_and_gradient(self, thetas_global):
        """(SSR [B], J^T r [B, n_proj]) on every rank; Jacobian rows stay on the rank that computed them."""
        lo, hi, th = self.local_block(thetas_global)
        out = self.evaluate_local(th, True)
        packed = torch.cat([out['ssr'].reshape(-1, 1), out['grad']], dim=1)
        allp = gather_rows(packed, thetas_global.shape[0], self.world, self.group)
        return allp[:, 0].contiguous(), allp[:, 1:].contiguous()

    def normal_equations(self, thetas_global):
        """Packed per-set normal equations [B, P (P + 1) / 2 + P + 1] (lower triangle of J^T J | J^T r | r^T r) of every
        parameter set on every rank - the Jacobian-block gather in the form the optimisers consume."""
        lo, hi, th = self.local_block(thetas_global)
        P = self.project.plan_tables.n_proj
        n = P * (P + 1) // 2 + P + 1
        if th.shape[0] == 0:
            local = torch.zeros((0, n), dtype=torch.float64, device=th.device)
        else:
            local = self._local_normal(self.evaluate_local(th, True))
        return gather_rows(local, thetas_global.shape[0], self.world, self.group)

    def jacobian(self, thetas_global, gather=False):
        """Scaled Jacobian blocks.  ``gather=False``: (lo, hi, J_local [hi-lo, n_rows, n_proj]);
        ``gather=True``: the full [B, n_rows, n_proj] tensor on every rank (one all-gather)."""
        lo, hi, th = self.local_block(thetas_global)
        out = self.evaluate_local(th, True)
        if not gather:
            return lo, hi, out['jac']
        return gather_rows(out['jac'], thetas_global.shape[0], self.world, self.group)

    # ---- weak-scaling form: every rank brings its OWN block (what bench.py times) --------------------------
    def step_local_blocks(self, thetas_local, out_all=None):
        """One objective evaluation of this rank's own block + the exchange an optimiser step needs: all-gather of the
        packed normal equations (which carry r^T r = 2 SSR).  Blocks must have the same size on every rank.
        Returns (local outputs, gathered [world * B_local, n] tensor)."""
        out = self.evaluate_local(thetas_local, True)
        packed = self._local_normal(out)
        if self.world == 1:
            return out, packed
        if out_all is None or out_all.shape[0] != self.world * packed.shape[0]:
            out_all = torch.empty((self.world * packed.shape[0], packed.shape[1]), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out_all, packed, group=self.group)
        return out, out_all


def slice_plan_tables(t, e_lo, e_hi):
    """The plan of experiments [e_lo, e_hi) of a project, with the project's GLOBAL column space: rows are renumbered
    from 0 (measurement rows are experiment-major and contiguous, base_project.py:307-314), every other table is cut
    and rebased.  Prior rows belong to no experiment and are not supported by the experiment shard."""
    from . import engine, _lib
    if np.any(t.row_kind != _lib.SBM_ROW_MEASUREMENT):
        raise NotImplementedError('the experiment shard does not support parameter / scale-factor prior rows')
    k_lo, k_hi = int(t.tout_off[e_lo]), int(t.tout_off[e_hi])
    o_lo, o_hi = int(t.orow_off[k_lo]), int(t.orow_off[k_hi])
    if o_hi > o_lo:
        r_lo, r_hi = int(t.row_sim[o_lo:o_hi].min()), int(t.row_sim[o_lo:o_hi].max()) + 1
    else:
        r_lo = r_hi = 0
    keep = (t.grp_rows >= r_lo) & (t.grp_rows < r_hi)
    grp_off = [0]
    for g in range(t.n_groups):
        grp_off.append(grp_off[-1] + int(keep[t.grp_off[g]:t.grp_off[g + 1]].sum()))
    return engine.PlanTables(
        E=e_hi - e_lo, n_rows=r_hi - r_lo, n_proj=t.n_proj, n_groups=t.n_groups, jac_atomic=t.jac_atomic,
        pmap=t.pmap.reshape(t.E, -1)[e_lo:e_hi], pfix=t.pfix.reshape(t.E, -1)[e_lo:e_hi], y0=t.y0.reshape(t.E, -1)[e_lo:e_hi],
        tout_off=t.tout_off[e_lo:e_hi + 1] - k_lo, tout=t.tout[k_lo:k_hi], orow_off=t.orow_off[k_lo:k_hi + 1] - o_lo,
        row_sim=t.row_sim[o_lo:o_hi] - r_lo, row_jac=t.row_jac[o_lo:o_hi] - r_lo, row_mask=t.row_mask[o_lo:o_hi],
        data=t.data[r_lo:r_hi], sigma=t.sigma[r_lo:r_hi], row_group=t.row_group[r_lo:r_hi], row_kind=t.row_kind[r_lo:r_hi],
        row_aux=t.row_aux[r_lo:r_hi], grp_off=grp_off, grp_rows=t.grp_rows[keep] - r_lo), r_lo, r_hi


class ExperimentShardedProject(object):
    """Secondary shard: the experiments of a project split over the ranks (contiguous blocks of the name-sorted list).

    Every rank holds the plan of ITS experiments in the project's global column space (``slice_plan_tables``): a scale
    factor's gradient couples the rows of one experiment to the parameters of all others (``squared_loss_function.py:76``:
    ``sim[:, None] * dB``), so the scaled Jacobian rows of a rank are dense over all project columns.  Parameter sets
    are replicated: every rank evaluates all B sets on its experiments.
    """

    def __init__(self, project, group=None, evaluator=None):
        self.project = project
        self.group = group
        self.world, self.rank = _world(group)
        t = project.plan_tables
        self.n_experiments, self.n_proj, self.n_groups = t.E, t.n_proj, t.n_groups
        self.exp_lo, self.exp_hi = shard_bounds(t.E, self.world, self.rank)
        self.tables, self.row_lo, self.row_hi = slice_plan_tables(t, self.exp_lo, self.exp_hi)
        self.row_counts = []
        for r in range(self.world):
            a, b = shard_bounds(t.E, self.world, r)
            _, lo, hi = slice_plan_tables(t, a, b) if b > a else (None, 0, 0)
            self.row_counts.append(hi - lo)
        self._evaluator = evaluator or _CudaExperimentEvaluator()
        self._plan = None

    @property
    def device_plan(self):
        """Device copy of this rank's plan (None when the rank holds no experiment)."""
        from . import engine
        if self._plan is None and self.exp_hi > self.exp_lo:
            self._plan = engine.DevicePlan(self.project._model.compiled, self.tables)
        return self._plan

    def evaluate(self, thetas, with_jacobian=True, gather_rows_too=False):
        """thetas [B, n_proj], identical on every rank.  Returns a dict with, on EVERY rank, ``ssr`` [B], ``sf`` [B, G],
        ``grad`` [B, n_proj] (if ``with_jacobian``), this rank's rows ``resid`` [B, rows_local] / ``jac``
        [B, rows_local, n_proj] (rows ``row_lo:row_hi`` of the project), and with ``gather_rows_too`` the full
        ``resid_all`` [B, n_rows]."""
        B, dev = thetas.shape[0], thetas.device
        f64 = dict(dtype=torch.float64, device=dev)
        G, P = self.n_groups, self.n_proj
        ev = self._evaluator
        have = self.exp_hi > self.exp_lo
        # ---- half 1: integrate the local experiments, group sums of the local rows ------------------------------------
        sums = torch.zeros((B, max(G, 1), 3 + 2 * P), **f64)
        state = None
        if have:
            state, sums_l = ev.solve_and_sums(self, thetas, with_jacobian)
            if G:
                sums = sums_l
        # ---- the one exchange before residual scaling (abstract_loss_function.py:62-91 gathers over ALL experiments) ----
        if self.world > 1 and G:
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=self.group)
        # ---- half 2: finish the local rows with the global sums; then SSR and J^T r are sums over the ranks' rows -------
        red = torch.zeros((B, 1 + P), **f64)
        out = {}
        if have:
            out = ev.finish(self, state, sums, with_jacobian)
            red[:, 0] = out['ssr']
            if with_jacobian:
                red[:, 1:] = out['grad']
        if self.world > 1:
            dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.group)
        res = {'ssr': red[:, 0].contiguous(), 'resid': out.get('resid'), 'jac': out.get('jac')}
        if with_jacobian:
            res['grad'] = red[:, 1:].contiguous()
        if G:
            # every rank derives the same scale factors from the same global sums
            res['sf'] = out['sf'] if have else None
        if gather_rows_too:
            res['resid_all'] = self._gather_resid(out.get('resid'), B, dev)
        return res

    def _gather_resid(self, resid_local, B, dev):
        if self.world == 1:
            return resid_local
        per = max(max(self.row_counts), 1)
        padded = torch.zeros((per, B), dtype=torch.float64, device=dev)
        if resid_local is not None:
            padded[:resid_local.shape[1]] = resid_local.t()
        allr = torch.empty((self.world * per, B), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allr, padded, group=self.group)
        pieces = [allr[r * per:r * per + self.row_counts[r]] for r in range(self.world)]
        return torch.cat(pieces, dim=0).t().contiguous()


class _CudaExperimentEvaluator(object):
    """The CUDA halves of ``ExperimentShardedProject.evaluate``: integrator + ``sbm_loss_sums``, then ``sbm_loss_finish``."""

    def solve_and_sums(self, esp, theta, with_jacobian):
        from . import _lib
        proj, plan, t = esp.project, esp.device_plan, esp.tables
        B, dev = theta.shape[0], theta.device
        f64 = dict(dtype=torch.float64, device=dev)
        sims = torch.zeros((B, t.n_rows), **f64)
        jac = torch.empty((B, t.n_rows, t.n_proj), **f64) if with_jacobian else None
        status = torch.zeros((B, t.E), dtype=torch.int32, device=dev)
        counters = torch.zeros((_lib.SBM_N_COUNTERS,), dtype=torch.int64, device=dev)
        plan.solve_device(theta, proj._opts(proj._jac_solver(theta) if with_jacobian else proj._concrete_solver(False, theta)), sims, jac, status, counters)
        sums = torch.zeros((B, max(t.n_groups, 1), 3 + 2 * t.n_proj), **f64)
        if t.n_groups:
            plan.loss_sums_device(theta, sims, jac, sums, loss=proj._loss_name(), flags=proj._flags())
        return {'theta': theta, 'sims': sims, 'jac': jac, 'status': status, 'counters': counters}, sums

    def finish(self, esp, state, sums, with_jacobian):
        proj, plan, t = esp.project, esp.device_plan, esp.tables
        theta = state['theta']
        B, dev = theta.shape[0], theta.device
        f64 = dict(dtype=torch.float64, device=dev)
        G = max(t.n_groups, 1)
        resid = torch.empty((B, t.n_rows), **f64)
        ssr = torch.empty((B,), **f64)
        sf = torch.empty((B, G), **f64)
        sfgrad = torch.empty((B, G, t.n_proj), **f64) if with_jacobian else None
        grad = torch.empty((B, t.n_proj), **f64) if with_jacobian else None
        if t.n_groups:
            plan.loss_finish_device(theta, state['sims'], state['jac'], sums.contiguous(), resid, ssr, sf, sfgrad, grad,
                                    loss=proj._loss_name(), flags=proj._flags())
        else:
            plan.loss_device(theta, state['sims'], state['jac'], resid, ssr, sf, sfgrad, grad, loss=proj._loss_name(),
                             flags=proj._flags())
        return {'resid': resid, 'ssr': ssr, 'sf': sf[:, :t.n_groups], 'jac': state['jac'], 'grad': grad,
                'status': state['status'], 'counters': state['counters']}
