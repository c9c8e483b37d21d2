"""Multi-GPU evaluation: parameter sets shard across the ranks of one box (one process per GPU).

Every (parameter set, experiment) initial-value problem is independent and all experiments of a
parameter set stay on one rank, so scale factors, residuals, SSR and Jacobian rows of a parameter set
are complete locally.  The only exchange is a gather of the per-set scalars
(SSR: 8 B per set) and, on request, of the gradient / Jacobian blocks - one NCCL all-gather on the
compute stream after the loss epilogue.  There is no data-path collective inside the integration.

``ShardedProject`` is backend-agnostic plumbing over ``torch.distributed``: with the default evaluator
it runs ``Project.evaluate_device`` on CUDA tensors over NCCL; the CPU tests drive the same code over
gloo with an injected evaluator.
"""
import torch
import torch.distributed as dist


def shard_bounds(n, world, rank):
    """Contiguous, balanced block [lo, hi) of ``n`` items for ``rank`` (first n % world blocks get one more)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class ShardedProject(object):
    def __init__(self, project, group=None, local_eval=None):
        self.project = project
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self._local_eval = local_eval or self._cuda_eval
        self._bufs = {}

    def _cuda_eval(self, thetas_local, with_jacobian):
        return self.project.evaluate_device(thetas_local, with_jacobian=with_jacobian, out=self._bufs)

    def _gather_rows(self, local, n_total):
        """all-gather variable-sized row blocks (sizes differ by at most one row): pad to the largest block."""
        if self.world == 1:
            return local
        per = -(-n_total // self.world)
        pad_shape = (per,) + tuple(local.shape[1:])
        padded = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
        padded[:local.shape[0]] = local
        out = torch.empty((self.world * per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, padded, group=self.group)
        pieces = []
        for r in range(self.world):
            lo, hi = shard_bounds(n_total, self.world, r)
            pieces.append(out[r * per:r * per + (hi - lo)])
        return torch.cat(pieces, dim=0)

    def local_block(self, thetas_global):
        lo, hi = shard_bounds(thetas_global.shape[0], self.world, self.rank)
        return lo, hi, thetas_global[lo:hi]

    def ssr(self, thetas_global):
        """SSR of every parameter set, on every rank."""
        lo, hi, th = self.local_block(thetas_global)
        out = self._local_eval(th, False)
        return self._gather_rows(out['ssr'].reshape(-1, 1), thetas_global.shape[0]).reshape(-1)

    def ssr_and_gradient(self, thetas_global):
        """(SSR [B], J^T r [B, n_proj]) on every rank; Jacobian rows stay on the rank that computed them."""
        lo, hi, th = self.local_block(thetas_global)
        out = self._local_eval(th, True)
        packed = torch.cat([out['ssr'].reshape(-1, 1), out['grad']], dim=1)
        allp = self._gather_rows(packed, thetas_global.shape[0])
        return allp[:, 0].contiguous(), allp[:, 1:].contiguous()

    def jacobian(self, thetas_global, gather=False):
        """Scaled Jacobian blocks.  ``gather=False``: (lo, hi, J_local [hi-lo, n_rows, n_proj]);
        ``gather=True``: the full [B, n_rows, n_proj] tensor on every rank (one all-gather)."""
        lo, hi, th = self.local_block(thetas_global)
        out = self._local_eval(th, True)
        if not gather:
            return lo, hi, out['jac']
        return self._gather_rows(out['jac'], thetas_global.shape[0])
