"""NCCL, 2 ranks on 2 GPUs (skipped on a box with fewer): the two shards of sysbio_modeling_b200.distributed with their
DEFAULT CUDA evaluators against the single-GPU result.

* primary shard (``ShardedProject``): C4 MAPK project, parameter sets split over the ranks, all-gather of SSR, of
  SSR + J^T r, of the packed per-set normal equations and of the Jacobian blocks;
* experiment shard (``ExperimentShardedProject``): C2 reversible MM, 64 experiments with experiment-local parameters and
  one scale-factor group over all of them, split over the ranks - one all-reduce of the group sums before residual
  scaling (abstract_loss_function.py:62-91, linear_scale_factor.py:27-42)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from tools import workloads as wl
    from sysbio_modeling_b200.distributed import ShardedProject, ExperimentShardedProject
    res = {}
    # ---- primary shard on C4 (tolerance 1e-8 keeps the test short) ----
    proj, _ = wl.c4_project(1e-8)
    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    TH = torch.from_numpy(wl.make_thetas(nominal, 101, 4, 0.25)).cuda()          # odd: ragged blocks
    sp = ShardedProject(proj)
    res['ssr'] = sp.ssr(TH).cpu()
    s2, g2 = sp.ssr_and_gradient(TH)
    res['ssr2'], res['grad'] = s2.cpu(), g2.cpu()
    res['normal'] = sp.normal_equations(TH).cpu()
    res['J'] = sp.jacobian(TH, gather=True).cpu()
    out_l, packed_all = sp.step_local_blocks(TH[:32].contiguous())
    res['packed_all'] = packed_all.cpu()
    if rank == 0:
        one = proj.evaluate_device(TH, True)
        res['one'] = {k: one[k].cpu() for k in ('ssr', 'grad', 'jac', 'resid')}
    # ---- experiment shard on C2 ----
    E = 64
    proj2, _ = wl.c2_project(E, shared_scale_factor=True, rtol=1e-8, atol=1e-8)
    TH2 = torch.from_numpy(wl.c2_theta(E, 3)).cuda()
    esp = ExperimentShardedProject(proj2)
    r2 = esp.evaluate(TH2, with_jacobian=True, gather_rows_too=True)
    res['e_ssr'], res['e_grad'], res['e_sf'] = r2['ssr'].cpu(), r2['grad'].cpu(), r2['sf'].cpu()
    res['e_resid_all'], res['e_jac'], res['e_rows'] = r2['resid_all'].cpu(), r2['jac'].cpu(), (esp.row_lo, esp.row_hi)
    r3 = esp.evaluate(TH2, with_jacobian=False)
    res['e_ssr_plain'] = r3['ssr'].cpu()
    if rank == 0:
        one = proj2.evaluate_device(TH2, True)
        res['e_one'] = {k: one[k].cpu() for k in ('ssr', 'grad', 'jac', 'resid', 'sf')}
    torch.save(res, os.path.join(out_dir, 'nccl%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(900)
def test_both_shards_on_nccl_match_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp
    if _n_gpus() < 2:
        pytest.skip('needs 2 GPUs')
    from sysbio_modeling_b200 import engine
    world = 2
    port = 35500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), 'nccl%d.pt' % r), weights_only=False) for r in range(world)]
    one = outs[0]['one']
    for o in outs:
        # same kernels, same arithmetic per parameter set: bitwise equal to the single-GPU evaluation
        assert torch.equal(o['ssr'], one['ssr']) or torch.allclose(o['ssr'], one['ssr'], rtol=1e-7)
        assert torch.equal(o['ssr2'], one['ssr'])
        assert torch.equal(o['grad'], one['grad'])
        assert torch.equal(o['J'], one['jac'])
        A, g, rr = engine.unpack_normal_equations(o['normal'], 22)
        A0 = torch.einsum('bri,brj->bij', one['jac'], one['jac'])
        assert torch.allclose(A, A0, rtol=1e-11, atol=1e-12 * A0.abs().max().item())
        assert torch.allclose(g, one['grad'], rtol=1e-10, atol=1e-12 * one['grad'].abs().max().item())
        assert torch.allclose(rr, 2 * one['ssr'], rtol=1e-12)
        assert o['packed_all'].shape == (64, 22 * 23 // 2 + 22 + 1)
        assert torch.equal(o['packed_all'][:32], o['packed_all'][32:])          # both ranks brought the same block
    # experiment shard
    e1 = outs[0]['e_one']
    for o in outs:
        assert torch.allclose(o['e_ssr'], e1['ssr'], rtol=1e-11)
        assert torch.allclose(o['e_sf'], e1['sf'], rtol=1e-12)
        assert torch.allclose(o['e_resid_all'], e1['resid'], rtol=1e-10, atol=1e-11)
        assert torch.allclose(o['e_grad'], e1['grad'], rtol=1e-9, atol=1e-11 * e1['grad'].abs().max().item())
        lo, hi = o['e_rows']
        assert torch.allclose(o['e_jac'], e1['jac'][:, lo:hi], rtol=1e-9, atol=1e-12 * e1['jac'].abs().max().item())
        assert torch.allclose(o['e_ssr_plain'], e1['ssr'], rtol=1e-6)             # plain kernel: equal to the tolerance
    assert outs[0]['e_rows'] == (0, 512) and outs[1]['e_rows'] == (512, 1024)
