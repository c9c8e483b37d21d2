"""CPU, world_size 2 over gloo: the parameter-set sharding and the SSR / gradient / Jacobian gathers of
sysbio_modeling_b200.distributed, driven by the CPU oracle as the per-rank evaluator (the product's default
evaluator is the CUDA engine; it is swapped here only because this box has no GPU)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from sysbio_modeling_b200.distributed import shard_bounds


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 64, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from fixtures import project_1state
    from sysbio_modeling_b200.distributed import ShardedProject
    proj, oracle = project_1state()

    def oracle_eval(th, with_jac):
        th = th.numpy()
        out = {'ssr': torch.tensor([oracle.sum_square_residuals(x) for x in th], dtype=torch.float64)}
        if with_jac:
            out['grad'] = torch.tensor(np.array([oracle.rss_gradient(x) for x in th]).reshape(len(th), 3))
            out['jac'] = torch.tensor(np.array([oracle.jacobian(x) for x in th]).reshape(len(th), 35, 3))
        return out
    sp = ShardedProject(proj, local_eval=oracle_eval)
    rng = np.random.default_rng(7)
    TH = torch.tensor(np.log([0.02, 0.015, 0.002]) + 0.2 * rng.standard_normal((B, 3)))
    ssr = sp.ssr(TH)
    ssr2, grad = sp.ssr_and_gradient(TH)
    J = sp.jacobian(TH, gather=True)
    lo, hi, Jl = sp.jacobian(TH)
    torch.save({'ssr': ssr, 'ssr2': ssr2, 'grad': grad, 'J': J, 'lo': lo, 'hi': hi, 'Jl': Jl},
               os.path.join(out_dir, 'rank%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_matches_single_process(tmp_path):
    B, world = 5, 2          # odd on purpose: ragged shards
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    from fixtures import project_1state
    _, oracle = project_1state()
    rng = np.random.default_rng(7)
    TH = np.log([0.02, 0.015, 0.002]) + 0.2 * rng.standard_normal((B, 3))
    ssr = np.array([oracle.sum_square_residuals(x) for x in TH])
    grad = np.array([oracle.rss_gradient(x) for x in TH])
    outs = [torch.load(os.path.join(str(tmp_path), 'rank%d.pt' % r)) for r in range(world)]
    for r, o in enumerate(outs):
        assert np.allclose(o['ssr'].numpy(), ssr, rtol=1e-12)
        assert np.allclose(o['ssr2'].numpy(), ssr, rtol=1e-12)
        assert np.allclose(o['grad'].numpy(), grad, rtol=1e-10, atol=1e-14)
        assert o['J'].shape == (B, 35, 3)
        assert (o['lo'], o['hi']) == shard_bounds(B, world, r)
        assert torch.equal(o['J'][o['lo']:o['hi']], o['Jl'])
    assert torch.equal(outs[0]['J'], outs[1]['J'])
