"""BASELINE.json's full sizes, checked through properties that do not need the CPU oracle (which takes 17 ms per solve):

  * a trajectory's arithmetic does not depend on where it sits in the batch: the results of a permuted batch are the
    permuted results, bit for bit, and the first sets evaluated alone equal the first sets of the batch
    (the persistent kernel hands trajectories to lanes through a global queue - which lanes share a CTA changes);
  * SSR = r.r / 2 and the gradient the epilogue returns is J^T r formed from its own outputs;
  * every trajectory ends with status 0 and finite rows.

The oracle comparison proper runs on 8 sets of the same projects (tests/test_gpu_bench_parity.py, bench.py `parity`)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _clone(o, keys):
    return {k: o[k].clone() for k in keys if k in o and o[k] is not None}


def test_c4_full_batch_permutation_subset_and_identities():
    import torch
    from tools import workloads as wl
    B = 65536
    proj, _ = wl.c4_project()                                   # rtol = atol = 1e-10: the bench project
    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    theta = torch.from_numpy(wl.make_thetas(nominal, B, 4, 0.25)).cuda()
    keys = ('sims', 'resid', 'ssr', 'sf', 'jac', 'grad', 'status')
    bufs = {}
    a = proj.evaluate_device(theta, True, out=bufs)
    torch.cuda.synchronize()
    assert int((a['status'] != 0).sum()) == 0
    assert bool(torch.isfinite(a['jac']).all()) and bool(torch.isfinite(a['resid']).all())
    # identities of the epilogue, in chunks (the einsum of the whole batch would double the Jacobian's 461 MB)
    for lo in range(0, B, 8192):
        J, r = a['jac'][lo:lo + 8192], a['resid'][lo:lo + 8192]
        g = torch.einsum('brc,br->bc', J, r)
        assert torch.allclose(a['grad'][lo:lo + 8192], g, rtol=1e-9, atol=1e-9 * float(g.abs().max()))
        assert torch.allclose(a['ssr'][lo:lo + 8192], 0.5 * (r * r).sum(1), rtol=1e-12)
    ck = {k: a[k].double().sum().item() for k in ('sims', 'resid', 'jac')}
    small = _clone(a, ('sims', 'resid', 'ssr', 'sf', 'grad'))
    jac_head = a['jac'][:64].clone()
    # permuted batch
    perm = torch.from_numpy(np.random.default_rng(5).permutation(B)).cuda()
    b = proj.evaluate_device(theta[perm].contiguous(), True, out=bufs)
    torch.cuda.synchronize()
    for k in small:
        assert torch.equal(b[k], small[k][perm]), k
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(B, device='cuda')
    assert torch.equal(b['jac'][inv[:64]], jac_head)
    for k in ck:                                                 # same multiset of numbers (sums in another order: not bitwise)
        assert abs(b[k].double().sum().item() - ck[k]) <= 1e-9 * abs(ck[k]) + 1e-9
    # the first 8 sets alone
    c = proj.evaluate_device(theta[:8].contiguous(), True)
    torch.cuda.synchronize()
    for k in small:
        assert torch.equal(c[k], small[k][:8]), k
    assert torch.equal(c['jac'], jac_head[:8])


def test_c3_full_batch_permutation_and_identities():
    import torch
    from tools import workloads as wl
    B = 1 << 20
    proj, _ = wl.c3_project()
    sm, nominal, y0 = wl.nominal_of('robertson')
    theta = torch.from_numpy(wl.make_thetas(nominal, B, 3, 0.25)).cuda()
    a = _clone(proj.evaluate_device(theta, False), ('sims', 'resid', 'ssr', 'status'))
    torch.cuda.synchronize()
    assert int((a['status'] != 0).sum()) == 0
    assert bool(torch.isfinite(a['resid']).all())
    assert torch.allclose(a['ssr'], 0.5 * (a['resid'] ** 2).sum(1), rtol=1e-12)
    perm = torch.from_numpy(np.random.default_rng(6).permutation(B)).cuda()
    b = proj.evaluate_device(theta[perm].contiguous(), False)
    torch.cuda.synchronize()
    for k in ('sims', 'resid', 'ssr'):
        assert torch.equal(b[k], a[k][perm]), k


def test_c2_full_size_experiment_local_parameters():
    """4,096 experiments with their own four parameters, 2 sets through the Project: block structure of the 65,536 x
    16,384 Jacobian (checked on slices - it is 8.6 GB per set), gradient = J^T r, and the rows of an experiment do not
    depend on the other experiments' parameters."""
    import torch
    from tools import workloads as wl
    E = 4096
    proj, _ = wl.c2_project(E)
    th = torch.from_numpy(wl.c2_theta(E, 2)).cuda()
    o = proj.evaluate_device(th, True)
    torch.cuda.synchronize()
    assert o['jac'].shape == (2, 16 * E, 4 * E)
    assert int((o['status'] != 0).sum()) == 0
    J, r = o['jac'], o['resid']
    for e in (0, 1, 777, E - 1):
        rows = slice(16 * e, 16 * e + 16)
        blk = J[:, rows, 4 * e:4 * e + 4]
        assert torch.count_nonzero(blk) > 0
        assert int(torch.count_nonzero(J[:, rows, :])) == int(torch.count_nonzero(blk))   # nothing outside the block
        g = torch.einsum('brc,br->bc', blk, r[:, rows])
        assert torch.allclose(o['grad'][:, 4 * e:4 * e + 4], g, rtol=1e-10, atol=1e-12)
    assert torch.allclose(o['ssr'], 0.5 * (r * r).sum(1), rtol=1e-12)
    keep = {k: o[k][:, :64].clone() for k in ('sims', 'resid')}
    jblk = J[:, :64, :16].clone()
    del o, J, r
    torch.cuda.empty_cache()
    th2 = th.clone()
    th2[:, 16:] += 0.1                                           # every experiment but the first four
    o2 = proj.evaluate_device(th2, True)
    torch.cuda.synchronize()
    for k in keep:
        assert torch.equal(o2[k][:, :64], keep[k]), k
    assert torch.equal(o2['jac'][:, :64, :16], jblk)
