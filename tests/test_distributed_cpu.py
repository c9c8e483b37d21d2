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


# ---------------------------------------------------------------------------------------------------------------
# fewer parameter sets than ranks (a rank with an empty block must still enter every collective) + the gather of the
# packed normal equations
# ---------------------------------------------------------------------------------------------------------------
def _pack_normal(J, r):
    P = J.shape[1]
    A = J.T.dot(J)
    ii, jj = np.tril_indices(P)
    return np.concatenate([A[ii, jj], J.T.dot(r), [r.dot(r)]])


def _worker_small(rank, world, port, B, out_dir):
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
    calls = []

    def oracle_eval(th, with_jac):
        calls.append(len(th))
        assert len(th) > 0, 'an empty block must not reach the evaluator'
        th = th.numpy()
        out = {'ssr': torch.tensor([oracle.sum_square_residuals(x) for x in th], dtype=torch.float64),
               'resid': torch.tensor(np.array([oracle.residuals(x) for x in th]))}
        if with_jac:
            out['grad'] = torch.tensor(np.array([oracle.rss_gradient(x) for x in th]).reshape(len(th), 3))
            out['jac'] = torch.tensor(np.array([oracle.jacobian(x) for x in th]).reshape(len(th), 35, 3))
        return out

    def oracle_normal(out):
        return torch.tensor(np.array([_pack_normal(J, r) for J, r in zip(out['jac'].numpy(), out['resid'].numpy())]))
    sp = ShardedProject(proj, local_eval=oracle_eval, local_normal=oracle_normal)
    rng = np.random.default_rng(9)
    TH = torch.tensor(np.log([0.02, 0.015, 0.002]) + 0.2 * rng.standard_normal((B, 3)))
    res = {'ssr': sp.ssr(TH), 'normal': sp.normal_equations(TH), 'J': sp.jacobian(TH, gather=True), 'calls': calls}
    torch.save(res, os.path.join(out_dir, 'small%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_fewer_sets_than_ranks_and_normal_equation_gather(tmp_path):
    B, world = 1, 2
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker_small, args=(world, port, B, str(tmp_path)), nprocs=world, join=True)
    from fixtures import project_1state
    _, oracle = project_1state()
    rng = np.random.default_rng(9)
    TH = np.log([0.02, 0.015, 0.002]) + 0.2 * rng.standard_normal((B, 3))
    outs = [torch.load(os.path.join(str(tmp_path), 'small%d.pt' % r)) for r in range(world)]
    expect = np.array([_pack_normal(oracle.jacobian(x), oracle.residuals(x)) for x in TH])
    for o in outs:
        assert o['ssr'].shape == (B,) and o['J'].shape == (B, 35, 3)
        assert np.allclose(o['ssr'].numpy(), [oracle.sum_square_residuals(x) for x in TH], rtol=1e-12)
        assert np.allclose(o['normal'].numpy(), expect, rtol=1e-10, atol=1e-14)
    assert outs[1]['calls'] == []            # rank 1 held no parameter set


# ---------------------------------------------------------------------------------------------------------------
# experiment shard: experiments split over 2 ranks, scale-factor group spanning both, one all-reduce of the group sums
# ---------------------------------------------------------------------------------------------------------------
class OracleExperimentEvaluator(object):
    """CPU stand-in for the CUDA halves (sbm_loss_sums / sbm_loss_finish): numpy restatement of
    linear_scale_factor.py:27-42 and squared_loss_function.py:27-80 on the rows ``row_lo:row_hi`` of the full
    OracleProject (test-only shortcut: the oracle integrates every experiment, the rank's rows are cut out)."""

    def __init__(self, oracle, perm):
        self.o, self.perm = oracle, perm            # perm: oracle column of every project column

    def solve_and_sums(self, esp, theta, with_jacobian):
        o, th = self.o, theta.numpy()
        B, P = th.shape
        lo, hi = esp.row_lo, esp.row_hi
        G = len(o.sf_groups)
        sums = np.zeros((B, max(G, 1), 3 + 2 * P))
        st = []
        inv = np.argsort(self.perm)
        for b in range(B):
            th_o = th[b][inv]
            sims, _ = o.sim_experiments(th_o)
            J = o.model_jacobian(th_o)[:, self.perm] if with_jacobian else None
            for g, key in enumerate(o.sf_groups):
                rows = np.array([r for r in o.group_rows[key] if lo <= r < hi], dtype=np.int64)
                s, d, sg = sims[rows], o.meas_mean[rows], o.meas_std[rows]
                sums[b, g, 0] = np.sum(s * d / sg ** 2)
                sums[b, g, 1] = np.sum(s * s / sg ** 2)
                if with_jacobian:
                    sums[b, g, 3:3 + P] = np.sum(J[rows].T * d / sg ** 2, axis=1)
                    sums[b, g, 3 + P:] = np.sum(J[rows].T * s / sg ** 2, axis=1)
            st.append((sims, J))
        return {'theta': theta, 'st': st}, torch.tensor(sums)

    def finish(self, esp, state, sums, with_jacobian):
        o, th = self.o, state['theta'].numpy()
        B, P = th.shape
        sums = sums.numpy()
        lo, hi = esp.row_lo, esp.row_hi
        R = hi - lo
        resid, ssr, sf = np.zeros((B, R)), np.zeros(B), np.zeros((B, len(o.sf_groups)))
        jac, grad = np.zeros((B, R, P)), np.zeros((B, P))
        for b in range(B):
            sims, J = state['st'][b]
            scaled = sims.copy()
            Js = None if J is None else J.copy()
            for g, key in enumerate(o.sf_groups):
                rows = o.group_rows[key]
                p1, p2 = sums[b, g, 0], sums[b, g, 1]
                Bf = p1 / p2
                sf[b, g] = Bf
                scaled[rows] *= Bf
                if with_jacobian:
                    gB = sums[b, g, 3:3 + P] / p2 - 2 * p1 * sums[b, g, 3 + P:] / p2 ** 2
                    Js[rows] = J[rows] * Bf + sims[rows][:, None] * gB
            resid[b] = ((scaled - o.meas_mean) / o.meas_std)[lo:hi]
            ssr[b] = 0.5 * np.sum(resid[b] ** 2)
            if with_jacobian:
                jac[b] = Js[lo:hi]
                grad[b] = (jac[b].T * resid[b]).sum(axis=1)
        return {'resid': torch.tensor(resid), 'ssr': torch.tensor(ssr), 'sf': torch.tensor(sf),
                'jac': torch.tensor(jac), 'grad': torch.tensor(grad)}


def _oracle_perm(full, columns):
    onames = [None] * full.n_project_params
    for g, d in full.project_param_idx.items():
        for st, i in d.items():
            onames[i] = g + ' ' + ''.join(str(x) for x in st)
    return np.array([onames.index(nm) for nm in columns], dtype=np.int64)


def _c2_small(E):
    """C2's shape at test size: E experiments with experiment-local parameters and ONE scale-factor group over all of
    them (data differ per experiment)."""
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('reversible_mm')
    r = np.random.default_rng(22)
    t = np.linspace(0.25, 4, 6)
    exps = []
    for k in range(E):
        data = 0.5 + 0.4 * r.random(len(t))
        exps.append(sb.Experiment('exp%03d' % k, sb.TimecourseMeasurement('P', data, t, 0.05 * data + 1e-3),
                                  initial_conditions=mod.initial_conditions))
    return exps


def _worker_exp(rank, world, port, out_dir):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, 'oracle'), os.path.join(root, 'tests')):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import sysbio_modeling_b200 as sb
    import sbm_oracle
    from sysbio_modeling_b200 import model_files
    from sysbio_modeling_b200.distributed import ExperimentShardedProject
    from fixtures import python_rhs
    sm, m, s = python_rhs('reversible_mm')
    settings = {'Local': list(sm.param_order)}
    om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
    E = 5                                                    # odd on purpose: 3 + 2 experiments
    proj = sb.Project(om, _c2_small(E), dict(settings), {'P': ('direct', 3)}, sf_groups=['P'])
    full = sbm_oracle.OracleProject(m, s, 4, sm.param_order, _c2_small(E), dict(settings), {'P': ('direct', 3)},
                                    sf_groups=['P'])
    columns = proj.get_ordered_project_params()
    perm = _oracle_perm(full, columns)
    esp = ExperimentShardedProject(proj, evaluator=OracleExperimentEvaluator(full, perm))
    assert (esp.exp_lo, esp.exp_hi) == ((0, 3) if rank == 0 else (3, 5))
    assert esp.tables.n_rows == 6 * (esp.exp_hi - esp.exp_lo) and esp.tables.n_proj == 4 * E
    rng = np.random.default_rng(3)
    nominal = np.array([10.0, 1.0, 5.0, 0.1])
    TH = np.zeros((2, len(columns)))
    for j, nm in enumerate(columns):
        TH[:, j] = np.log(nominal[list(sm.param_order).index(nm.split('_exp')[0])]) + 0.3 * rng.standard_normal(2)
    res = esp.evaluate(torch.tensor(TH), with_jacobian=True, gather_rows_too=True)
    torch.save({'ssr': res['ssr'], 'grad': res['grad'], 'sf': res['sf'], 'resid_all': res['resid_all'], 'TH': torch.tensor(TH),
                'columns': columns, 'jac': res['jac'], 'rows': (esp.row_lo, esp.row_hi)}, os.path.join(out_dir, 'exp%d.pt' % rank))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_experiment_shard_two_ranks_equals_single_process(tmp_path):
    import sbm_oracle
    from fixtures import python_rhs
    world = 2
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_worker_exp, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    outs = [torch.load(os.path.join(str(tmp_path), 'exp%d.pt' % r), weights_only=False) for r in range(world)]
    sm, m, s = python_rhs('reversible_mm')
    settings = {'Local': list(sm.param_order)}
    full = sbm_oracle.OracleProject(m, s, 4, sm.param_order, _c2_small(5), dict(settings), {'P': ('direct', 3)},
                                    sf_groups=['P'])
    TH = outs[0]['TH'].numpy()
    perm = _oracle_perm(full, outs[0]['columns'])
    inv = np.argsort(perm)
    for b in range(2):
        th_o = TH[b][inv]
        r0 = full.residuals(th_o)
        J0 = full.jacobian(th_o)[:, perm]                    # project column order
        g0 = (J0.T * r0).sum(axis=1)
        for o in outs:
            assert abs(o['ssr'][b].item() - 0.5 * np.sum(r0 ** 2)) <= 1e-10 * np.sum(r0 ** 2)
            assert np.allclose(o['resid_all'][b].numpy(), r0, rtol=1e-9, atol=1e-9)
            assert np.allclose(o['grad'][b].numpy(), g0, rtol=1e-8, atol=1e-8 * np.abs(g0).max())
            assert abs(o['sf'][b, 0].item() - full.sf['P']) <= 1e-10 * abs(full.sf['P'])
            lo, hi = o['rows']
            # the rank's rows of the scaled Jacobian are dense over ALL project columns (sim * dB couples experiments)
            assert np.allclose(o['jac'][b].numpy(), J0[lo:hi], rtol=1e-8, atol=1e-9 * np.abs(J0).max())
    assert outs[0]['rows'] == (0, 18) and outs[1]['rows'] == (18, 30)
    assert np.count_nonzero(outs[0]['jac'][0].numpy()[:, 12:]) > 0
