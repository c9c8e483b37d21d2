#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path on B200 (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[3], the one north_star quotes its target on): Kholodenko MAPK cascade,
8 states x 22 parameters -> 184 sensitivity-augmented states, one experiment observing MAPK_PP and MKK_PP at
20 time points, 65,536 parameter sets per GPU (theta = log(nominal) + 0.25 N(0,1)), one step = one
evaluation of the full scaled Jacobian [40 x 22] + residuals + SSR + J^T r for every parameter set
(sensitivity-augmented Dormand-Prince solve + loss epilogue) + the per-set normal equations J^T J | J^T r that
the optimiser consumes (all-gathered over the ranks at N > 1).  Tolerances default to the reference's
hard-coded rtol = atol = 1e-10 (model/ode_model.py:122-123).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's own CPU path on the host cores

After the timed headline region the N = 1 run also times the other BASELINE configs (C2, C3, C5; reduced step count)
and checks 8 parameter sets of the headline against the CPU oracle (``parity``).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, 'oracle')):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np   # noqa: E402

METRIC = 'sensitivity_ode_solves_per_sec'
UNIT = 'solves/s'
SURVEY_FLOPS_SENS_RHS_C4 = 1004        # SymPy count_ops of f + sensitivity RHS after CSE


# ------------------------------------------------------------------------------------------------
# CPU reference path: the reference's UNMODIFIED OdeModel.calc_jacobian (baseline/_ref, see oracle/build_ref.py) -
# SciPy LSODA rtol = atol = 1e-10 on the 1000-node grid with the numba-compiled generated sensitivity RHS; falls back
# to the oracle port when the copy is absent
# ------------------------------------------------------------------------------------------------
_W = {}


def _single_thread_env():
    # LSODA's BDF mode factors dense 184x184 Jacobians through LAPACK; one BLAS thread per worker process
    for k in ('OMP_NUM_THREADS', 'OPENBLAS_NUM_THREADS', 'MKL_NUM_THREADS', 'NUMEXPR_NUM_THREADS'):
        os.environ[k] = '1'


def _cpu_init(tol):
    try:
        from threadpoolctl import threadpool_limits
        _W['tp'] = threadpool_limits(1)
    except Exception:
        pass
    from tools import workloads as wl
    from sysbio_modeling_b200.symbolic import make_ode_model
    import build_ref
    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    model, sens = make_ode_model(sm, False), make_ode_model(sm, True)
    RefOdeModel = build_ref.load_reference_odemodel() if tol == 1e-10 else None     # the reference hard-codes 1e-10
    kind = 'port'
    if RefOdeModel is not None:
        try:
            _W['ref'] = RefOdeModel(model, sens, sm.n_vars, list(sm.param_order), use_jit=True)     # numba, ode_model.py:46-57
            kind = 'reference'
        except Exception:
            _W.pop('ref', None)
    if 'ref' not in _W:
        import sbm_oracle
        try:
            import numba
            sens = numba.jit("void(f8[:], f8, f8[:], f8[:])", nopython=True)(sens)   # ode_model.py:50-51
        except Exception:
            pass
        _W.update(dict(oracle=sbm_oracle, sens=sens))
    _W.update(dict(sm=sm, y0=y0, tol=tol, kind=kind, t_end=float(wl.C4_T[-1])))
    _cpu_solve(np.log(nominal))          # warm up (numba compile excluded from timing)
    return kind


def _cpu_solve(theta):
    """One reference sensitivity solve: OdeModel.calc_jacobian on the 1000-node grid (base_project.py:510-515)."""
    sm = _W['sm']
    p = np.exp(theta)
    t_sim = np.linspace(0, _W['t_end'], 1000)
    y0 = np.concatenate([_W['y0'], np.zeros(sm.n_vars * sm.n_free)])
    if 'ref' in _W:
        S = _W['ref'].calc_jacobian(p, t_sim, y0)
    else:
        S = _W['oracle'].calc_jacobian(_W['sens'], sm.n_vars, p, t_sim, y0, rtol=_W['tol'], atol=_W['tol'])
    return float(S[-1, 0])


def _cpu_kind(_):
    return _W.get('kind', 'port')


def cpu_pool(tol, cores):
    import multiprocessing as mp
    _single_thread_env()
    ctx = mp.get_context('spawn')
    return ctx.Pool(cores, initializer=_cpu_init, initargs=(tol,))


def cpu_reference_rate(thetas, tol, cores=None):
    """Solves/s of the reference CPU path on `cores` host processes over the given parameter sets."""
    cores = cores or os.cpu_count() or 1
    with cpu_pool(tol, cores) as pool:
        kind = pool.map(_cpu_kind, range(cores))[0]
        pool.map(_cpu_solve, [thetas[0]] * cores)           # make sure every worker is warm
        t0 = time.perf_counter()
        pool.map(_cpu_solve, list(thetas), chunksize=max(1, len(thetas) // (cores * 4)))
        dt = time.perf_counter() - t0
    return len(thetas) / dt, dt, cores, kind


def cpu_sample_text(kind, n, tol, dt, cores):
    who = ('the reference\'s unmodified model/ode_model.py OdeModel.calc_jacobian (baseline/_ref)' if kind == 'reference'
           else 'oracle port of OdeModel.calc_jacobian')
    return ('%d parameter sets of the bench distribution; %s: SciPy odeint (LSODA) rtol=atol=%g on the 1000-node grid, '
            'numba-compiled generated sensitivity RHS; %.1f s wall on %d processes' % (n, who, tol, dt, cores))


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        if not self.samples:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        mhz = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith('active') for s in self.samples)]
        return {'sm_mhz': mhz[len(mhz) // 2] if mhz else None,
                'sm_max_mhz': float(self.samples[0][1]) if self.samples[0][1].replace('.', '').isdigit() else None,
                'reasons': reasons, 'samples': len(self.samples)}


# ------------------------------------------------------------------------------------------------
def workload_config(args, batch):
    return {'workload': 'C4 Kholodenko MAPK (8 states, 22 params, 184 augmented states), %d parameter sets per GPU, '
                        '1 experiment x 2 observables x 20 time points, full scaled Jacobian 40x22 + residuals + SSR + J^T r'
                        % batch,
            'param_sets_per_gpu': batch, 'rtol': args.tol, 'atol': args.tol, 'solver': 'dp5',
            'cache': 'L2 flushed (256 MiB write) between timed steps; Jacobian output (%.0f MB) exceeds L2'
                     % (batch * 40 * 22 * 8 / 1e6)}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from tools import workloads as wl
    sm, nominal, _ = wl.nominal_of('mapk_kholodenko')
    cores = os.cpu_count() or 1
    per_step = args.cpu_sample if args.cpu_sample > 0 else 100 * cores      # ~3 s of CPU work per step
    thetas = wl.make_thetas(nominal, per_step * (args.steps + args.warmup), 4, 0.25)
    with cpu_pool(args.tol, cores) as pool:
        kind = pool.map(_cpu_kind, range(cores))[0]
        pool.map(_cpu_solve, [thetas[0]] * cores)
        k = 0
        for _ in range(args.warmup):
            pool.map(_cpu_solve, list(thetas[k:k + per_step]))
            k += per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            pool.map(_cpu_solve, list(thetas[k:k + per_step]))
            k += per_step
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': workload_config(args, args.batch),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                             'sample': 'each step: ' + cpu_sample_text(kind, per_step, args.tol, dt / args.steps, cores)},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------
# the other BASELINE configs (N = 1 only, after the headline): solves/s, residual / Jacobian evaluations per second
# including the loss epilogue, algorithmic flops from the kernels' counters, fraction of the measured FP64 peak
# ------------------------------------------------------------------------------------------------
def _time_eval(proj, theta, with_jac, solver, reps):
    import torch
    bufs = {}
    proj.evaluate_device(theta, with_jacobian=with_jac, solver=solver, out=bufs)
    torch.cuda.synchronize()
    best = None
    for _ in range(reps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        opts = proj._opts(proj._jac_solver() if with_jac else solver)
        bufs['sims'].zero_()
        bufs['counters'].zero_()
        ev[0].record()
        proj.device_plan.solve_device(theta, opts, bufs['sims'], bufs.get('jac') if with_jac else None, bufs['status'],
                                      bufs['counters'])
        ev[1].record()
        proj.device_plan.loss_device(theta, bufs['sims'], bufs.get('jac') if with_jac else None, bufs['resid'], bufs['ssr'],
                                     bufs['sf'], bufs.get('sfgrad') if with_jac else None,
                                     bufs.get('grad') if with_jac else None, loss=proj._loss_name(), flags=proj._flags())
        ev[2].record()
        torch.cuda.synchronize()
        t = (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]))
        if best is None or sum(t) < sum(best):
            best = t
    from sysbio_modeling_b200 import _lib
    cnt = dict(zip(_lib.COUNTER_NAMES, [int(x) for x in bufs['counters'].cpu().numpy()[:5]]))
    failed = int((bufs['status'] != 0).sum().item())
    return best, cnt, failed


def extra_configs(peak_tf, quick=False):
    import torch
    from tools import workloads as wl
    dev = torch.device('cuda')
    out = {}
    # ---- C3: Robertson, 1,048,576 parameter sets, residuals only, Rosenbrock kernel (rtol 1e-6, atol 1e-10) ----
    B3 = 131072 if quick else 1048576
    proj, _ = wl.c3_project()
    sm, nominal, _ = wl.nominal_of('robertson')
    th = torch.from_numpy(wl.make_thetas(nominal, B3, 3, 0.5)).to(dev)
    (ms_i, ms_l), cnt, failed = _time_eval(proj, th, False, 'rodas4', 2)
    cs = proj._model.compiled.cuda
    N, s_ = 3, 6
    f_step = cs.flops_jac_dense + (2.0 / 3.0) * N ** 3 + s_ * (2 * N * N + cs.flops_f) + N * (2 * s_ * (s_ - 1) + 2 * s_ + 5)
    flops = cnt['steps'] * f_step
    out['c3'] = {'workload': 'C3 Robertson (3 states), %d parameter sets, 2 observables x 12 log-spaced times, residuals only, '
                             'RODAS4 rtol 1e-6 atol 1e-10' % B3,
                 'solves_per_sec': B3 / ms_i * 1e3, 'residual_evals_per_sec': B3 / (ms_i + ms_l) * 1e3,
                 'sbm_rodas4_plain_ms': ms_i, 'sbm_loss_ms': ms_l, 'loss_share_of_step': ms_l / (ms_i + ms_l),
                 'steps_per_solve': cnt['steps'] / B3, 'lu_per_solve': cnt['lu'] / B3, 'failed': failed,
                 'algorithmic_flops': flops, 'flops_per_step': f_step, 'tflops': flops / ms_i / 1e9,
                 'frac_fp64_peak': flops / ms_i / 1e9 / peak_tf}
    # (beyond the config: the stiff model's scaled Jacobian through RODAS4 with forward sensitivities, 3 x 3 columns)
    try:
        Bj = B3 // 8
        (ms_ij, ms_lj), cntj, failedj = _time_eval(proj, th[:Bj].contiguous(), True, 'rodas4', 2)
        out['c3'].update({'jacobian_sets': Bj, 'jacobian_evals_per_sec': Bj / (ms_ij + ms_lj) * 1e3,
                          'sbm_rodas4_sens_ms': ms_ij, 'jacobian_steps_per_solve': cntj['steps'] / Bj,
                          'jacobian_failed': failedj})
    except Exception as exc:
        out['c3']['jacobian_error'] = repr(exc)
    del th
    # ---- C5: EGFR-scale network (25 states, 52 parameters), 131,072 walkers, SSR only, states-only DP5 in push form ----
    B5 = 16384 if quick else 131072
    proj, _ = wl.c5_project()
    sm, nominal, _ = wl.nominal_of('egfr_signalling')
    th = torch.from_numpy(wl.make_thetas(nominal, B5, 5, 0.1)).to(dev)
    (ms_i, ms_l), cnt, failed = _time_eval(proj, th, False, 'dp5', 2)
    cs = proj._model.compiled.cuda
    flops = cnt['rhs_evals'] * cs.flops_f + cnt['steps'] * 57 * sm.n_vars
    out['c5'] = {'workload': 'C5 EGFR-scale network (25 states, 52 parameters), %d walkers, 2 observables x 12 times, SSR only, '
                             'DP5 rtol = atol = 1e-10' % B5,
                 'solves_per_sec': B5 / ms_i * 1e3, 'residual_evals_per_sec': B5 / (ms_i + ms_l) * 1e3,
                 'sbm_dp5_push_ms': ms_i, 'sbm_loss_ms': ms_l, 'steps_per_solve': cnt['steps'] / B5, 'failed': failed,
                 'algorithmic_flops': flops, 'flops_per_rhs': cs.flops_f, 'tflops': flops / ms_i / 1e9,
                 'frac_fp64_peak': flops / ms_i / 1e9 / peak_tf}
    del th
    # ---- C2: reversible MM, experiment-local parameters, sensitivities.  (a) the integrator on 4,096 experiments x 64 sets
    #      through OdeModel (raw dy/dp), (b) the Project objective with 4,096 experiments: one residual / Jacobian evaluation
    #      = 4,096 solves; the dense scaled Jacobian [65,536 x 16,384] is 8.6 GB per parameter set ----
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('reversible_mm')
    om = sb.OdeModel.from_model_file(model_files.path('reversible_mm'))
    nominal = np.array([mod.nominal[k] for k in om.param_order])
    n2 = 4096 * (8 if quick else 64)
    P = nominal * np.exp(0.3 * np.random.default_rng(2).standard_normal((n2, 4)))
    t2 = np.concatenate([[0.0], wl.C2_T])
    om.calc_jacobian_batch(P[:256], t2, mod.initial_conditions)
    best = 1e30
    for _ in range(2):
        om.calc_jacobian_batch(P, t2, mod.initial_conditions)
        best = min(best, om.last_solve_ms)
    cs = om.compiled.cuda
    cnt = om.last_counters
    flops = cnt['rhs_evals'] * cs.flops_sens_rhs + cnt['steps'] * 57 * 20
    out['c2'] = {'workload': 'C2 reversible MM (4 states, 4 params, 20 augmented states), 4,096 experiments x %d parameter sets, '
                             '16 time points, sensitivities, DP5 rtol = atol = 1e-10' % (n2 // 4096),
                 'solves_per_sec': n2 / best * 1e3, 'sbm_dp5_ct_ms': best, 'steps_per_solve': cnt['steps'] / n2,
                 'failed': int((om.last_status != 0).sum()), 'algorithmic_flops': flops, 'flops_per_rhs': cs.flops_sens_rhs,
                 'tflops': flops / best / 1e9, 'frac_fp64_peak': flops / best / 1e9 / peak_tf}
    if not quick:
        E = 4096
        proj, _ = wl.c2_project(E)
        th = torch.from_numpy(wl.c2_theta(E, 2)).to(dev)
        (ms_i, ms_l), cnt, failed = _time_eval(proj, th, True, 'dp5', 1)
        out['c2'].update({'project_experiments': E, 'project_sets': 2,
                          'jacobian_evals_per_sec': 2 / (ms_i + ms_l) * 1e3, 'project_solve_ms': ms_i, 'project_loss_ms': ms_l,
                          'project_jacobian_bytes_per_set': 16 * E * 4 * E * 8, 'project_failed': failed})
        (ms_i, ms_l), cnt, failed = _time_eval(proj, th, False, 'dp5', 2)
        out['c2'].update({'residual_evals_per_sec': 2 / (ms_i + ms_l) * 1e3})
    return out


def parity_check(proj, make_oracle, thetas, out_gpu):
    """8 parameter sets of the timed batch against the CPU oracle (OracleProject: LSODA 1e-10 on the 1000-node grid)."""
    oracle = make_oracle()
    worst = {'ssr_rel': 0.0, 'resid_abs': 0.0, 'jac_rel_to_max': 0.0, 'grad_rel_to_max': 0.0, 'sf_rel': 0.0}
    for b in range(len(thetas)):
        r0 = oracle.residuals(thetas[b])
        sf0 = np.array([oracle.sf[g] for g in oracle.sf_groups])
        J0 = oracle.jacobian(thetas[b])
        g0 = (J0.T * r0).sum(axis=1)
        ssr0 = 0.5 * np.sum(r0 ** 2)
        worst['ssr_rel'] = max(worst['ssr_rel'], abs(out_gpu['ssr'][b] - ssr0) / ssr0)
        worst['resid_abs'] = max(worst['resid_abs'], float(np.abs(out_gpu['resid'][b] - r0).max()))
        worst['jac_rel_to_max'] = max(worst['jac_rel_to_max'], float(np.abs(out_gpu['jac'][b] - J0).max() / np.abs(J0).max()))
        worst['grad_rel_to_max'] = max(worst['grad_rel_to_max'], float(np.abs(out_gpu['grad'][b] - g0).max() / np.abs(g0).max()))
        worst['sf_rel'] = max(worst['sf_rel'], float(np.abs(out_gpu['sf'][b] / sf0 - 1).max()))
    ok = worst['ssr_rel'] <= 1e-6 and worst['jac_rel_to_max'] <= 1e-5 and worst['sf_rel'] <= 1e-7
    return dict(worst, sets=len(thetas), ok=bool(ok),
                oracle='OracleProject: the reference path restated (LSODA rtol = atol = 1e-10, 1000-node grid); '
                       'bars: SSR 1e-6 relative (north_star), Jacobian 1e-5 of its largest entry, scale factors 1e-7')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=65536, help='parameter sets per GPU')
    ap.add_argument('--tol', type=float, default=1e-10, help='rtol = atol (reference: 1e-10)')
    ap.add_argument('--cpu-sample', type=int, default=0, help='parameter sets in the CPU baseline sample (0 = auto)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-configs', action='store_true', help='skip the C2 / C3 / C5 measurements after the headline')
    ap.add_argument('--quick-configs', action='store_true', help='reduced batch sizes for C2 / C3 / C5')
    ap.add_argument('--shape', default='0,0,0', help='group_lanes,cols_per_lane,sens_layout of the sensitivity kernel (0,0,0 = auto; layout 1 = lane-per-column groups, 2 = CTA-transposed)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = max(args.warmup, 1)

    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sysbio_modeling_b200 import _lib, engine
    from sysbio_modeling_b200.distributed import ShardedProject
    from tools import workloads as wl
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    _lib.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    sm, nominal, y0 = wl.nominal_of('mapk_kholodenko')
    B = args.batch
    shape = tuple(int(x) for x in args.shape.split(','))
    proj, make_oracle = wl.c4_project(args.tol, shape)      # data: the model at the nominal parameters with 5 % noise (seeded)
    t = proj.plan_tables
    info = proj._model.compiled.info()
    cuda_src = proj._model.compiled.cuda
    sharded = ShardedProject(proj)

    # every rank draws the SAME sample of the parameter distribution (weak scaling: fixed work per GPU; different samples
    # differ by ~1 % in their total step count, which only shows up as skew at the gather)
    theta_np = wl.make_thetas(nominal, B, 4, 0.25)
    theta_h = torch.from_numpy(theta_np).pin_memory()
    theta = theta_h.to(dev)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    n_packed = t.n_proj * (t.n_proj + 1) // 2 + t.n_proj + 1
    packed_all = torch.empty((B * world, n_packed), dtype=torch.float64, device=dev) if world > 1 else None

    for _ in range(max(args.warmup, 1)):
        sharded.step_local_blocks(theta, packed_all)
    torch.cuda.synchronize()
    bufs = sharded._bufs

    # ---- timed region: K steps, device time by CUDA events on the launching (current) stream ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    opts_sens = proj._opts('dp5')
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        for k in range(args.steps):
            flush.fill_(float(k))
            ev[k][0].record()
            # the launches of ShardedProject.step_local_blocks, with events between them
            proj.device_plan.solve_device(theta, opts_sens, bufs['sims'].zero_(), bufs['jac'], bufs['status'],
                                          bufs['counters'].zero_())
            ev[k][1].record()
            proj.device_plan.loss_device(theta, bufs['sims'], bufs['jac'], bufs['resid'], bufs['ssr'], bufs['sf'],
                                         bufs['sfgrad'], bufs['grad'], loss='square')
            ev[k][2].record()
            packed = engine.normal_equations_device(bufs['jac'], bufs['resid'], out=bufs['normal'])
            if world > 1:
                dist.all_gather_into_tensor(packed_all, packed)
            ev[k][3].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    ms_step = [ev[k][0].elapsed_time(ev[k][3]) for k in range(args.steps)]
    ms_solve = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    ms_loss = [ev[k][1].elapsed_time(ev[k][2]) for k in range(args.steps)]
    ms_gather = [ev[k][2].elapsed_time(ev[k][3]) for k in range(args.steps)]
    total_ms = float(sum(ms_step))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    per_rank = torch.tensor([float(np.mean(ms_solve)), float(np.mean(ms_loss)), float(np.mean(ms_gather))],
                            dtype=torch.float64, device=dev)
    per_rank_all = [torch.zeros_like(per_rank) for _ in range(world)]
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank_all, per_rank)
    else:
        per_rank_all = [per_rank]
    total_ms_max = float(tmax.item())
    n_solves_rank = B * t.E
    value = n_solves_rank * world * args.steps / (total_ms_max * 1e-3)

    counters = bufs['counters'].cpu().numpy()
    status_bad = int((bufs['status'] != 0).sum().item())
    cnt = dict(zip(_lib.COUNTER_NAMES, [int(x) for x in counters[:5]]))
    n_aug = sm.n_vars * (sm.n_free + 1)
    flops_survey = cnt['rhs_evals'] * SURVEY_FLOPS_SENS_RHS_C4 + cnt['steps'] * 57 * n_aug
    flops_emitter = cnt['rhs_evals'] * cuda_src.flops_sens_rhs + cnt['steps'] * 57 * n_aug
    solve_ms = float(np.mean(ms_solve))
    gpu_first = {k: bufs[k][:8].cpu().numpy() for k in ('ssr', 'resid', 'jac', 'grad', 'sf')}

    # ---- end-to-end through the C-ABI host entry points: pinned host theta in, pinned host results out.  Evaluations are
    #      issued back to back through the two workspace slots of the plan (sbm_jacobian_host_begin / sbm_host_end): every
    #      step's copy-in and copy-out are inside the timed region, the copy-out of step i overlaps the integration of
    #      step i + 1 (it cannot overlap its own: the persistent integrator fills every SM).  The serial single-call form
    #      (sbm_jacobian_host) is timed too. ----
    import ctypes
    L = _lib.lib()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream
    hb = []
    for _ in range(2):
        hb.append({'jac': torch.empty((B, t.n_rows, t.n_proj), dtype=torch.float64).pin_memory(),
                   'resid': torch.empty((B, t.n_rows), dtype=torch.float64).pin_memory(),
                   'ssr': torch.empty((B,), dtype=torch.float64).pin_memory(),
                   'grad': torch.empty((B, t.n_proj), dtype=torch.float64).pin_memory()})

    def e2e_begin(slot):
        h = hb[slot]
        _lib.check(L.sbm_jacobian_host_begin(proj.device_plan.handle, ctypes.byref(opts_sens), 0, 0, B,
                                             vp(theta_h.data_ptr()), vp(h['jac'].data_ptr()), vp(h['resid'].data_ptr()),
                                             vp(h['ssr'].data_ptr()), vp(h['grad'].data_ptr()), None, None, None, None,
                                             vp(st), slot))

    def e2e_end(slot):
        _lib.check(L.sbm_host_end(proj.device_plan.handle, slot))

    def e2e_serial():
        h = hb[0]
        _lib.check(L.sbm_jacobian_host(proj.device_plan.handle, ctypes.byref(opts_sens), 0, 0, B,
                                       vp(theta_h.data_ptr()), vp(h['jac'].data_ptr()), None, vp(h['resid'].data_ptr()),
                                       vp(h['ssr'].data_ptr()), vp(h['grad'].data_ptr()), None, None, None, None, vp(st)))
    e2e_serial()                                      # warm-up (allocates the workspaces of slot 0)
    e2e_begin(1)
    e2e_end(1)
    n_e2e = max(3, min(args.steps, 5))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_begin(i & 1)
        if i > 0:
            e2e_end((i - 1) & 1)
    e2e_end((n_e2e - 1) & 1)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    e2e_serial()
    torch.cuda.synchronize()
    e2e_serial_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt, e2e_serial_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_solves_rank * world * n_e2e / float(e2e_t[0].item())
    e2e_serial_value = n_solves_rank * world / float(e2e_t[1].item())
    h2d = B * t.n_proj * 8
    d2h = B * (t.n_rows * t.n_proj + t.n_rows + 1 + t.n_proj) * 8
    # both paths give the same numbers (same kernels, same arithmetic per parameter set)
    same = all(bool(torch.equal(h['ssr'].to(dev), bufs['ssr'])) and bool(torch.equal(h['jac'][:64].to(dev), bufs['jac'][:64]))
               for h in hb)

    peak_tf, peak_ms = _lib.fp64_peak()
    achieved_tf = flops_survey / (solve_ms * 1e-3) / 1e12
    achieved_tf_emitter = flops_emitter / (solve_ms * 1e-3) / 1e12

    ct = info['sens_layout'] == 2
    kname = 'sbm_dp5_ct' if ct else 'sbm_dp5_sens'
    kern_info = {kname + '_ms': solve_ms, 'sbm_loss_tile_kernel_ms': float(np.mean(ms_loss)),
                 'sbm_gram_kernel_plus_gather_ms': float(np.mean(ms_gather)),
                 'per_rank_ms': [{'solve': float(x[0]), 'loss': float(x[1]), 'normal_eq_gather': float(x[2])} for x in per_rank_all],
                 'loss_algorithmic_gbs': B * (2 * t.n_rows * t.n_proj + 3 * t.n_rows + 3 * t.n_proj + 4) * 8 / (np.mean(ms_loss) * 1e-3) / 1e9,
                 'sens_layout': 'cta-transposed (lane = trajectory, warp = column group)' if ct
                 else 'lane-per-column warp groups'}
    if ct:
        kern_info.update({'warps_per_cta': info['ct_warps'], 'cols_per_warp': info['ct_cols'],
                          'cols_on_state_warp': info['ct_cols_state'], 'dyn_smem_bytes': info['ct_dyn_smem_bytes'],
                          'ctas_per_sm': info['ct_ctas_per_sm'], 'regs': info['regs_dp5_ct'],
                          'local_bytes': info['local_dp5_ct']})
    else:
        kern_info.update({'group_lanes': info['group_lanes'], 'cols_per_lane': info['cols_per_lane'],
                          'block_threads': info['block_threads'], 'dyn_smem_bytes': info['dyn_smem_bytes'],
                          'regs': info['regs_dp5_sens'], 'local_bytes': info['local_dp5_sens']})
    # DRAM traffic of the dominant kernel: from the committed ncu capture of THIS configuration (same batch), else null
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r02_traffic.json')))
        if tr.get('kernel') == kname and int(tr['param_sets']) == B:
            traffic = tr['dram_bytes_read'] + tr['dram_bytes_write']
            traffic_note = 'dram__bytes_read.sum + dram__bytes_write.sum of one launch at %d parameter sets (%s)' % (
                tr['param_sets'], tr['source'])
    except Exception:
        pass
    line = None
    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': total_ms_max / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic', 'config': workload_config(args, B),
            'jacobian_evals_per_sec': B * world * args.steps / (total_ms_max * 1e-3),
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h,
                    'calls': n_e2e, 'matches_device_path': same,
                    'serial_single_call': e2e_serial_value,
                    'how': 'sbm_jacobian_host_begin / sbm_host_end (C ABI), pinned host buffers, %d evaluations back to back '
                           'through two workspace slots: the copy-out of one overlaps the integration of the next; '
                           'serial_single_call = one sbm_jacobian_host call (copies not overlapped)' % n_e2e},
            'gpu_launches': 3 * args.steps,
            'multi_gpu': {'path': 'ShardedProject.step_local_blocks (sysbio_modeling_b200/distributed.py)',
                          'exchange': 'all_gather_into_tensor of the packed per-set normal equations J^T J | J^T r | r^T r '
                                      '(%d doubles per set, %.1f MB per rank)' % (n_packed, B * n_packed * 8 / 1e6),
                          'theta': 'same sample of the distribution on every rank'},
            'kernels': kern_info,
            'counters_per_step': cnt, 'failed_trajectories': status_bad,
            'roofline': {'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': traffic, 'traffic_note': traffic_note,
                         'algorithmic_bytes_per_launch': B * t.E * (8 * (sm.n_params + sm.n_vars) + 8 * 40 * (1 + t.n_proj) + 4),
                         'kernel': kname,
                         'peak_source': 'measured: sbm_fp64_peak DFMA microbenchmark in this run '
                                        '(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)',
                         'algorithmic_flops_per_launch': flops_survey,
                         'flops_model': 'SURVEY.md 8(d): rhs_evals x %d + attempted steps x 57 x %d'
                                        % (SURVEY_FLOPS_SENS_RHS_C4, n_aug),
                         'achieved_emitter_count': achieved_tf_emitter, 'frac_emitter_count': achieved_tf_emitter / peak_tf,
                         'flops_model_emitter_count': 'rhs_evals x %d (op count of the emitted f, J, G + 22 columns x (2 nnz(J) FMA flops) '
                                                      '+ nnz(G)) + steps x 57 x %d' % (cuda_src.flops_sens_rhs, n_aug)},
            'clocks': clocks.summary(),
        }
    if world > 1:
        dist.barrier()

    if rank == 0:
        # (at N > 1 too: rank 0's first 8 sets; the other ranks have left the timed region and wait in the teardown)
        try:
            line['parity'] = parity_check(proj, make_oracle, theta_np[:8], gpu_first)
        except Exception as exc:                      # the bench line must survive a broken checker
            line['parity'] = {'error': repr(exc)}
    if rank == 0 and world == 1:
        if not args.no_configs:
            del hb, flush
            for k in list(bufs.keys()):
                del bufs[k]
            torch.cuda.empty_cache()
            try:
                line['configs'] = extra_configs(peak_tf, args.quick_configs)
            except Exception as exc:
                line['configs'] = {'error': repr(exc)}
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample = args.cpu_sample if args.cpu_sample > 0 else min(B, 400 * cores)       # ~12 s of CPU work
            rate, dt, cores, kind = cpu_reference_rate(theta_np[:sample], args.tol, cores)
            line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': kind,
                                    'sample': 'first ' + cpu_sample_text(kind, sample, args.tol, dt, cores)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
