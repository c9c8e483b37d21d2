#!/usr/bin/env python
"""bench.py - headline benchmark of the hot path on B200 (contract: see the task statement / DESIGN.md).

Workload (BASELINE.json configs[3], the one north_star quotes its target on): Kholodenko MAPK cascade,
8 states x 22 parameters -> 184 sensitivity-augmented states, one experiment observing MAPK_PP and MKK_PP at
20 time points, 65,536 parameter sets per GPU (theta = log(nominal) + 0.25 N(0,1)), one step = one
evaluation of the full scaled Jacobian [40 x 22] + residuals + SSR + J^T r for every parameter set
(sensitivity-augmented Dormand-Prince solve + loss epilogue).  Tolerances default to the reference's
hard-coded rtol = atol = 1e-10 (model/ode_model.py:122-123).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores
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
OBS = {'MAPK_PP': 7, 'MKK_PP': 4}
T_OBS = np.linspace(150.0, 3000.0, 20)


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def mapk_model_info():
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('mapk_kholodenko')
    from sysbio_modeling_b200.symbolic import symbolic_model_from_file
    sm = symbolic_model_from_file(model_files.path('mapk_kholodenko'), name='mapk')
    nominal = np.array([mod.nominal[p] for p in sm.param_order])
    return sm, nominal, np.array(mod.initial_conditions)


def make_thetas(nominal, B, seed):
    rng = np.random.default_rng(seed)
    return np.log(nominal)[None, :] + 0.25 * rng.standard_normal((B, len(nominal)))


def build_project(tol, data=None, shape=(0, 0, 0)):
    import sysbio_modeling_b200 as sb
    from sysbio_modeling_b200 import model_files
    mod = model_files.module('mapk_kholodenko')
    om = sb.OdeModel.from_model_file(model_files.path('mapk_kholodenko'), model_name='mapk', group_lanes=shape[0],
                                     cols_per_lane=shape[1], sens_layout=shape[2])
    if data is None:
        data = {k: np.ones_like(T_OBS) for k in OBS}
    meas = [sb.TimecourseMeasurement(k, data[k], T_OBS, 0.05 * np.abs(data[k]) + 1e-3) for k in OBS]
    exp = sb.Experiment('MAPK_cascade', meas, initial_conditions=mod.initial_conditions)
    return sb.Project(om, [exp], {}, {k: ('direct', v) for k, v in OBS.items()}, sf_groups=list(OBS.keys()),
                      rtol=tol, atol=tol)


# ------------------------------------------------------------------------------------------------
# CPU reference path (oracle port: SciPy LSODA + numba-compiled generated sensitivity RHS)
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
    import sbm_oracle
    from sysbio_modeling_b200.symbolic import make_ode_model
    sm, nominal, y0 = mapk_model_info()
    sens = make_ode_model(sm, True)
    try:
        import numba
        sens = numba.jit("void(f8[:], f8, f8[:], f8[:])", nopython=True)(sens)   # ode_model.py:50-51
        jitted = True
    except Exception:
        jitted = False
    _W.update(dict(oracle=sbm_oracle, sens=sens, sm=sm, y0=y0, tol=tol, jitted=jitted))
    # warm up (numba compile excluded from timing)
    _cpu_solve(np.log(nominal))
    return jitted


def _cpu_solve(theta):
    """One reference sensitivity solve: OdeModel.calc_jacobian on the 1000-node grid (base_project.py:510-515)."""
    O, sm = _W['oracle'], _W['sm']
    p = np.exp(theta)
    t_sim = np.linspace(0, T_OBS[-1], 1000)
    y0 = np.concatenate([_W['y0'], np.zeros(sm.n_vars * sm.n_free)])
    S = O.calc_jacobian(_W['sens'], sm.n_vars, p, t_sim, y0, rtol=_W['tol'], atol=_W['tol'])
    return float(S[-1, 0])


def cpu_reference_rate(thetas, tol, cores=None):
    """Solves/s of the reference CPU path on `cores` host processes over the given parameter sets."""
    import multiprocessing as mp
    cores = cores or os.cpu_count() or 1
    _single_thread_env()
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(tol,)) as pool:
        pool.map(_cpu_solve, [thetas[0]] * cores)           # make sure every worker is warm
        t0 = time.perf_counter()
        pool.map(_cpu_solve, list(thetas), chunksize=max(1, len(thetas) // (cores * 4)))
        dt = time.perf_counter() - t0
    return len(thetas) / dt, dt, cores


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
def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    sm, nominal, _ = mapk_model_info()
    cores = os.cpu_count() or 1
    per_step = args.cpu_sample if args.cpu_sample > 0 else 100 * cores      # ~3 s of CPU work per step
    thetas = make_thetas(nominal, per_step * (args.steps + args.warmup), 4)
    import multiprocessing as mp
    _single_thread_env()
    ctx = mp.get_context('spawn')
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(args.tol,)) as pool:
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
            'config': dict(workload_config(args, args.batch), sample_per_step=per_step),
            'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d parameter sets per step, SciPy odeint (LSODA) rtol=atol=%g on the reference\'s '
                                       '1000-node grid, numba-compiled generated sensitivity RHS, %d processes'
                                       % (per_step, args.tol, cores)},
            'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line))
    return 0


def workload_config(args, batch):
    return {'workload': 'C4 Kholodenko MAPK (8 states, 22 params, 184 augmented states), %d parameter sets per GPU, '
                        '1 experiment x 2 observables x 20 time points, full scaled Jacobian 40x22 + residuals + SSR + J^T r'
                        % batch,
            'param_sets_per_gpu': batch, 'rtol': args.tol, 'atol': args.tol, 'solver': 'dp5',
            'cache': 'L2 flushed (256 MiB write) between timed steps; Jacobian output (%.0f MB) exceeds L2'
                     % (batch * 40 * 22 * 8 / 1e6)}


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
    ap.add_argument('--shape', default='0,0,0', help='group_lanes,cols_per_lane,sens_layout of the sensitivity kernel (0,0,0 = auto; layout 1 = lane-per-column groups, 2 = CTA-transposed)')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'b200':
        args.warmup = max(args.warmup, 1)

    if args.impl == 'reference':
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from sysbio_modeling_b200 import _lib
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    _lib.require_gpu()
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    sm, nominal, y0 = mapk_model_info()
    B = args.batch
    # synthetic data: the model at the nominal parameters with 5 % noise (seeded)
    shape = tuple(int(x) for x in args.shape.split(','))
    proj0 = build_project(args.tol, shape=shape)
    sims = proj0.residuals_batch(np.log(nominal)[None, :])['sims'][0]
    rng = np.random.default_rng(40)
    labels = [l[1] for l in proj0._labels]
    data = {k: np.array([sims[i] for i, l in enumerate(labels) if l == k]) * (1 + 0.05 * rng.standard_normal(20))
            for k in OBS}
    proj = build_project(args.tol, data, shape)
    t = proj.plan_tables
    info = proj._model.compiled.info()
    cuda_src = proj._model.compiled.cuda

    theta_h = torch.from_numpy(make_thetas(nominal, B, 4 + rank)).pin_memory()
    theta = theta_h.to(dev)
    bufs = {}
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    def step():
        proj.evaluate_device(theta, with_jacobian=True, out=bufs)
        if world > 1:
            dist.all_gather_into_tensor(ssr_all, bufs['ssr'])

    ssr_all = torch.empty(B * world, dtype=torch.float64, device=dev) if world > 1 else None
    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device time by CUDA events on the launching (current) stream ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    opts_sens = proj._opts('dp5')
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        for k in range(args.steps):
            flush.fill_(float(k))
            ev[k][0].record()
            # same two launches as proj.evaluate_device, with an event between them
            proj.device_plan.solve_device(theta, opts_sens, bufs['sims'].zero_(), bufs['jac'], bufs['status'],
                                          bufs['counters'].zero_())
            ev[k][1].record()
            proj.device_plan.loss_device(theta, bufs['sims'], bufs['jac'], bufs['resid'], bufs['ssr'], bufs['sf'],
                                         bufs['sfgrad'], bufs['grad'], loss='square')
            if world > 1:
                dist.all_gather_into_tensor(ssr_all, bufs['ssr'])
            ev[k][2].record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
    ms_step = [ev[k][0].elapsed_time(ev[k][2]) for k in range(args.steps)]
    ms_solve = [ev[k][0].elapsed_time(ev[k][1]) for k in range(args.steps)]
    total_ms = float(sum(ms_step))
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    n_solves_rank = B * t.E
    value = n_solves_rank * world * args.steps / (total_ms_max * 1e-3)

    counters = bufs['counters'].cpu().numpy()
    status_bad = int((bufs['status'] != 0).sum().item())
    cnt = dict(zip(_lib.COUNTER_NAMES, [int(x) for x in counters[:5]]))
    n_aug = sm.n_vars * (sm.n_free + 1)
    flops_launch = cnt['rhs_evals'] * cuda_src.flops_sens_rhs + cnt['steps'] * 57 * n_aug
    solve_ms = float(np.mean(ms_solve))

    # ---- end-to-end through the C-ABI host entry point: pinned host theta in, pinned host results out ----
    jac_h = torch.empty((B, t.n_rows, t.n_proj), dtype=torch.float64).pin_memory()
    resid_h = torch.empty((B, t.n_rows), dtype=torch.float64).pin_memory()
    ssr_h = torch.empty((B,), dtype=torch.float64).pin_memory()
    grad_h = torch.empty((B, t.n_proj), dtype=torch.float64).pin_memory()
    import ctypes
    L = _lib.lib()
    vp = ctypes.c_void_p
    st = torch.cuda.current_stream().cuda_stream

    def e2e_call():
        _lib.check(L.sbm_jacobian_host(proj.device_plan.handle, ctypes.byref(opts_sens), 0, 0, B,
                                       vp(theta_h.data_ptr()), vp(jac_h.data_ptr()), None, vp(resid_h.data_ptr()),
                                       vp(ssr_h.data_ptr()), vp(grad_h.data_ptr()), None, None, None, None, vp(st)))
    e2e_call()
    n_e2e = max(2, min(args.steps, 3))
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_e2e):
        e2e_call()
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_t = torch.tensor([e2e_dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_value = n_solves_rank * world * n_e2e / float(e2e_t.item())
    h2d = B * t.n_proj * 8
    d2h = B * (t.n_rows * t.n_proj + t.n_rows + 1 + t.n_proj) * 8
    # spot-check that both paths agree
    same = bool(torch.allclose(ssr_h.to(dev), bufs['ssr'], rtol=1e-9, atol=0))

    peak_tf, peak_ms = _lib.fp64_peak()
    achieved_tf = flops_launch / (solve_ms * 1e-3) / 1e12

    ct = info['sens_layout'] == 2
    kname = 'sbm_dp5_ct' if ct else 'sbm_dp5_sens'
    kern_info = {kname + '_ms': solve_ms, 'sbm_loss_kernel_ms': float(np.mean(ms_step)) - solve_ms,
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
    # DRAM traffic of the dominant kernel from the committed ncu capture, scaled to this launch's batch
    traffic, traffic_note = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'r01_traffic.json')))
        if tr.get('kernel') == kname:
            traffic = (tr['dram_bytes_read'] + tr['dram_bytes_write']) * B / float(tr['param_sets'])
            traffic_note = 'dram read+write of one launch at %d parameter sets (%s), scaled by batch' % (
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
                    'calls': n_e2e, 'matches_device_path': same},
            'gpu_launches': 2 * args.steps,
            'kernels': kern_info,
            'counters_per_step': cnt, 'failed_trajectories': status_bad,
            'roofline': {'bound': 'fp64', 'achieved': achieved_tf, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': achieved_tf / peak_tf, 'traffic': traffic, 'traffic_note': traffic_note,
                         'algorithmic_bytes_per_launch': B * t.E * (8 * (sm.n_params + sm.n_vars) + 8 * 40 * (1 + t.n_proj) + 4),
                         'kernel': kname,
                         'peak_source': 'measured: sbm_fp64_peak DFMA microbenchmark in this run '
                                        '(MEASURED_PEAKS.json has no FP64 entry; nominal 148 SM x 64 DFMA/clk x 2 x 1.965 GHz = 37.2)',
                         'algorithmic_flops_per_launch': flops_launch,
                         'flops_model': 'rhs_evals x %d (CSE op count of f,J,G once + %d columns x J*s + G) + steps x 57 x %d'
                                        % (cuda_src.flops_sens_rhs, sm.n_free, n_aug)},
            'clocks': clocks.summary(),
        }
    if world > 1:
        dist.barrier()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample = args.cpu_sample if args.cpu_sample > 0 else min(B, 400 * cores)       # ~12 s of CPU work
        th = make_thetas(nominal, sample, 4)
        rate, dt, cores = cpu_reference_rate(th, args.tol, cores)
        line['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                                'sample': 'first %d of the %d parameter sets; SciPy odeint (LSODA) rtol=atol=%g on the '
                                          'reference\'s 1000-node grid with the numba-compiled generated sensitivity RHS '
                                          '(reference OdeModel.calc_jacobian path), %.1f s wall on %d processes'
                                          % (sample, B, args.tol, dt, cores)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
