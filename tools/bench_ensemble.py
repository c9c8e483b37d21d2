"""C5-shaped ensemble run: W Metropolis walkers on the 25-species EGFR-scale network (SSR / free energy only, no
sensitivities), data generated at the nominal parameters with the product's own integrator.  Prints walker-steps per
second (one walker-step = one residual evaluation of one candidate point + accept / reject)."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import sysbio_modeling_b200 as sb
from sysbio_modeling_b200 import model_files
from sysbio_modeling_b200.project import ensembles

W = int(sys.argv[1]) if len(sys.argv) > 1 else 131072            # walkers PER GPU (weak scaling under torchrun)
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, world = int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1'))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
    dist.init_process_group('nccl')
mod = model_files.module('egfr_signalling')
om = sb.OdeModel.from_model_file(model_files.path('egfr_signalling'))
nominal = np.array([mod.nominal[k] for k in om.param_order])
y0 = np.array(mod.initial_conditions)
tm = np.linspace(50, 600, 12)
ynom = om.simulate(nominal, np.concatenate([[0.0], tm]), y0)[1:]
exps = [sb.Experiment('EGF_stimulation',
                      [sb.TimecourseMeasurement('ERKPP', ynom[:, 24], tm, 0.05 * ynom[:, 24] + 0.1),
                       sb.TimecourseMeasurement('RasT', ynom[:, 16], tm, 0.05 * ynom[:, 16] + 0.1)],
                      initial_conditions=mod.initial_conditions)]
proj = sb.Project(om, exps, {}, {'ERKPP': ('direct', 24), 'RasT': ('direct', 16)})
th0 = np.log(nominal)
# the sampling matrix comes from the Gauss-Newton Hessian J^T J of the scaled Jacobian at the starting point (reference
# Ensembles.py:226-258): one sensitivity solve of the 25-state, 52-parameter model (1,325 augmented states, warp-group kernel);
# singular values below CUTOFF x the largest are raised to it (the reference's sing_val_cutoff): the quadratic model of a
# sloppy 52-parameter cost only holds along its stiff directions
hess = ensembles.gauss_newton_hessian(proj, th0)
CUTOFF = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-2
ensembles.ensemble_log_params_batch(proj, th0, hess=hess, steps=1, n_walkers=1024, seeds=1, sing_val_cutoff=CUTOFF)      # warm-up
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.time()
# independent walkers: every rank runs its own shard with its own random stream, no data-path collective
ens, fes, ratio = ensembles.ensemble_log_params_batch(proj, th0, hess=hess, steps=steps, n_walkers=W, seeds=2 + rank,
                                                      skip_elems=steps - 1, sing_val_cutoff=CUTOFF)
torch.cuda.synchronize()
dt = torch.tensor([time.time() - t0], dtype=torch.float64, device='cuda')
acc = torch.tensor([float(ratio.mean())], dtype=torch.float64, device='cuda')
if world > 1:
    dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    acc /= world
if rank == 0:
    print(json.dumps({'n_gpus': world, 'walkers_per_gpu': W, 'steps': steps, 'wall_s_max_over_ranks': float(dt),
                      'walker_steps_per_s': world * W * (steps + 1) / float(dt), 'acceptance_mean': float(acc),
                      'free_energy_finite_rank0': bool(np.isfinite(fes).all()), 'sing_val_cutoff': CUTOFF,
                      'hessian': 'Gauss-Newton (J / sigma)^T (J / sigma) at the starting point, warp-group sensitivity kernel'}))
if world > 1:
    dist.destroy_process_group()
