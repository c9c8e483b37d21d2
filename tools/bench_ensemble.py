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

W = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
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
# a diagonal stand-in for the Hessian keeps the set-up cheap: (the real J^T J needs the sensitivity kernel of a 25-state,
# 52-parameter model, which the CTA-transposed layout does not cover yet)
hess = np.eye(len(th0)) * 2.0e4
ensembles.ensemble_log_params_batch(proj, th0, hess=hess, steps=1, n_walkers=1024, seeds=1)      # warm-up
torch.cuda.synchronize()
t0 = time.time()
ens, fes, ratio = ensembles.ensemble_log_params_batch(proj, th0, hess=hess, steps=steps, n_walkers=W, seeds=2,
                                                      skip_elems=steps - 1)
torch.cuda.synchronize()
dt = time.time() - t0
print(json.dumps({'walkers': W, 'steps': steps, 'wall_s': dt, 'walker_steps_per_s': W * (steps + 1) / dt,
                  'acceptance_mean': float(ratio.mean()), 'free_energy_finite': bool(np.isfinite(fes).all()),
                  'free_energy_start': float(fes[0].mean()), 'free_energy_end': float(fes[-1].mean())}))
