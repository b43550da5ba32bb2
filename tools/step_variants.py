"""Dev tool: device time per step for flag combinations (serial / concurrent / skip-gibbs / cached phi)."""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde import image as im, util
from mspde.core import STEP_CACHE_PHI_CUR, STEP_SKIP_GIBBS, STEP_SERIAL
sim = im.Simulation(n_layers=3, n_samples=8, n=32, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4,
                    meas_std=0.2, evaluation_interval=10, printProgress=False, seed=1, burn_percentage=25.0, enable_step_adaptation=True,
                    pcn_variant="dunlop", phantom_name="shepp.png", n_theta=90)
sim.pcn.set_chol_epsilon(1e-6)
pcn, ctx, Layers = sim.pcn, sim.pcn.ctx, sim.Layers
pcn._ensure_inputs(Layers[0].stdev_sym_host)
cb = pcn._bind(Layers)
rng = np.random.Generator(np.random.PCG64(3)); nr, M, K = ctx.n_ravel, ctx.M, 3
def noise():
    w = [ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))) for _ in range(K - 1)]
    return w, -0.3, ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))), ctx.to_dev(rng.standard_normal(M))
nz = [noise() for _ in range(4)]
def run(flags, reps=3):
    for i in range(2): ctx.pcn_step(cb, *nz[i], flags=flags)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): ctx.pcn_step(cb, *nz[i % 4], flags=flags)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for name, fl in [("serial", STEP_SERIAL), ("concurrent", 0), ("serial+skipgibbs", STEP_SERIAL | STEP_SKIP_GIBBS), ("conc+skipgibbs", STEP_SKIP_GIBBS),
                 ("serial+cache", STEP_SERIAL | STEP_CACHE_PHI_CUR), ("conc+cache", STEP_CACHE_PHI_CUR),
                 ("serial+cache+skip", STEP_SERIAL | STEP_CACHE_PHI_CUR | STEP_SKIP_GIBBS), ("conc+cache+skip", STEP_CACHE_PHI_CUR | STEP_SKIP_GIBBS)]:
    print(f"{name:20s} {run(fl):8.2f} ms")
