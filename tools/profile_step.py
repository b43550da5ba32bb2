"""One warm pCN step at config-2 size between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde import image as im, util
mode = sys.argv[1] if len(sys.argv) > 1 else "step"
sim = im.Simulation(n_layers=3, n_samples=8, n=32, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4,
                    meas_std=0.2, evaluation_interval=10, printProgress=False, seed=1, burn_percentage=25.0, enable_step_adaptation=True,
                    pcn_variant="dunlop", phantom_name="shepp.png", n_theta=90)
sim.pcn.set_chol_epsilon(1e-6)
pcn, ctx, Layers = sim.pcn, sim.pcn.ctx, sim.Layers
pcn.one_step(Layers); pcn.one_step(Layers)
torch.cuda.synchronize()
if mode == "step":
    torch.cuda.profiler.start()
    pcn.one_step(Layers)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
else:   # the dominant launch alone: Q_inv = I - Ht^H Ht (upper), m=n=11430, k=3969; and the assembly kernel
    X = ctx.buffer("X", ctx.N, ctx.M); Q = ctx.buffer("Q", ctx.M, ctx.M)
    ctx.gemm_tn(X, X, True, -1.0, 0.0, C=Q, upper=True); torch.cuda.synchronize()
    torch.cuda.profiler.start()
    ctx.gemm_tn(X, X, True, -1.0, 0.0, C=Q, upper=True)
    ctx.construct_L(Layers[1].current_sample_sym, 1e5, out=Layers[2].LMat.latest_computed_L)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
