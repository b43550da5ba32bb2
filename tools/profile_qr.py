import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde.core import MspdeContext
ctx = MspdeContext(32, 64, 11430, 3)
N, M = ctx.N, ctx.M
ctx.set_inputs(np.zeros((M, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(M), np.ones(N), np.ones(N), 0.0)
torch.manual_seed(0)
L = torch.randn(N, N, dtype=torch.complex128, device="cuda"); w = torch.randn(N, dtype=torch.complex128, device="cuda")
ctx.H.copy_(torch.randn(M, N, dtype=torch.complex128, device="cuda"))
rhs = torch.randn(M + N, dtype=torch.complex128, device="cuda")
x = ctx.solve(L, w); v = ctx.gibbs_lstsq(L, rhs); torch.cuda.synchronize()
torch.cuda.profiler.start()
if sys.argv[1] == "solve":
    x = ctx.solve(L, w)
else:
    v = ctx.gibbs_lstsq(L, rhs)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
