import sys, os, shutil
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
from mspde import _ffi
_ffi.LIB_PATH = os.path.join(ROOT, "tools", "libmspde_prof.so")
import numpy as np, torch
from mspde.core import MspdeContext
ctx = MspdeContext(32, 64, 11430, 3)
N, M = ctx.N, ctx.M
ctx.set_inputs(np.zeros((M, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(M), np.ones(N), np.ones(N), 0.0)
torch.manual_seed(0)
L = torch.randn(N, N, dtype=torch.complex128, device="cuda"); w = torch.randn(N, dtype=torch.complex128, device="cuda")
ctx.H.copy_(torch.randn(M, N, dtype=torch.complex128, device="cuda"))
x = ctx.solve(L, w); torch.cuda.synchronize()
print("---- gibbs")
v = ctx.gibbs_lstsq(L, torch.randn(M + N, dtype=torch.complex128, device="cuda")); torch.cuda.synchronize()
