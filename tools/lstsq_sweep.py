import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde.core import MspdeContext
ctx = MspdeContext(8, 64, 3048, 3)   # N=225, workspace (M+N) x N
N = ctx.N
ctx.set_inputs(np.zeros((3048, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(3048), np.ones(N), np.ones(N), 0.0)
torch.manual_seed(0)
for rows, cols in [(25, 25), (100, 25), (1041, 25), (1041, 24), (1041, 26), (1041, 32), (1041, 33), (1065, 49), (1041, 63), (1041, 64), (1041, 65), (2000, 130), (3000, 225), (129, 25), (65, 25), (64, 25)]:
    a = torch.randn(rows, cols, dtype=torch.complex128, device="cuda"); b = torch.randn(rows, dtype=torch.complex128, device="cuda")
    x = ctx.lstsq(a, b)
    ref = torch.linalg.lstsq(a, b.unsqueeze(1)).solution.squeeze(1)
    print(rows, cols, "rel", float((x - ref).abs().max() / ref.abs().max()))
