"""torchrun check of the block-distributed LU solve and QR least squares against the single-GPU paths."""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from mspde.core import MspdeContext
from mspde.dist import DistLU, DistQR
rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
M = int(sys.argv[2]) if len(sys.argv) > 2 else 900
ctx = MspdeContext(n, 64, M, 3, device=local); N = ctx.N
g = torch.Generator(device=dev).manual_seed(3)
rnd = lambda *s: torch.complex(torch.randn(*s, dtype=torch.float64, device=dev, generator=g), torch.randn(*s, dtype=torch.float64, device=dev, generator=g))
L = rnd(N, N); H = rnd(M, N); w = rnd(N); rhs = rnd(M + N)
ctx.set_inputs(H, torch.zeros(N, N, dtype=torch.complex128, device=dev), np.zeros(M), np.ones(N), np.ones(N), 0.0)
Lp = ctx.new_matrix(N, N); Lp.copy_(L)
x1, ld1 = ctx.solve(Lp, w, want_logdet=True)
v1 = ctx.gibbs_lstsq(Lp, rhs)
torch.cuda.synchronize(); t0 = time.perf_counter()
x2, ld2 = DistLU(dev).solve(L, w); torch.cuda.synchronize(); t1 = time.perf_counter()
v2 = DistQR(dev).lstsq([H, L], rhs); torch.cuda.synchronize(); t2 = time.perf_counter()
if rank == 0:
    print(f"N={N} M={M} world={world}: LU x rel {float((x2 - x1).abs().max() / x1.abs().max()):.2e} logdet {ld1!r} {ld2!r} ({t1 - t0:.3f} s)   "
          f"QR v rel {float((v2 - v1).abs().max() / v1.abs().max()):.2e} ({t2 - t1:.3f} s)")
dist.destroy_process_group()
