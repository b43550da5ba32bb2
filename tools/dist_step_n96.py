"""One block-distributed pCN iteration's worth of dense work at config 5 (n=96, N=36481, M=17190) -- or any n -- on P GPUs:
assembly (replicated), middle-layer solve (distributed pivoted LU), Phi(current) and Phi(latest) (distributed Cholesky
pipeline), Gibbs least squares (distributed Householder QR).  Prints per-piece times and the aggregate FP64 rate."""
import os, sys, time, ctypes, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from mspde import image as im, _ffi, util
from mspde.core import MspdeContext
from mspde.dist import DistPhi, DistLU, DistQR

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
n_ext = max(n, 64); n_theta = 90
il = 2 * n - 1; N = il * il; m = 2 * n_ext - 1; M = m * n_theta; nr = 2 * n * n - 2 * n + 1
f = im.FourierAnalysis_2D(n, n_ext)
sino = im.Sinogram("none.png", target_size=m, n_theta=n_theta, stdev=0.2)
H = sino.get_measurement_matrix_device(f.ix.ravel("F"), f.iy.ravel("F"), dev) / 0.2
ctx = MspdeContext(n, n_ext, M, 3, device=local, assembly_only=True)
D = torch.as_tensor(f.D_diag).to(dev)
_ffi.check(ctx.lib.mspde_ctx_set_inputs(ctx._h, None, 0, None, 0, None, 0, None, _ffi.ptr(D), None, ctypes.c_double(0.0)))
# hyper-parameters of Simulation.__init__ (image_cupy.py:839-880)
pi = np.float64(util.PI); beta_u = (4e7 ** 2) * 4 * pi; sb0 = np.float32(np.sqrt(beta_u)); sbv = np.float32(np.sqrt(beta_u * (3.2e4 / 4e7) ** 2))
Lu = ((f.D_diag / 5e9 - 5e9) / np.float64(sb0)).astype(np.float32); sd = (-1.0 / Lu.astype(np.float64)); sd[nr - 1] /= 2
rng = np.random.Generator(np.random.PCG64(7))
def w_sym():
    return torch.as_tensor(util.symmetrize(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr)))).to(dev)
y = torch.as_tensor(rng.standard_normal(M) * 30).to(dev)
rhs = torch.as_tensor(rng.standard_normal(M + N) + 0j).to(dev)
u0 = torch.as_tensor(sd).to(dev) * w_sym()
torch.cuda.synchronize(); dist.barrier()
T = {}
def tic(): torch.cuda.synchronize(); dist.barrier(); return time.perf_counter()
t = tic(); dp = DistPhi(N, M, nb=512, device=dev); dp.setup(H, y, 0.0)
nrm2 = torch.zeros(1, dtype=torch.float64, device=dev)
for i in dp.HtH_rows.owned:
    blk = dp.HtH_rows.block(i)[:, i * 512:]
    nrm2 += 2 * (torch.triu(blk, 1).abs() ** 2).sum() + (blk.diagonal().abs() ** 2).sum()
dist.all_reduce(nrm2); dp.delta = 1e-6 * float(nrm2.sqrt()); T["setup_HtH"] = tic() - t
lu, qr = DistLU(dev), DistQR(dev)
for rep in range(2):
    t0 = tic()
    L1 = ctx.construct_L(u0, float(np.float64(sbv) * np.sqrt(0.4))); T["assemble_mid"] = tic() - t0
    t = tic(); u1, logdet = lu.solve(L1, w_sym()); T["lu_solve"] = tic() - t
    t = tic(); L2 = ctx.construct_L(u1, float(sbv)); T["assemble_last"] = tic() - t
    del L1
    t = tic(); phi_a = dp.phi(L2); T["phi_current"] = tic() - t
    t = tic(); phi_b = dp.phi(L2); T["phi_latest"] = tic() - t
    t = tic(); v = qr.lstsq([H, L2], rhs); T["gibbs_qr"] = tic() - t
    total = tic() - t0
    F = 8 / 3 * N**3 + 2 * (4 * N**3 + 4 / 3 * N**3 + 4 * N * N * M + 4 * M * M * N + 4 / 3 * M**3) + 8 * (M + N) * N * N - 8 / 3 * N**3
    if rank == 0:
        print(json.dumps({"n": n, "N": N, "M": M, "world": world, "rep": rep, "seconds_per_iteration": total, "iters_per_s": 1 / total,
                          "flops": F, "aggregate_tflops": F / total / 1e12, "per_gpu_tflops": F / total / 1e12 / world,
                          "pieces_s": {k: round(x, 4) for k, x in T.items()}, "phi": phi_a, "phi_repeat_equal": phi_a == phi_b,
                          "logdet_L1": logdet, "u1_max": float(u1.abs().max()), "v_max": float(v.abs().max())}), flush=True)
    del L2
dist.destroy_process_group()
