"""torchrun check of the block-distributed Phi: equality with the single-GPU path at n=32, timing at n=96 (optional)."""
import os, sys, time, ctypes
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from mspde import image as im, _ffi
from mspde.core import MspdeContext
from mspde.dist import DistPhi, BlockRows

rank = int(os.environ.get("RANK", 0)); local = int(os.environ.get("LOCAL_RANK", 0)); world = int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
mode = sys.argv[1] if len(sys.argv) > 1 else "n32"

if mode == "n32":
    sim = im.Simulation(n_layers=3, n_samples=4, n=32, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4,
                        meas_std=0.2, evaluation_interval=10, printProgress=False, seed=1, burn_percentage=25.0, enable_step_adaptation=True,
                        pcn_variant="dunlop", phantom_name="shepp.png", n_theta=90, device=local)
    sim.pcn.set_chol_epsilon(1e-6)
    pcn, ctx = sim.pcn, sim.pcn.ctx
    pcn._ensure_inputs(sim.Layers[0].stdev_sym_host)
    L = sim.Layers[2].LMat.current_L
    ref = ctx.phi(L)
    dp = DistPhi(ctx.N, ctx.M, nb=512, device=dev)
    # same HtH bits as the single-GPU path: slice the context's matrix into the owned row blocks
    hr = BlockRows(ctx.N, 512, world, rank, dev)
    for i in hr.owned:
        hr.block(i).copy_(ctx.HtH[i * 512:i * 512 + hr.rows(i), :])
    dp.setup(ctx.H, ctx.y, ctx.delta, HtH_rows=hr)
    val = dp.phi(L); val = dp.phi(L)
    t0 = time.perf_counter(); val = dp.phi(L); torch.cuda.synchronize(); t1 = time.perf_counter()
    if rank == 0:
        print(f"n=32 world={world}: single-GPU Phi {ref!r} distributed {val!r} rel {abs(val - ref) / abs(ref):.2e}  time {1e3 * (t1 - t0):.1f} ms  {dp.timings}")
    # own-HtH variant (distributed Gram of H)
    dp2 = DistPhi(ctx.N, ctx.M, nb=512, device=dev); dp2.setup(ctx.H, ctx.y, ctx.delta)
    v2 = dp2.phi(L)
    if rank == 0:
        print(f"   with distributed HtH: rel {abs(v2 - ref) / abs(ref):.2e}")
else:
    n, n_ext, n_theta = 96, 96, 90
    il = 2 * n - 1; N = il * il; m = 2 * n_ext - 1; M = m * n_theta
    f = im.FourierAnalysis_2D(n, n_ext)
    sino = im.Sinogram("none.png", target_size=m, n_theta=n_theta, stdev=0.2)
    t0 = time.perf_counter()
    H = sino.get_measurement_matrix_device(f.ix.ravel("F"), f.iy.ravel("F"), dev) / 0.2
    ctx = MspdeContext(n, n_ext, M, 3, device=local, assembly_only=True)
    D = torch.as_tensor(f.D_diag).to(dev)
    _ffi.check(ctx.lib.mspde_ctx_set_inputs(ctx._h, None, 0, None, 0, None, 0, None, _ffi.ptr(D), None, ctypes.c_double(0.0)))
    g = torch.Generator(device=dev).manual_seed(5)
    # smooth random field of O(1): low modes only
    u = torch.zeros(N, dtype=torch.complex128, device=dev)
    half = torch.complex(torch.randn(N // 2 + 1, dtype=torch.float64, device=dev, generator=g), torch.randn(N // 2 + 1, dtype=torch.float64, device=dev, generator=g))
    kx = torch.as_tensor(f.ix.ravel("F")[N // 2:]).to(dev); ky = torch.as_tensor(f.iy.ravel("F")[N // 2:]).to(dev)
    half = half * 30.0 / (1.0 + (kx.double() ** 2 + ky.double() ** 2)) ; half[0] = half[0].real.clone()
    u[N // 2:] = half; u[:N // 2] = torch.flip(half[1:], dims=(0,)).conj()
    L = ctx.construct_L(u, 1.13e5)
    y = torch.randn(M, dtype=torch.float64, device=dev, generator=g) * 30
    torch.cuda.synchronize()
    if rank == 0: print(f"n=96 setup (H on device, L assembled): {time.perf_counter() - t0:.1f} s; N={N} M={M}", flush=True)
    dp = DistPhi(N, M, nb=512, device=dev)
    t0 = time.perf_counter(); dp.setup(H, y, 0.0); torch.cuda.synchronize()
    # delta = 1e-6 * ||HtH||_F from the distributed row blocks
    nrm2 = torch.zeros(1, dtype=torch.float64, device=dev)
    for i in dp.HtH_rows.owned:
        blk = dp.HtH_rows.block(i)[:, i * 512:]
        up = torch.triu(blk, 1); dg = blk.diagonal()
        nrm2 += 2 * (up.abs() ** 2).sum() + (dg.abs() ** 2).sum()
    dist.all_reduce(nrm2); dp.delta = 1e-6 * float(nrm2.sqrt())
    if rank == 0: print(f"distributed HtH: {time.perf_counter() - t0:.1f} s, delta={dp.delta:.4g}", flush=True)
    for rep in range(2):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        val = dp.phi(L); torch.cuda.synchronize(); t1 = time.perf_counter()
        fl = 4 * N**3 + 4 / 3 * N**3 + 4 * N * N * M + 4 * M * M * N + 4 / 3 * M**3
        if rank == 0:
            print(f"n=96 world={world}: Phi={val!r}  {t1 - t0:.3f} s  {fl / (t1 - t0) / 1e12:.1f} TFLOP/s aggregate ({fl / (t1 - t0) / 1e12 / world:.1f} per GPU)  {dp.timings}", flush=True)
dist.destroy_process_group()
