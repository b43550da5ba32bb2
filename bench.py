#!/usr/bin/env python
"""bench.py -- pCN iterations/s of the B200-native hot path (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # ours (one process per GPU under torchrun for N>1)
  python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the numpy oracle port on the host cores

Workload (config.workload): BASELINE.json configs[1] -- 2-D Shepp-Logan tomography, 3 layers, n=32 Fourier
modes/axis (N=3969), n_ext=64 (m=127), 90 angles (M=11430), single chain per GPU; synthetic sinogram.
A "step" is one pCN.one_step (reference-faithful formulation: Phi(current) recomputed every iteration).
  value : iterations/s with the injected noise already resident in HBM (device-timed, CUDA events)
  e2e   : iterations/s through the public API pCN.one_step(Layers, noise=...) with the step's noise copied from
          pinned host memory and the accept flag + recorded samples read back every step
For N>1 the chains are independent (one per rank, different seeds): no data-path collective, one final NCCL
all_gather of the recorded samples inside the timed region; scaling is weak.
"""
from __future__ import annotations

import argparse
import datetime
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200"))
sys.path.insert(0, ROOT)

import numpy as np

CONFIGS = {
    1: dict(n=16, n_ext=64, n_theta=45, n_layers=2),
    2: dict(n=32, n_ext=64, n_theta=90, n_layers=3),
    3: dict(n=64, n_ext=64, n_theta=90, n_layers=3),
    5: dict(n=96, n_ext=96, n_theta=90, n_layers=3),   # one chain, every dense piece block-distributed (needs >= 2 GPUs)
}
FP64_PEAK_FALLBACK = 37.15  # DMMA m8n8k4 microbenchmark of round 1 (profiles/r01_fp64_probe.log); bench.py re-measures it live
REFERENCE_BUDGET_S = 200.0  # wall budget of the timed region of --impl reference
BENCH_BETA = 0.1            # pCN step of the bench chains: 25 % acceptance over the timed region (profiles/r02_beta_sweep.log)


def flops_iter(n, n_ext, n_theta, n_layers):
    """F_iter = (layers-2)*8/3 N^3 + 2 F_Phi + F_QR (real flops, structure-exploiting count)."""
    N = (2 * n - 1) ** 2
    M = (2 * n_ext - 1) * n_theta
    f_phi = 4 * N ** 3 + 4 / 3 * N ** 3 + 4 * N * N * M + 4 * M * M * N + 4 / 3 * M ** 3
    f_qr = 8 * (M + N) * N * N - 8 / 3 * N ** 3
    return (n_layers - 2) * 8 / 3 * N ** 3 + 2 * f_phi + f_qr, N, M


_GUARD_CHILD = r"""
import sys
fallback = sys.stdin.readline()
rest = sys.stdin.read()                 # returns when the parent closes the pipe -- or dies
if "DISARM" not in rest and fallback.strip():
    sys.stdout.write(fallback if fallback.endswith("\n") else fallback + "\n")
    sys.stdout.flush()
"""


class LineGuard:
    """Insurance for the one JSON line while rank 0 is inside a collective phase that another rank can stall: a small child
    process holds a complete fallback line (everything measured so far) and prints it to the result descriptor if this process
    ends without disarming it -- e.g. killed by NCCL's watchdog.  disarm() is called on every normal path, so stdout still
    carries exactly one line."""

    def __init__(self, fallback_line):
        out = _RESULT_FD if _RESULT_FD is not None else 1
        sys.stdout.flush()
        self.proc = subprocess.Popen([sys.executable, "-c", _GUARD_CHILD], stdin=subprocess.PIPE, stdout=out, close_fds=True)
        self.proc.stdin.write((json.dumps(fallback_line) + "\n").encode())
        self.proc.stdin.flush()

    def disarm(self):
        if self.proc is None:
            return
        try:
            self.proc.stdin.write(b"DISARM\n")
            self.proc.stdin.close()
            self.proc.wait(timeout=30)
        except Exception as ex:                                   # never lose the real line to the insurance
            sys.stderr.write(f"line guard: {ex!r}\n")
        self.proc = None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


_RESULT_FD = None


def reserve_stdout():
    """stdout carries exactly one JSON line: keep the real stdout aside and send everything else that libraries write to
    file descriptor 1 (NCCL's version banner, for one) to stderr."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    text = json.dumps(line) + "\n"
    if _RESULT_FD is None:
        sys.stdout.write(text)
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, text.encode())


# ----------------------------------------------------------------------------------------------------------------
# CPU arm: the numpy oracle (port of the reference step) on the host cores, on a bounded sample
# ----------------------------------------------------------------------------------------------------------------
def cpu_oracle_rate(cfg, steps=1, warmup=0, budget_s=30.0):
    """Times the oracle's one_step AT THE LABELLED CONFIG (same n, angles, layers; never a smaller geometry scaled up).
    Runs at most `steps` iterations and stops early once `budget_s` of wall time is used (at least one iteration always
    completes), so a K-step request stays bounded.  Returns (iters_per_s, description, cores, seconds_per_iter, steps_timed)."""
    from oracle import mspde_oracle as o
    try:
        # torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU arm is meant to use every host core it may run on
        from threadpoolctl import threadpool_info, threadpool_limits
        avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        threadpool_limits(limits=avail)
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        cores = int(os.environ.get("OMP_NUM_THREADS", 0)) or os.cpu_count() or 1
    a = (np.random.rand(1500, 1500) + 1j * np.random.rand(1500, 1500))
    a @ a                                                   # BLAS thread pool warm-up
    hp = o.Hyper(n_layers=cfg["n_layers"], n=cfg["n"], n_ext=cfg["n_ext"], n_theta=cfg["n_theta"], beta=BENCH_BETA)
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    ns = o.NoiseStream(11, hp.n_layers, ft.n_ravel, H.shape[0])
    K = hp.n_layers
    st = o.init_chain(hp, ft, [ns.half() for _ in range(K)], [ns.half() for _ in range(K)])
    for _ in range(warmup):
        o.one_step(st, ft, H, HtH, y, delta, ns.draw_iteration())
    done, t0 = 0, time.perf_counter()
    while done < max(1, steps):
        o.one_step(st, ft, H, HtH, y, delta, ns.draw_iteration())
        done += 1
        if time.perf_counter() - t0 >= budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    desc = (f"{done} full oracle iteration(s) of the labelled workload itself (n={cfg['n']}, {cfg['n_theta']} angles, "
            f"{cfg['n_layers']} layers; {dt:.2f} s each, numpy/OpenBLAS on {cores} threads); no scaling or extrapolation")
    return 1.0 / dt, desc, cores, dt, done


def run_reference(args, cfg, rank, world):
    """--impl reference: the CPU port of the reference step on all host cores, at the labelled config.  One n=32 iteration is
    25-40 s of CPU, so the timed run is bounded by wall time (REFERENCE_BUDGET_S) and reports how many steps it timed."""
    if rank != 0:
        return
    rate, desc, cores, dt, done = cpu_oracle_rate(cfg, steps=max(1, args.steps), warmup=0, budget_s=REFERENCE_BUDGET_S)
    line = {"impl": "reference", "metric": f"pCN iters/sec (n={cfg['n']}, {cfg['n_layers']} layers)", "value": rate, "unit": "iters/s",
            "n_gpus": args.gpus, "steps": done, "warmup": 0, "steps_requested": args.steps, "warmup_requested": args.warmup,
            "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(cfg, args.gpus),
            "cpu_baseline": {"value": rate, "unit": "iters/s", "cores": cores, "kind": "port", "sample": desc},
            "e2e": {"value": rate, "unit": "iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": f"timed {done} of the requested {args.steps} steps (wall budget {REFERENCE_BUDGET_S:.0f} s; BLAS needs no warm-up "
                    "steps); value = steps / their wall time, ms_per_step x steps = the timed wall time"}
    emit(line)


def library_standin_ms(sim, noise, reps=2):
    """The same pCN iteration written with library calls only -- torch.fft (cuFFT), `@` (cuBLAS ZGEMM), torch.linalg cholesky /
    solve / solve_triangular / lstsq (cuSOLVER) -- i.e. what the reference's CuPy path does on this box (BASELINE.md 4.4,
    SURVEY 8d: "the reference's library-call GPU path on the same box our kernels must beat").  Structure-exploiting where the
    libraries allow it (triangular solve instead of the reference's LU of a triangle, lstsq instead of an explicit Q), same
    inputs, same process.  Returns ms per step (CUDA events, best of `reps` after one warm-up)."""
    import torch
    from mspde import util
    pcn, ctx, Layers, f = sim.pcn, sim.pcn.ctx, sim.Layers, sim.fourier
    N, M = ctx.N, ctx.M
    n, il = f.basis_number, 2 * f.basis_number - 1
    dev = ctx.device
    wh, log_u, wl, e = noise
    idx = torch.arange(il * il, device=dev)
    blk, inn = idx // il, idx % il
    dij, dkl = blk[:, None] - blk[None, :], inn[:, None] - inn[None, :]
    mask = (dij.abs() <= n - 1) & (dkl.abs() <= n - 1)
    r_i, c_i = (dkl + n - 1).clamp(0, il - 1), (dij + n - 1).clamp(0, il - 1)
    del dij, dkl
    zero = torch.zeros((), dtype=torch.complex128, device=dev)
    I_N = torch.eye(N, dtype=torch.complex128, device=dev)
    I_M = torch.eye(M, dtype=torch.complex128, device=dev)
    Hc, Hct, yc = ctx.H.contiguous(), ctx.Hct.contiguous(), ctx.y.to(torch.complex128)

    def assemble(u_sym, sb):
        uh = util.from_u_2D_ravel_to_uHalf_2D(u_sym, n)
        img = f.irfft2(uh)
        tp = util.symmetrize_2D(f.rfft2(torch.exp(img)))
        tm = util.symmetrize_2D(f.rfft2(torch.exp(-img)))
        return (torch.where(mask, tp[r_i, c_i], zero) * ctx.D_diag - torch.where(mask, tm[r_i, c_i], zero)) / sb

    def phi(L):
        r = L.conj().T @ L + ctx.HtH + ctx.delta * I_N
        c = torch.linalg.cholesky(r)
        Ht = torch.linalg.solve_triangular(c, Hct, upper=False)
        C = torch.linalg.cholesky(I_M - Ht.conj().T @ Ht)
        return 0.5 * torch.linalg.norm(yc @ C) ** 2 - torch.log(torch.diagonal(C).real).sum()

    def step():
        K = len(Layers)
        u_prev = None
        for k in range(K - 1):
            nn = pcn.betaZ * Layers[k]._noise_cur + pcn.beta * util.symmetrize(wh[k])
            if k == 0:
                u_prev = ctx.stdev_sym * nn
            else:
                Lk = assemble(u_prev, Layers[k].LMat.sqrt_beta)
                u_prev = torch.linalg.solve(Lk, nn)
                del Lk
        phi_cur = phi(Layers[-1].LMat.current_L.contiguous())
        L2 = assemble(u_prev, Layers[-1].LMat.sqrt_beta)
        phi_new = phi(L2)
        acc = bool((phi_cur - phi_new) > log_u)                              # the reference's host sync (573)
        nn = pcn.betaZ * (Layers[-1]._noise_new if acc else Layers[-1]._noise_cur) + pcn.beta * util.symmetrize(wl)
        rhs = torch.cat(((ctx.y - e).to(torch.complex128), -nn))
        A = torch.cat((Hc, L2), 0)
        return torch.linalg.lstsq(A, rhs.unsqueeze(1)).solution

    step()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def config5_block(cfg, dev, rank, world, steps, warmup, fp64_peak, selfcheck_n=(32,), local_rank=0):
    """BASELINE configs[4] as a sampler: one n=96 chain (N=36481, M=17190) stepped by mspde.dist.DistPCN with every dense piece
    block-distributed over the ranks.  Returns the record (collective: every rank must call it)."""
    import torch
    import torch.distributed as dist
    from mspde import _ffi
    from mspde.dist import build_dist_chain
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import dist_chain_check as dcc
    f_iter, N, M = flops_iter(**cfg)
    out = {"config": {"workload": f"BASELINE configs[4]: 3 layers, n={cfg['n']} (N={N}, M={M}), single chain, every factorization "
                                  "block-distributed over NCCL/NVLink (mspde.dist.DistPCN: state, injected noise, accept on the "
                                  "all-reduced Phi values, Gibbs draw with both outcomes' right-hand sides)"}}
    # (0) every rank must enter or skip together: a rank that runs out of memory mid-factorization leaves the others waiting in
    # a broadcast.  Give back what earlier phases of this process kept, then agree on the smallest free HBM of any rank.
    torch.cuda.synchronize()
    _ffi.release_gemm_workspaces()
    torch.cuda.empty_cache()
    need = config5_bytes_per_rank(N, M, world)
    free = torch.tensor([float(torch.cuda.mem_get_info(dev)[0])], dtype=torch.float64, device=dev)
    dist.all_reduce(free, op=dist.ReduceOp.MIN)
    out["hbm"] = {"needed_bytes_per_rank_estimate": need, "min_free_bytes_over_ranks": float(free.item())}
    if float(free.item()) < need:
        out["unavailable"] = (f"needs about {need / 2**30:.0f} GiB of free HBM on every rank, the fullest rank has "
                              f"{float(free.item()) / 2**30:.0f} GiB: skipped on all ranks")
        return out
    # The distributed sampler's numbers were taken with the in-loop 3M kernel (profiles/r02_bench_8gpu_b.json); its per-rank
    # panel updates are short-k products the packed kernel does not help, and the packed operands' workspaces are HBM this
    # configuration needs.  Process-wide switch, restored on the way out.
    pre_min0 = _ffi.get_gemm_pre_min()
    _ffi.set_gemm_pre_min(0)
    try:
        return _config5_measure(out, cfg, dev, rank, world, steps, warmup, fp64_peak, selfcheck_n, local_rank, f_iter, N, M)
    finally:
        _ffi.set_gemm_pre_min(pre_min0)
        torch.cuda.empty_cache()


def config5_bytes_per_rank(N, M, world):
    """Free HBM a rank needs for the n=96 distributed step: the peak seen by torch's allocator on 8 B200 was 126.5 GiB at
    N=36481, M=17190 (profiles/r02_bench_8gpu_c.json's out-of-memory report: 106.7 GiB live + a 19.8 GiB request; most of it
    is replicated N x N matrices, so it shrinks slowly with the rank count).  Other sizes are scaled by the bytes of the
    problem's dense matrices -- an (M+N) x (M+N) stacked matrix, four N x N, two M x N complex128: 141 GiB at n=96 -- and by
    sqrt(8 / world); 15 % is added for allocator fragmentation (17.9 GiB were reserved-but-unallocated in that report)."""
    blocks = 16.0 * ((M + N) ** 2 + 4.0 * N * N + 2.0 * M * N)
    at8 = 16.0 * ((17190 + 36481) ** 2 + 4.0 * 36481 ** 2 + 2.0 * 17190 * 36481)
    return 1.15 * 126.5 * 2**30 * (blocks / at8) * (8.0 / world) ** 0.5


def _config5_measure(out, cfg, dev, rank, world, steps, warmup, fp64_peak, selfcheck_n, local_rank, f_iter, N, M):
    import torch
    import torch.distributed as dist
    from mspde import _ffi
    from mspde.dist import build_dist_chain
    import dist_chain_check as dcc
    # (1) correctness evidence recorded with the number: distributed sampler == single-GPU kernels on the same inputs
    out["selfcheck_vs_single_gpu"] = [dcc.single_vs_dist(n, dev, rank) for n in selfcheck_n]
    torch.cuda.empty_cache()
    # (2) the n=96 chain
    t0 = time.perf_counter()
    chain, noise_fn = build_dist_chain(cfg["n"], cfg["n_ext"], cfg["n_theta"], cfg["n_layers"], dev, seed=7, beta=0.3)
    torch.cuda.synchronize()
    out["setup_s"] = time.perf_counter() - t0

    def timed_step():
        nz = noise_fn()
        torch.cuda.synchronize()
        dist.barrier()
        t = time.perf_counter()
        d = chain.one_step(*nz)
        torch.cuda.synchronize()
        dist.barrier()
        return time.perf_counter() - t, d
    for _ in range(warmup):
        timed_step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    runs = [timed_step() for _ in range(steps)]
    out["clocks"] = sampler.stop()
    tt = torch.tensor([sum(r[0] for r in runs)], dtype=torch.float64, device=dev)
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    sec = float(tt.item()) / steps
    out.update({"value": 1.0 / sec, "ms_per_step": 1e3 * sec, "gemm_scheme": "3m" if _ffi.get_gemm_mode() == _ffi.GEMM_3M else "4m",
                "steps_detail": [dict(seconds=r[0], accepted=bool(r[1]["accepted"]), phi_cur=r[1]["phi_cur"], phi_new=r[1]["phi_new"],
                                      phase_s=r[1].get("phase_s"))
                                 for r in runs],
                "finite": bool(all(np.isfinite(r[1]["phi_new"]) for r in runs) and torch.isfinite(chain.u_new[-1].abs()).all()),
                "roofline": {"bound": "tensor", "achieved": f_iter / sec / 1e12 / world, "peak": fp64_peak, "unit": "TFLOP/s per GPU",
                             "frac": f_iter / sec / 1e12 / world / fp64_peak, "aggregate_tflops": f_iter / sec / 1e12,
                             "executed_frac": (0.75 if _ffi.get_gemm_mode() == _ffi.GEMM_3M else 1.0) * f_iter / sec / 1e12 / world / fp64_peak,
                             "traffic": None}})
    # (3) the collective that bounds a panel step: ncclBroadcast of the widest Cholesky panel (512 x N complex128)
    pan = torch.zeros((512, N), dtype=torch.complex128, device=dev)
    for _ in range(2):
        dist.broadcast(torch.view_as_real(pan), src=0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        dist.broadcast(torch.view_as_real(pan), src=0)
    e1.record()
    torch.cuda.synchronize()
    ms_b = e0.elapsed_time(e1) / 5
    out["panel_broadcast"] = {"what": "ncclBroadcast of one 512-row Cholesky panel (512 x N complex128), the largest message of a panel "
                                      "step; hidden behind the K=512 trailing update by the one-panel look-ahead",
                              "bytes": 512 * N * 16, "ms": ms_b, "GB_per_s": 512 * N * 16 / ms_b * 1e-6,
                              "trailing_update_ms_at_peak": 0.75 * 4.0 * 512 * (N / world) * N / (fp64_peak * 1e12) * 1e3}
    del pan, chain
    torch.cuda.empty_cache()
    return out


def config4_block(cfg, dev, rank, world, local_rank, total_chains=64, steps_per_chain=2):
    """BASELINE configs[3]: 64 independent chains sharded over the GPUs (chain c -> rank c mod G), stepped round-robin on each
    GPU through one shared context (mspde.parallel.MultiChain), then ONE NCCL gather of the recorded samples and the pooled
    acceptance / Welford field statistics.  Strong scaling: the 64-chain workload is fixed, the GPUs vary."""
    import torch
    import torch.distributed as dist
    from mspde import image as im, parallel
    K = cfg["n_layers"]
    mine = parallel.shard_chains(total_chains, world, rank)
    sim = im.Simulation(n_layers=K, n_samples=steps_per_chain + 2, n=cfg["n"], n_extended=cfg["n_ext"], beta=BENCH_BETA, kappa=5e9,
                        sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4, meas_std=0.2, evaluation_interval=10, printProgress=False,
                        seed=parallel.chain_seed(1, mine[0]), burn_percentage=25.0, enable_step_adaptation=True, pcn_variant="dunlop",
                        phantom_name="shepp.png", meas_type="tomo", n_theta=cfg["n_theta"], verbose=False, device=local_rank)
    sim.pcn.set_chol_epsilon(1e-6)
    mc = parallel.MultiChain(sim, mine, base_seed=1, accumulate_stats=True)
    for ch in mc.chains:
        make_state_consistent(ch["Layers"])
    mc.step_all()                                   # one warm-up iteration of every chain
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps_per_chain):
        mc.step_all()
    hist = torch.as_tensor(mc.histories()).to(dev)                           # [local chains, K, rows, n_ravel]
    full = parallel.gather_samples(hist, total_chains, world)               # the final NCCL gather (configs[3])
    stats = mc.pooled_field_statistics()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    a, n = mc.accepted_and_steps()
    acc = parallel.pooled_acceptance(a, n)
    sim.pcn.ctx.check_info("config 4 chains")
    ms = float(t.item())
    out = {"workload": f"BASELINE configs[3]: {total_chains} independent chains, {K} layers, n={cfg['n']}, sharded c mod G over {world} GPUs, "
                       "final NCCL gather of samples + pooled acceptance and Welford field statistics inside the timed region",
           "chains": total_chains, "chains_per_gpu": len(mine), "steps_per_chain": steps_per_chain, "scaling": "strong",
           "value": total_chains * steps_per_chain / (ms * 1e-3), "unit": "iters/s", "ms_total": ms,
           "gathered_shape": list(full.shape), "pooled_acceptance": acc, "pooled_stats_count": int(stats[0]["count"]),
           "pooled_u_mean_abs_max": float(stats[-1]["u_mean"].abs().max())}
    sim.pcn.ctx.close()
    del mc, sim, hist, full
    torch.cuda.empty_cache()
    return out


def make_state_consistent(Layers):
    """Bench setup only.  Simulation.__init__ leaves every layer's current sample and current noise from DIFFERENT draws
    (image_cupy.py:735 vs 755), so until a first acceptance every proposal -- built from the noise -- is a jump
    to an unrelated state and the accept branch is never timed.  Here the noise is set to the one that reproduces the current
    samples (xi_0 = u_0 / stdev, xi_k = L_k u_k): a valid chain state whose small-beta proposals are local moves, so the timed
    region accepts at the usual 15-35 %."""
    import torch
    for k, l in enumerate(Layers):
        if k == 0:
            sd = torch.as_tensor(l.stdev_sym_host, device=l._u_cur.device)
            l._noise_cur.copy_(l._u_cur / sd)
        else:
            l._noise_cur.copy_(torch.mv(l.LMat.current_L.contiguous(), l._u_cur))
        l._noise_new.copy_(l._noise_cur)


def measure_fp64_peak(_ffi):
    """FP64 DMMA peak of THIS box, measured now by mspde_fp64_peak_probe (MEASURED_PEAKS.json has no FP64 entry)."""
    import ctypes
    v = ctypes.c_double(0.0)
    try:
        _ffi.check(_ffi.lib().mspde_fp64_peak_probe(_ffi.current_stream(), ctypes.byref(v)), "fp64_peak_probe")
        if 10.0 < v.value < 100.0:
            return float(v.value), ("measured live on this GPU by mspde_fp64_peak_probe (register-resident DMMA m8n8k4 chain, 16 "
                                    "warps/SM, best of 3); MEASURED_PEAKS.json has no FP64 entry")
    except Exception as ex:                                   # never lose the bench line to the probe
        sys.stderr.write(f"fp64 probe failed: {ex}\n")
    return FP64_PEAK_FALLBACK, "round-1 probe value (profiles/r01_fp64_probe.log); live probe unavailable"


def workload_config(cfg, gpus):
    idx = {16: 0, 32: 1, 64: 2, 96: 4}.get(cfg["n"], 1)       # position in BASELINE.json "configs"
    N, M = (2 * cfg["n"] - 1) ** 2, (2 * cfg["n_ext"] - 1) * cfg["n_theta"]
    return {"workload": f"BASELINE configs[{idx}]: Shepp-Logan tomography, {cfg['n_layers']} layers, n={cfg['n']} modes/axis, "
                        f"n_ext={cfg['n_ext']}, {cfg['n_theta']} angles, one chain per GPU",
            "formulation": "reference-faithful (Phi(current) recomputed every step, Gibbs QR every step)",
            "cache": f"working set per step (L {16e-6 * N * N:.0f} MB, Q_inv {16e-9 * M * M:.1f} GB, H {16e-6 * N * M:.0f} MB) >> 126 MB L2; "
                     "no explicit flush",
            "parallelism": f"{gpus} independent chain(s), no data-path collective; final NCCL all_gather of samples"}


# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cache-phi", action="store_true", help="also report the results-preserving cached-Phi(current) rate")
    ap.add_argument("--no-extras", action="store_true", help="skip the separately-reported results-preserving formulations")
    ap.add_argument("--gemm", default=None, choices=["3m", "4m"],
                    help="complex product scheme of the GEMM kernel for the headline (default: the library's, 3m)")
    ap.add_argument("--with-config5", action="store_true", help="also run the n=96 block-distributed chain at fewer than 8 GPUs")
    ap.add_argument("--beta", type=float, default=None, help="pCN step size of the bench chains (default BENCH_BETA)")
    ap.add_argument("--chains-per-gpu", type=int, default=1,
                    help="config 4 (many independent chains): also report the throughput with this many chains in flight per GPU")
    args = ap.parse_args()
    if args.beta is not None:
        global BENCH_BETA
        BENCH_BETA = args.beta
    reserve_stdout()
    cfg = CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, cfg, rank, world)
        return

    import torch
    import torch.distributed as dist
    from mspde import image as im
    from mspde import util, _ffi

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep NCCL's version banner off stdout: one JSON line only
        # a rank lost mid-collective should end the run in minutes, not after NCCL's default 10-minute watchdog
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    f_iter, N, M = flops_iter(**cfg)
    FP64_PEAK_TFLOPS, fp64_peak_source = measure_fp64_peak(_ffi)
    if args.config == 5:
        # BASELINE configs[4]: a single n=96 chain whose dense work is block-distributed over the GPUs (strong scaling)
        if world < 2:
            raise SystemExit("config 5 (n=96) needs at least 2 GPUs: python -m torch.distributed.run ... bench.py --config 5")
        if args.gemm is not None:
            _ffi.set_gemm_mode(_ffi.GEMM_3M if args.gemm == "3m" else _ffi.GEMM_4M)
        blk = config5_block(cfg, dev, rank, world, steps=args.steps, warmup=min(args.warmup, 1), fp64_peak=FP64_PEAK_TFLOPS,
                            selfcheck_n=(32,), local_rank=local_rank)
        if rank == 0:
            line = {"metric": "pCN iters/sec (n=96, 3 layers, one chain block-distributed)", "value": blk.get("value"), "unit": "iters/s",
                    "n_gpus": world, "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": blk.get("ms_per_step"),
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
            line.update({k: v for k, v in blk.items() if k not in ("value", "ms_per_step")})
            emit(line)
        dist.destroy_process_group()
        return
    K = cfg["n_layers"]
    t_setup = time.perf_counter()
    sim = im.Simulation(n_layers=K, n_samples=args.steps * 2 + args.warmup * 2 + 8, n=cfg["n"], n_extended=cfg["n_ext"], beta=BENCH_BETA,
                        kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4, meas_std=0.2, evaluation_interval=10,
                        printProgress=False, seed=1 + rank, burn_percentage=25.0, enable_step_adaptation=True,
                        pcn_variant="dunlop", phantom_name="shepp.png", meas_type="tomo", n_theta=cfg["n_theta"],
                        verbose=False, device=local_rank)
    sim.pcn.set_chol_epsilon(1e-6)
    pcn, ctx, Layers = sim.pcn, sim.pcn.ctx, sim.Layers
    make_state_consistent(Layers)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    nr = sim.fourier.basis_number_2D_ravel

    # ---- injected noise for every step, generated once on the host (pinned) in the reference's RNG call order
    rng = np.random.Generator(np.random.PCG64(1000 + rank))
    total = 2 * (args.warmup + args.steps)

    def host_noise():
        w = [util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr)) for _ in range(K - 1)]
        log_u = float(np.log(rng.random()))
        w_last = util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))
        e = rng.standard_normal(M)
        return ([torch.from_numpy(x).pin_memory() for x in w], log_u, torch.from_numpy(w_last).pin_memory(),
                torch.from_numpy(e).pin_memory())
    noises = [host_noise() for _ in range(total)]
    h2d_bytes = sum(x.numel() * x.element_size() for x in noises[0][0]) + noises[0][2].numel() * 16 + noises[0][3].numel() * 8
    d2h_bytes = 4 * 8 + K * nr * 16

    def to_dev(nz):
        return [x.to(dev, non_blocking=True) for x in nz[0]], nz[1], nz[2].to(dev, non_blocking=True), nz[3].to(dev, non_blocking=True)

    lib = _ffi.lib()
    lib.mspde_launch_count.restype = __import__("ctypes").c_longlong
    if args.gemm is not None:
        _ffi.set_gemm_mode(_ffi.GEMM_3M if args.gemm == "3m" else _ffi.GEMM_4M)
    gemm_mode = _ffi.get_gemm_mode()
    executed = 0.75 if gemm_mode == _ffi.GEMM_3M else 1.0     # DMMA flops executed per conventional (8-flop) GEMM flop

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- arm 1: device-resident inputs, no host sync inside the loop -------------------------------------------
    pcn._ensure_inputs(Layers[0].stdev_sym_host)
    cb = pcn._bind(Layers)
    dev_noise = [to_dev(nz) for nz in noises[:args.warmup + args.steps]]
    rec = torch.zeros((args.steps + args.warmup, K, nr), dtype=torch.complex128, device=dev)
    outs = torch.zeros((args.steps + args.warmup, 4), dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    for i in range(args.warmup):
        w, lu, wl, e = dev_noise[i]
        ctx.pcn_step(cb, w, lu, wl, e, record=rec[i], out=outs[i])
    ctx.check_info("warm-up")
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    launches0 = lib.mspde_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for i in range(args.warmup, args.warmup + args.steps):
        w, lu, wl, e = dev_noise[i]
        ctx.pcn_step(cb, w, lu, wl, e, record=rec[i], out=outs[i])
    if world > 1:
        gathered = [torch.empty_like(rec) for _ in range(world)]
        dist.all_gather(gathered, rec)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    launches = lib.mspde_launch_count() - launches0
    ctx.check_info("timed region")
    ms_dev = ev0.elapsed_time(ev1)
    t = torch.tensor([ms_dev], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev = float(t.item())
    value = world * args.steps / (ms_dev * 1e-3)
    acc_rate = float(outs[args.warmup:, 0].mean().item())

    # ---- arm 2: end to end through the public API, host noise in pinned memory, results read back every step ----
    base = args.warmup + args.steps
    for i in range(min(args.warmup, 2)):
        pcn.one_step(Layers, noise=to_dev(noises[base + i]))
    barrier()
    t0 = time.perf_counter()
    for i in range(args.warmup, args.warmup + args.steps):
        pcn.one_step(Layers, noise=to_dev(noises[base + i]))
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t = torch.tensor([t_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    t_e2e = float(t.item())
    e2e_value = world * args.steps / t_e2e

    extra = {}
    if args.cache_phi:
        pcn.cache_phi_current = True
        pcn.one_step(Layers, noise=to_dev(noises[0]))
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pcn.one_step(Layers, noise=to_dev(noises[1 + i]))
        torch.cuda.synchronize()
        extra["e2e_cached_phi_current"] = {"value": world * args.steps / (time.perf_counter() - t0), "unit": "iters/s",
                                           "note": "results-preserving: Phi(L_cur) reused until the next acceptance"}
        pcn.cache_phi_current = False

    if not args.no_extras:
        # The same step with the other complex product scheme of the GEMM kernel (4M = 8 flop per complex multiply-add)
        other = _ffi.GEMM_4M if gemm_mode == _ffi.GEMM_3M else _ffi.GEMM_3M
        _ffi.set_gemm_mode(other)
        pcn.one_step(Layers, noise=to_dev(noises[0]))
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pcn.one_step(Layers, noise=to_dev(noises[1 + i]))
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        extra["e2e_gemm_4m" if other == _ffi.GEMM_4M else "e2e_gemm_3m"] = {
            "value": world * args.steps / float(tt.item()), "unit": "iters/s",
            "note": "same step, GEMM kernel switched to the other complex product scheme (mspde_set_gemm_mode)"}
        _ffi.set_gemm_mode(gemm_mode)
        # Results-preserving reformulation, reported SEPARATELY (never the headline): determinant-lemma Phi, no M-sized work
        pcn.fast_phi = True
        pcn.one_step(Layers, noise=to_dev(noises[0]))
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            pcn.one_step(Layers, noise=to_dev(noises[1 + i]))
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        extra["e2e_determinant_lemma_phi"] = {"value": world * args.steps / float(tt.item()), "unit": "iters/s",
                                              "note": "opt-in pCN.fast_phi: same Phi to <=1e-10 (tests), 4N^3+8/3N^3 flops per Phi "
                                                      "instead of F_Phi; NOT the reference-faithful headline"}
        pcn.fast_phi = False

    if args.chains_per_gpu > 1:
        # BASELINE configs[3]: independent chains sharded over the GPUs.  Several chains per GPU are kept in flight on their
        # own streams (each with its own context), which fills the latency-bound gaps of one chain with the GEMMs of another.
        C = args.chains_per_gpu
        sims = [sim] + [im.Simulation(n_layers=K, n_samples=8, n=cfg["n"], n_extended=cfg["n_ext"], beta=BENCH_BETA, kappa=5e9, sigma_0=4e7,
                                      sigma_v=3.2e4, sigma_scaling=0.4, meas_std=0.2, evaluation_interval=10, printProgress=False,
                                      seed=100 + rank * C + c, burn_percentage=25.0, enable_step_adaptation=True,
                                      pcn_variant="dunlop", phantom_name="shepp.png", meas_type="tomo", n_theta=cfg["n_theta"],
                                      verbose=False, device=local_rank) for c in range(1, C)]
        streams = [torch.cuda.Stream(device=dev) for _ in range(C)]
        binds = []
        for sm_ in sims:
            sm_.pcn.set_chol_epsilon(1e-6)
            sm_.pcn._ensure_inputs(sm_.Layers[0].stdev_sym_host)
            binds.append(sm_.pcn._bind(sm_.Layers))
        torch.cuda.synchronize()

        def sweep(i):
            for c in range(C):
                with torch.cuda.stream(streams[c]):
                    w, lu, wl, e = dev_noise[i % len(dev_noise)]
                    sims[c].pcn.ctx.pcn_step(binds[c], w, lu, wl, e)
        for i in range(2):
            sweep(i)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            sweep(i)
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        for sm_ in sims:
            sm_.pcn.ctx.check_info("multi-chain run")
        extra["multi_chain"] = {"chains_per_gpu": C, "value": world * C * args.steps / float(tt.item()), "unit": "iters/s",
                                "note": "aggregate over all chains; independent chains, reference-faithful formulation"}

    # ---- roofline of the dominant kernel: zgemm_tn_kernel on its largest launch, Q_inv = I - Ht^H Ht (upper) ----
    roofline = None
    if rank == 0:
        X = ctx.buffer("X", N, M)
        Q = ctx.buffer("Q", M, M)
        reps = 3
        ctx.gemm_tn(X, X, conjA=True, alpha=-1.0, beta=0.0, C=Q, upper=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            ctx.gemm_tn(X, X, conjA=True, alpha=-1.0, beta=0.0, C=Q, upper=True)
        e1.record()
        torch.cuda.synchronize()
        ms_k = e0.elapsed_time(e1) / reps
        fl8 = 4.0 * M * M * N         # ZHERK m x m x k = 4 m^2 k real flops in the 8-flop convention
        fl = executed * fl8           # what this kernel's algorithm needs and the DMMA pipe executes (3M: 6 of every 8)
        ach = fl / (ms_k * 1e-3) / 1e12
        step_rate = f_iter / (ms_dev * 1e-3 / args.steps) / 1e12
        scheme = "3M: three real DMMA products per complex product" if gemm_mode == _ffi.GEMM_3M else "4M: four real DMMA products"
        packed = gemm_mode == _ffi.GEMM_3M and 0 < _ffi.get_gemm_pre_min() <= min(M, N)
        kname = ("zgemm3m_pre_tn_kernel + 2 x pack_operand_kernel (operands repacked into the stage-tile image, T3 planes pre-summed)"
                 if packed else ("zgemm3m_tn_kernel" if gemm_mode == _ffi.GEMM_3M else "zgemm_tn_kernel"))
        if packed:
            traffic, traffic_src = 23.6e9, ("profiles/r02_zgemm3m_packed_full.ncu-rep: zgemm3m_pre_tn_kernel 19.00e9 read + 1.04e9 written "
                                            "(44.60 ms, tensor pipe 95.5 % active, L2 hit 79.9 %) + two pack passes of 0.73e9 read + 1.04e9 "
                                            "written each (0.27 ms each); algorithmic 0.73e9 + 1.05e9; DRAM runs at ~6 % of its peak")
        elif gemm_mode == _ffi.GEMM_3M:
            traffic, traffic_src = 15.44e9, ("profiles/r02_zgemm3m_full.ncu-rep (in-loop sums: 14.40e9 read + 1.05e9 written, tensor pipe "
                                             "88.6 % active, L2 hit 85.5 %)")
        else:
            traffic, traffic_src = 13.1e9, "profiles/r01_zgemm_full_v2.ncu-rep (grouped tile rasterisation; 56.2e9 before it)"
        roofline = {"kernel": "%s (DMMA m8n8k4, %s), launch Q_inv = I - Ht^H Ht, upper triangle, m=n=%d k=%d" % (kname, scheme, M, N),
                    "bound": "tensor", "achieved": ach, "peak": FP64_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_PEAK_TFLOPS,
                    "flops_per_launch": fl, "zgemm_equivalent": {"flops_per_launch": fl8, "achieved": fl8 / (ms_k * 1e-3) / 1e12,
                                                                 "frac": fl8 / (ms_k * 1e-3) / 1e12 / FP64_PEAK_TFLOPS,
                                                                 "note": "same launch counted at 8 flop per complex multiply-add (SURVEY 8d's "
                                                                         "ZHERK = 4 n^2 k), the convention ZGEMM rates are quoted in; above 1.0 "
                                                                         "under 3M because the kernel's algorithm needs 6"},
                    "traffic": traffic,
                    "traffic_source": "dram__bytes_read+write of this launch from one ncu --set full capture: " + traffic_src,
                    "timing_note": "ms_per_launch is the whole mspde_zgemm_tn call (pack passes included when the packed kernel is used)",
                    "ms_per_launch": ms_k,
                    "peak_source": fp64_peak_source,
                    "step_fp64": {"flops_per_step": f_iter, "achieved": step_rate, "frac": step_rate / FP64_PEAK_TFLOPS,
                                  "executed_frac": executed * step_rate / FP64_PEAK_TFLOPS,
                                  "note": "flops_per_step is SURVEY 8(d)'s F_iter (8-flop convention); executed_frac scales it by the "
                                          "share of those flops the GEMM scheme actually issues to the FP64 pipe"}}

    if rank == 0 and roofline is not None:
        # HBM-bound kernels of the path (north_star: achieved HBM GB/s for assembly and misfit), timed live with CUDA events
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak, hbm_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json (of measured)"
        except Exception:
            hbm_peak, hbm_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
        import ctypes as _ct

        def ev_ms(fn, reps=5):
            fn(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                fn()
            b.record(); torch.cuda.synchronize()
            return a.elapsed_time(b) / reps
        tp, tm = ctx.field_coeffs(Layers[1].current_sample_sym)
        Lbuf = Layers[2].LMat.latest_computed_L
        ms_a = ev_ms(lambda: ctx.assemble_L(tp, tm, 1.0e5, out=Lbuf))
        Qv = ctx.buffer("Q", M, M)
        rs, rl = torch.empty(M, dtype=torch.float64, device=dev), torch.empty(M, dtype=torch.float64, device=dev)
        ms_v = ev_ms(lambda: lib.mspde_upper_rows_times_y(_ffi.current_stream(), _ffi.ptr(Qv), Qv.stride(0), M, M, _ffi.ptr(ctx.y),
                                                          _ffi.ptr(rs), _ffi.ptr(rl)))
        roofline["hbm_kernels"] = [
            {"kernel": "assemble_L_kernel (a3+a4, one write pass of L)", "bound": "hbm", "bytes": 16.0 * N * N,
             "achieved": 16.0 * N * N / ms_a * 1e-6, "peak": hbm_peak, "unit": "GB/s", "frac": 16.0 * N * N / ms_a * 1e-6 / hbm_peak,
             "peak_source": hbm_src},
            {"kernel": "vy_rows_kernel (misfit 0.5*||y@C||^2 - sum log diag C over the upper M x M factor)", "bound": "hbm",
             "bytes": 8.0 * M * M, "achieved": 8.0 * M * M / ms_v * 1e-6, "peak": hbm_peak, "unit": "GB/s",
             "frac": 8.0 * M * M / ms_v * 1e-6 / hbm_peak, "peak_source": hbm_src}]

    library = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            ms_lib = library_standin_ms(sim, to_dev(noises[0]))
            library = {"ms_per_step": ms_lib, "value": 1e3 / ms_lib, "unit": "iters/s", "speedup_of_ours": ms_lib / (ms_dev / args.steps),
                       "what": "the same iteration through library calls only (torch.fft = cuFFT, @ = cuBLAS ZGEMM, torch.linalg = "
                               "cuSOLVER), same inputs, same process: the stand-in for the reference's CuPy path on this box"}
        except Exception as ex:
            library = {"unavailable": repr(ex)[:200]}
        torch.cuda.empty_cache()

    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, desc, cores, dt, done = cpu_oracle_rate(cfg, steps=1, warmup=0, budget_s=30.0)
        cpu_baseline = {"value": rate, "unit": "iters/s", "cores": cores, "kind": "port", "sample": desc}

    def headline(config4, config5):
        line = {"metric": f"pCN iters/sec (n={cfg['n']}, {cfg['n_layers']} layers)", "value": value, "unit": "iters/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(cfg, world),
                "e2e": {"value": e2e_value, "unit": "iters/s", "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": d2h_bytes},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline,
                "library_standin": library, "config4_64_chains": config4, "config5_n96_distributed": config5,
                "gemm_scheme": "3m" if gemm_mode == _ffi.GEMM_3M else "4m", "acceptance_rate": acc_rate, "setup_s": t_setup}
        line.update(extra)
        return line

    config4 = config5 = None
    if world > 1 and args.config == 2 and not args.no_extras:
        # the two multi-GPU splits north_star names, measured in the same driver-visible run
        sim.pcn.ctx.close()
        del sim, pcn, ctx, Layers, cb, dev_noise, rec, outs
        _ffi.release_gemm_workspaces()
        torch.cuda.empty_cache()
        lost = {"unavailable": "rank 0 did not return from this block (another rank stalled or failed); the rest of the line was "
                               "measured before it started"}
        def arm(make_line):
            try:
                return LineGuard(make_line()) if rank == 0 else None
            except Exception as ex:                     # the insurance must never cost the run
                sys.stderr.write(f"line guard unavailable: {ex!r}\n")
                return None
        guard = arm(lambda: headline(dict(lost), dict(lost)))
        try:
            try:
                config4 = config4_block(cfg, dev, rank, world, local_rank)
            except Exception as ex:                     # keep the headline line even if an extra block fails
                config4 = {"unavailable": repr(ex)[:300]}
            if world >= 8 or args.with_config5:
                if rank == 0:
                    if guard is not None:
                        guard.disarm()
                    guard = arm(lambda: headline(config4, dict(lost)))
                try:
                    config5 = config5_block(CONFIGS[5], dev, rank, world, steps=2, warmup=1, fp64_peak=FP64_PEAK_TFLOPS,
                                            selfcheck_n=(32,), local_rank=local_rank)
                except Exception as ex:
                    config5 = {"unavailable": repr(ex)[:300]}
        finally:
            if guard is not None:
                guard.disarm()

    if rank == 0:
        emit(headline(config4, config5))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
