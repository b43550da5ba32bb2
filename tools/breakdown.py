"""Per-phase device timings of one pCN step at full size (dev tool)."""
import sys, os, time, ctypes
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde import image as im, _ffi
n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
nth = int(sys.argv[2]) if len(sys.argv) > 2 else 90
sim = im.Simulation(n_layers=3, n_samples=16, n=n, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4,
                    meas_std=0.2, evaluation_interval=10, printProgress=False, seed=1, burn_percentage=25.0, enable_step_adaptation=True,
                    pcn_variant="dunlop", phantom_name="shepp.png", n_theta=nth)
sim.pcn.set_chol_epsilon(1e-6)
pcn, ctx, Layers = sim.pcn, sim.pcn.ctx, sim.Layers
pcn._ensure_inputs(Layers[0].stdev_sym_host)
N, M = ctx.N, ctx.M
lib = _ffi.lib()
def t_ms(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
L = Layers[2].LMat.current_L; u = Layers[1].current_sample_sym; w = Layers[1].current_noise_sample
print("N", N, "M", M)
print("construct_L  %.3f ms  (%.1f GB/s of 16N^2)" % ((t := t_ms(lambda: ctx.construct_L(u, 1e5, out=Layers[2].LMat.latest_computed_L if Layers[2].LMat.latest_computed_L is not L else None))), 16 * N * N / t * 1e-6))
tp, tm = ctx.field_coeffs(u)
print("field_coeffs %.3f ms" % t_ms(lambda: ctx.field_coeffs(u)))
Lbuf = ctx.new_matrix(N, N)
print("assemble_L   %.3f ms  (%.1f GB/s)" % ((t := t_ms(lambda: ctx.assemble_L(tp, tm, 1e5, out=Lbuf))), 16 * N * N / t * 1e-6))
print("phi          %.3f ms" % t_ms(lambda: lib.mspde_phi(ctx._h, _ffi.ptr(L), L.stride(0), _ffi.ptr(ctx.scal))))
print("solve (QR)   %.3f ms" % t_ms(lambda: ctx.solve(L, w)))
rhs = torch.randn(M + N, dtype=torch.complex128, device=ctx.device)
print("gibbs lstsq  %.3f ms" % t_ms(lambda: ctx.gibbs_lstsq(L, rhs)))
# inside phi
r = ctx.buffer("r", N, N); X = ctx.buffer("X", N, M); Q = ctx.buffer("Q", M, M)
st = _ffi.current_stream()
winv = torch.empty(((max(N, M) + 63) // 64) * 64 * 64, dtype=torch.complex128, device=ctx.device)
info = torch.zeros(4, dtype=torch.int32, device=ctx.device)
print("  herk N     %.3f ms" % t_ms(lambda: ctx.gemm_tn(L, L, True, 1.0, 1.0, C=r, upper=True)))
def mk_r():
    r.copy_(ctx.HtH); r.diagonal().add_(ctx.delta); ctx.gemm_tn(L, L, True, 1.0, 1.0, C=r, upper=True)
mk_r(); rsave = r.clone()
def potrf_N():
    r.copy_(rsave); lib.mspde_zpotrf_upper(st, _ffi.ptr(r), r.stride(0), N, _ffi.ptr(winv), _ffi.ptr(info))
print("  copy N     %.3f ms" % t_ms(lambda: r.copy_(rsave)))
print("  potrf N+cp %.3f ms" % t_ms(potrf_N))
def trsm():
    X.copy_(ctx.Hct); lib.mspde_ztrsm_uh(st, _ffi.ptr(r), r.stride(0), _ffi.ptr(winv), N, _ffi.ptr(X), X.stride(0), M)
print("  copy X     %.3f ms" % t_ms(lambda: X.copy_(ctx.Hct)))
print("  trsm+cp    %.3f ms" % t_ms(trsm))
print("  herk M     %.3f ms" % t_ms(lambda: ctx.gemm_tn(X, X, True, -1.0, 0.0, C=Q, upper=True)))
Q.diagonal().add_(1.0); Qsave = Q.clone()
def potrf_M():
    Q.copy_(Qsave); lib.mspde_zpotrf_upper(st, _ffi.ptr(Q), Q.stride(0), M, _ffi.ptr(winv), _ffi.ptr(info))
print("  copy Q     %.3f ms" % t_ms(lambda: Q.copy_(Qsave)))
print("  potrf M+cp %.3f ms" % t_ms(potrf_M))
print("info", info.tolist())
