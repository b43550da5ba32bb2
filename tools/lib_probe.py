"""On-box library stand-in timings (cuBLAS/cuSOLVER through torch): FP64 GEMM peak and the
torch.linalg rendition of the dense pieces of one pCN step.  Not part of the product path."""
import json, sys, time
import torch

def t_ms(f, reps=3):
    f(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best

out = {}
dev = "cuda"
for n in (4096, 8192):
    a = torch.randn(n, n, dtype=torch.float64, device=dev); b = torch.randn(n, n, dtype=torch.float64, device=dev)
    ms = t_ms(lambda: a @ b)
    out[f"dgemm_{n}_tflops"] = 2 * n**3 / ms * 1e-9
    ac = torch.randn(n, n, dtype=torch.complex128, device=dev); bc = torch.randn(n, n, dtype=torch.complex128, device=dev)
    ms = t_ms(lambda: ac @ bc)
    out[f"zgemm_{n}_tflops"] = 8 * n**3 / ms * 1e-9
    del a, b, ac, bc
N, M = 3969, 11430
L = torch.randn(N, N, dtype=torch.complex128, device=dev)
H = torch.randn(M, N, dtype=torch.complex128, device=dev) / 50
ms = t_ms(lambda: L.conj().T @ L); out["LhL_ms"] = ms; out["LhL_tflops"] = 8 * N**3 / ms * 1e-9
r = L.conj().T @ L + H.conj().T @ H + torch.eye(N, device=dev, dtype=torch.complex128)
ms = t_ms(lambda: torch.linalg.cholesky(r)); out["potrf_N_ms"] = ms; out["potrf_N_tflops"] = 4 / 3 * N**3 / ms * 1e-9
c = torch.linalg.cholesky(r)
Hh = H.conj().T.contiguous()
ms = t_ms(lambda: torch.linalg.solve_triangular(c, Hh, upper=False)); out["trsm_ms"] = ms; out["trsm_tflops"] = 4 * N * N * M / ms * 1e-9
Ht = torch.linalg.solve_triangular(c, Hh, upper=False)
ms = t_ms(lambda: Ht.conj().T @ Ht); out["gram_M_ms"] = ms; out["gram_M_tflops"] = 8 * M * M * N / ms * 1e-9
Q = torch.eye(M, device=dev, dtype=torch.complex128) - Ht.conj().T @ Ht
ms = t_ms(lambda: torch.linalg.cholesky(Q)); out["potrf_M_ms"] = ms; out["potrf_M_tflops"] = 4 / 3 * M**3 / ms * 1e-9
del Q
w = torch.randn(N, 1, dtype=torch.complex128, device=dev)
ms = t_ms(lambda: torch.linalg.solve(L, w)); out["lu_solve_ms"] = ms; out["lu_tflops"] = 8 / 3 * N**3 / ms * 1e-9
LBar = torch.cat([H, L], 0)
ms = t_ms(lambda: torch.linalg.qr(LBar), reps=2); out["qr_ms"] = ms; out["qr_tflops_2x"] = 2 * (8 * (M + N) * N * N - 8 / 3 * N**3) / ms * 1e-9
print(json.dumps(out, indent=1))
