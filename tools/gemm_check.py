"""GPU check + timing of mspde_zgemm_tn against torch (cuBLAS) -- development tool."""
import ctypes, sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "mcmc-multispde_b200"))
import torch
from mspde import _ffi
L = _ffi.lib()
dev = "cuda"
torch.manual_seed(0)

def gemm(conjA, A, B, C, alpha=1.0, beta=0.0, uplo=0, ws=None, nsplit=1):
    k, m = A.shape; n = B.shape[1]
    rc = L.mspde_zgemm_tn(_ffi.current_stream(), int(conjA), m, n, k, ctypes.c_double(alpha), _ffi.ptr(A), A.stride(0),
                          _ffi.ptr(B), B.stride(0), ctypes.c_double(beta), _ffi.ptr(C), C.stride(0), uplo, _ffi.ptr(ws), nsplit)
    _ffi.check(rc, "zgemm_tn")

def rnd(*s): return torch.randn(*s, dtype=torch.complex128, device=dev)

ok = True
for (m, n, k, conj, uplo, beta) in [(64, 64, 16, 1, 0, 0.0), (64, 64, 64, 0, 0, 0.0), (100, 77, 35, 1, 0, 0.0), (200, 130, 129, 1, 0, 1.0),
                                    (333, 333, 211, 1, 1, 1.0), (49, 1016, 49, 1, 0, -1.0), (961, 961, 961, 1, 1, 0.0), (64, 500, 3001, 1, 0, 0.0)]:
    A = rnd(k, m); B = rnd(k, n) if not uplo else A; C = rnd(m, n); C0 = C.clone()
    ref = (A.conj().T if conj else A.T) @ B
    alpha = -1.0 if beta else 1.0
    ref = alpha * ref + beta * C0
    gemm(conj, A, B, C, alpha, beta, uplo)
    torch.cuda.synchronize()
    if uplo:
        mask = torch.triu(torch.ones(m, n, dtype=torch.bool, device=dev))
        err = ((C - ref)[mask]).abs().max().item(); low = ((C - C0)[~mask]).abs().max().item()
    else:
        err = (C - ref).abs().max().item(); low = 0.0
    print(f"m{m} n{n} k{k} conj{conj} uplo{uplo} beta{beta}: maxerr {err:.3e} lower-untouched {low:.1e} scale {ref.abs().max().item():.2e}")
    ok &= err < 1e-10 * max(1.0, ref.abs().max().item()) and low == 0.0
# split-K
A = rnd(3001, 64); B = rnd(3001, 500); C = rnd(64, 500); ws = torch.empty(8 * 64 * 500, dtype=torch.complex128, device=dev)
gemm(1, A, B, C, 1.0, 0.0, 0, ws, 8); torch.cuda.synchronize()
err = (C - A.conj().T @ B).abs().max().item(); print("splitk err", err); ok &= err < 1e-10
print("CHECK", "PASS" if ok else "FAIL")

def t_ms(f, reps=5):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for (m, n, k, uplo) in [(3969, 3969, 3969, 0), (3969, 3969, 3969, 1), (11430, 11430, 3969, 1), (3969, 11430, 2048, 0), (8192, 8192, 128, 0), (8192, 8192, 64, 0), (5000, 5000, 128, 1)]:
    A = rnd(k, m); B = A if uplo else rnd(k, n); C = torch.zeros(m, n, dtype=torch.complex128, device=dev)
    ms = t_ms(lambda: gemm(1, A, B, C, -1.0, 1.0, uplo))
    fl = 8.0 * m * n * k * (0.5 if uplo else 1.0)
    print(f"time m{m} n{n} k{k} uplo{uplo}: {ms:.3f} ms  {fl / ms * 1e-9:.2f} TFLOP/s")
    del A, B, C

# ---- 3M: pre-summed operand planes (P kernel) against the in-loop sums: identical arithmetic -> identical bits, and timing
if L.mspde_get_gemm_mode() == 1:
    default_pre = L.mspde_get_gemm_pre_min()
    same = True
    for (m, n, k, conj, uplo, beta) in [(65, 33, 17, 1, 0, 0.0), (127, 63, 35, 0, 0, 1.0), (333, 333, 211, 1, 1, 1.0), (961, 961, 961, 1, 1, 0.0),
                                        (200, 1031, 129, 1, 0, -0.5), (1, 1, 1, 1, 0, 0.0), (64, 500, 3001, 1, 0, 0.0)]:
        A = rnd(k, m); B = rnd(k, n) if not uplo else A; C0 = rnd(m, n)
        outs = []
        for pre in (0, 1):
            L.mspde_set_gemm_pre_min(pre)
            C = C0.clone(); gemm(conj, A, B, C, 0.75, beta, uplo); torch.cuda.synchronize(); outs.append(C)
        eq = torch.equal(outs[0], outs[1]); same &= eq
        print(f"pre vs in-loop m{m} n{n} k{k} conj{conj} uplo{uplo} beta{beta}: bit-identical {eq}")
    print("PRE CHECK", "PASS" if same else "FAIL")
    for (m, n, k, uplo) in [(11430, 11430, 3969, 1), (3969, 3969, 3969, 0), (3969, 11430, 2048, 0), (10000, 10000, 512, 1), (6000, 6000, 512, 1),
                            (3000, 3000, 512, 1), (1536, 1536, 512, 1), (8192, 8192, 128, 0), (8192, 8192, 64, 0), (2048, 2048, 2048, 0)]:
        A = rnd(k, m); B = A if uplo else rnd(k, n); C = torch.zeros(m, n, dtype=torch.complex128, device=dev)
        fl = 8.0 * m * n * k * (0.5 if uplo else 1.0)
        res = []
        for pre in (0, 1):
            L.mspde_set_gemm_pre_min(pre)
            ms = t_ms(lambda: gemm(1, A, B, C, -1.0, 1.0, uplo))
            res.append(ms)
        print(f"pre-timing m{m} n{n} k{k} uplo{uplo}: in-loop {res[0]:.3f} ms {fl / res[0] * 1e-9:.2f} TF | pre-summed {res[1]:.3f} ms {fl / res[1] * 1e-9:.2f} TF | gain {100 * (res[0] / res[1] - 1):+.1f} %")
        del A, B, C
    L.mspde_set_gemm_pre_min(default_pre)
