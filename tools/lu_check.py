import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from mspde.core import MspdeContext
torch.manual_seed(0)
def t_ms(f, reps=3):
    f(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best
for n in (2, 3, 4, 5, 8, 12, 32):
    ctx = MspdeContext(n, 64, 256, 3); N = ctx.N
    ctx.set_inputs(np.zeros((256, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(256), np.ones(N), np.ones(N), 0.0)
    for kind in ("random", "diagdom"):
        L = torch.randn(N, N, dtype=torch.complex128, device="cuda")
        if kind == "diagdom": L += 2 * N ** 0.5 * torch.eye(N, dtype=torch.complex128, device="cuda")
        w = torch.randn(N, dtype=torch.complex128, device="cuda")
        Lp = ctx.new_matrix(N, N); Lp.copy_(L)
        x, ld = ctx.solve(Lp, w, want_logdet=True)
        xq, ldq = ctx.solve(Lp, w, want_logdet=True, method="qr")
        ref = torch.linalg.solve(L, w); ldr = float(torch.linalg.slogdet(L)[1])
        print(f"n={n} N={N} {kind}: lu rel {float((x-ref).abs().max()/ref.abs().max()):.2e} qr rel {float((xq-ref).abs().max()/ref.abs().max()):.2e} logdet {ld:.10f} {ldq:.10f} ref {ldr:.10f}")
    if n == 32:
        print("LU solve ms", t_ms(lambda: ctx.solve(Lp, w)), "QR solve ms", t_ms(lambda: ctx.solve(Lp, w, method="qr")))
    # singular
    if n == 4:
        Ls = Lp.clone(); Ls[:, 5] = 0
        try:
            ctx.solve(Ls, w, want_logdet=True); print("singular: no error?!")
        except np.linalg.LinAlgError as e: print("singular ->", e)
    ctx.close()
