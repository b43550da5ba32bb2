"""Dev check: assembly + Phi on the GPU vs the numpy oracle at small n."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import mspde_oracle as o
from mspde.core import MspdeContext
n, nth = int(sys.argv[1]), int(sys.argv[2])
hp = o.Hyper(n_layers=3, n=n, n_ext=64, n_theta=nth)
pb = o.synthetic_problem(hp); ft = pb['ft']; H = pb['H']; HtH = pb['HtH']; y = pb['y']; delta = pb['delta']
pt = o.prior_tables(hp, ft)
ns = o.NoiseStream(7, 3, ft.n_ravel, H.shape[0])
st = o.init_chain(hp, ft, [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
ctx = MspdeContext(n, 64, H.shape[0], 3)
ctx.set_inputs(H, HtH, y, ft.D_diag, pt['stdev_sym'], delta)
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
for k in (1, 2):
    u = st.u_cur_sym[k - 1]; sb = st.sqrt_betas[k]
    tp, tm = o.field_coeffs(u, ft)
    ud = ctx.to_dev(u)
    tpd, tmd = ctx.field_coeffs(ud)
    print(f"layer {k}: coef+ rel {rel(tpd.cpu().numpy(), tp):.2e} coef- rel {rel(tmd.cpu().numpy(), tm):.2e}")
    Lo = o.assemble_L_from_tables(tp, tm, ft.D_diag, sb)
    Ld = ctx.assemble_L(ctx.to_dev(tp), ctx.to_dev(tm), sb)
    print("  L from oracle tables bit-exact:", bool((Ld.cpu().numpy() == Lo).all()), "rel", rel(Ld.cpu().numpy(), Lo))
    Ld2 = ctx.construct_L(ud, sb)
    print("  L end-to-end rel", rel(Ld2.cpu().numpy(), st.L_cur[k]))
L = st.L_cur[2]
po, parts = o.phi_woodbury(L, H, HtH, y, delta, return_parts=True)
Ld = ctx.new_matrix(ft.N, ft.N); Ld.copy_(ctx.to_dev(L))
t = time.time(); pg = ctx.phi(Ld, parts=True); print("gpu phi time", time.time() - t)
print("phi oracle", po, parts['quad'], parts['logdet'])
print("phi gpu   ", pg, "rel", abs(pg[0] - po) / abs(po))
N, M = ft.N, H.shape[0]
U = torch.triu(ctx.buffer("r", N, N)).cpu().numpy()
print("  c rel", rel(U.conj().T, parts['c']))
print("  Ht rel", rel(ctx.buffer("X", N, M).cpu().numpy(), parts['Ht']))
V = torch.triu(ctx.buffer("Q", M, M)).cpu().numpy()
print("  C rel", rel(V.conj().T, parts['C']))
