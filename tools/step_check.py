"""Dev check: QR solve / lstsq / full pCN steps on the GPU vs the numpy oracle at small n."""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import numpy as np, torch
from oracle import mspde_oracle as o
from mspde.core import MspdeContext, ChainBuffers
n, nth, nsteps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
hp = o.Hyper(n_layers=3, n=n, n_ext=64, n_theta=nth)
pb = o.synthetic_problem(hp); ft = pb['ft']; H = pb['H']; HtH = pb['HtH']; y = pb['y']; delta = pb['delta']
pt = o.prior_tables(hp, ft); K = 3; N = ft.N; M = H.shape[0]
ns = o.NoiseStream(7, 3, ft.n_ravel, M)
st = o.init_chain(hp, ft, [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
ctx = MspdeContext(n, 64, M, 3)
ctx.set_inputs(H, HtH, y, ft.D_diag, pt['stdev_sym'], delta)
rel = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))
# solve + logdet
L = st.L_cur[1]; w = st.noise_cur[1]
Ld = ctx.new_matrix(N, N); Ld.copy_(ctx.to_dev(L))
x, ld = ctx.solve(Ld, ctx.to_dev(w), want_logdet=True)
print("solve rel", rel(x.cpu().numpy(), np.linalg.solve(L, w)), "logdet", ld, o.logdet_L(L), "cond", np.linalg.cond(L))
# gibbs lstsq
rng = np.random.default_rng(0); rhs = rng.standard_normal(M + N) + 1j * rng.standard_normal(M + N)
L2 = st.L_cur[2]; L2d = ctx.new_matrix(N, N); L2d.copy_(ctx.to_dev(L2))
v = ctx.gibbs_lstsq(L2d, ctx.to_dev(rhs))
vo = o.lstsq_qr(np.vstack((H, L2)), rhs)
print("gibbs lstsq rel", rel(v.cpu().numpy(), vo))
# chain
cb = ChainBuffers(ctx, st.sqrt_betas, st.beta)
for k in range(K):
    cb.noise_cur[k].copy_(ctx.to_dev(st.noise_cur[k])); cb.noise_new[k].copy_(ctx.to_dev(st.noise_new[k]))
    cb.u_cur[k].copy_(ctx.to_dev(st.u_cur_sym[k])); cb.u_new[k].copy_(ctx.to_dev(st.u_new_sym[k]))
    if k >= 1:
        cb.L_cur[k].copy_(ctx.to_dev(st.L_cur[k])); cb.L_latest[k].copy_(ctx.to_dev(st.L_latest[k]))
rec = torch.zeros((K, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
worst = 0.0
for it in range(nsteps):
    nz = ns.draw_iteration()
    d = o.one_step(st, ft, H, HtH, y, delta, nz)
    cb.set_step(st.beta)
    out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz['w_half']], nz['log_u'], ctx.to_dev(nz['w_last']), ctx.to_dev(nz['e']), record=rec)
    ctx.check_info("step")
    g = out.cpu().numpy()
    if (it + 1) % 5 == 0: o.adapt_step(st, hp)
    halves = o.current_halves(st, ft)
    rr = max(rel(rec[k].cpu().numpy(), halves[k]) for k in range(K))
    ru = max(rel(cb.u_new[k].cpu().numpy(), st.u_new_sym[k]) for k in range(K))
    worst = max(worst, rr, ru, abs(g[2] - d['phi_cur']) / abs(d['phi_cur']), abs(g[3] - d['phi_new']) / abs(d['phi_new']))
    print(it, "acc", int(g[0]), d['accepted'], "phi_cur rel %.2e phi_new rel %.2e" % (abs(g[2] - d['phi_cur']) / abs(d['phi_cur']), abs(g[3] - d['phi_new']) / abs(d['phi_new'])),
          "rec rel %.2e u_new rel %.2e" % (rr, ru), "margin %.3g" % abs(d['log_ratio'] - d['log_u']))
    assert int(g[0]) == int(d['accepted'])
print("WORST", worst)
