"""Generates tests/golden/*.npz in THIS container (needs /root/reference; never runs on the GPU box).

  NUMBA_ENABLE_CUDASIM=1 NUMBA_CACHE_DIR=/tmp/nbcache python oracle/make_golden.py

1. ref_construct_U_n{n}.npz, ref_H_tomography.npz : outputs of the reference's own numba.cuda kernel bodies
   (`_construct_U`, `_calculate_H_Tomography`), AST-extracted from /root/reference/mcmc/util_cupy.py and executed
   verbatim under the numba CUDA simulator.  They pin the BTTB index map bit-exactly and the H entries.
2. ref_util_helpers.npz : symmetrize / symmetrize_2D / extend2D from /root/reference/mcmc/util.py, util_2D.py
   (imported with a stub h5py module).
3. oracle_chain_n4.npz : a 6-step, 3-layer toy chain from the oracle itself (regression fixture; also lets the GPU
   tests run without re-deriving inputs).
"""
import ast
import os
import sys
import types

os.environ.setdefault("NUMBA_ENABLE_CUDASIM", "1")
os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/nbcache")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")

import numpy as np
from numba import cuda

from oracle import mspde_oracle as o


def extract_kernels(names):
    src = open(os.path.join(REF, "mcmc", "util_cupy.py")).read()
    tree = ast.parse(src)
    ns = {"cuda": cuda}
    exec("from math import sin,cos,sqrt,pi\nfrom cmath import exp", ns)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            mod = ast.Module(body=[node], type_ignores=[])
            exec(compile(mod, "util_cupy.py", "exec"), ns)
    return [ns[n] for n in names]


def main():
    os.makedirs(OUT, exist_ok=True)
    construct_U, calc_H = extract_kernels(["_construct_U", "_calculate_H_Tomography"])
    # 1a. BTTB scatter: label every table entry with a unique integer so the gather map can be read back exactly
    for n in (2, 3, 4):
        il = 2 * n - 1
        N = il * il
        labels = (np.arange(il * il).reshape(il, il) + 1).astype(np.complex64)
        U = np.zeros((N, N), dtype=np.complex64)
        bpg = ((N + 7) // 8, (N + 7) // 8)
        construct_U[bpg, (8, 8)](labels, n, il, U)
        np.savez_compressed(os.path.join(OUT, f"ref_construct_U_n{n}.npz"), n=n, labels=labels.real.astype(np.int32),
                            U=U.real.astype(np.int32))
        r, c, mask = o.bttb_index_map(n)
        mine = np.where(mask, labels.real.astype(np.int32)[np.clip(r, 0, il - 1), np.clip(c, 0, il - 1)], 0)
        assert (mine == U.real.astype(np.int32)).all(), "oracle BTTB map differs from the reference kernel"
    # 1b. tomography kernel body on a small geometry (n=3, n_ext=4 -> m=7, 8 angles)
    ft = o.FourierTables(3, 4)
    n_theta, n_r = 8, ft.m
    theta = np.linspace(0.0, 180.0, n_theta, endpoint=False)
    shifted = theta + theta[1]
    r = np.linspace(-0.5, 0.5, n_r, endpoint=False)[::-1]
    tg, rg = np.meshgrid(shifted * np.float64(o.PI32) / 180.0, r)
    ix, iy = ft.ix.ravel("F"), ft.iy.ravel("F")
    H = np.zeros((rg.size, ix.size), dtype=np.complex128)
    bpg = ((H.shape[0] + 31) // 32, (H.shape[1] + 31) // 32)
    calc_H[bpg, (32, 32)](rg.ravel("F"), tg.ravel("F"), ix, iy, H)
    np.savez_compressed(os.path.join(OUT, "ref_H_tomography.npz"), n=3, n_ext=4, n_theta=n_theta, H_kernel=H)
    Ho = o.tomography_H(ft, n_theta)
    Hk = H * (n_r / (2 * (n_r // 2)))
    for b in range(n_theta // 4):
        Hk[b * n_r:(b + 1) * n_r] = np.roll(Hk[b * n_r:(b + 1) * n_r], 1, axis=0)
    print("H closed form vs reference kernel body: max abs diff", np.abs(Ho - Hk).max())
    assert np.abs(Ho - Hk).max() < 1e-14

    # 2. reference CPU helpers
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    sys.path.insert(0, REF)
    import mcmc.util as ru
    import mcmc.util_2D as ru2
    rng = np.random.default_rng(0)
    half = rng.standard_normal(13) + 1j * rng.standard_normal(13)
    uh = rng.standard_normal((5, 3)) + 1j * rng.standard_normal((5, 3))
    sym2 = np.asarray(ru2.symmetrize_2D(uh))
    np.savez_compressed(os.path.join(OUT, "ref_util_helpers.npz"), half=half, sym=np.asarray(ru.symmetrize(half)), uh=uh,
                        sym2D=sym2, ext2D=np.asarray(ru2.extend2D(sym2, 6)))
    assert np.array_equal(o.symmetrize(half), np.asarray(ru.symmetrize(half)))
    assert np.array_equal(o.symmetrize_2D(uh), sym2)
    assert np.array_equal(o.extend2D(sym2, 6), np.asarray(ru2.extend2D(sym2, 6)))

    # 3. oracle toy chain
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8)
    pb = o.synthetic_problem(hp)
    ftc, Hc, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    ns = o.NoiseStream(7, 3, ftc.n_ravel, Hc.shape[0])
    st = o.init_chain(hp, ftc, [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    L0 = st.L_cur[2].copy()
    phi0 = o.phi_woodbury(L0, Hc, HtH, y, delta)
    steps = []
    for it in range(6):
        d = o.one_step(st, ftc, Hc, HtH, y, delta, ns.draw_iteration())
        d["halves"] = np.stack(o.current_halves(st, ftc))
        d["u_new_last"] = st.u_new_sym[2].copy()
        steps.append(d)
    np.savez_compressed(os.path.join(OUT, "oracle_chain_n4.npz"), y=y, delta=delta, phi0=phi0,
                        L0_checksum=np.array([L0.real.sum(), L0.imag.sum(), np.abs(L0).sum()]),
                        accepted=np.array([s["accepted"] for s in steps]), phi_cur=np.array([s["phi_cur"] for s in steps]),
                        phi_new=np.array([s["phi_new"] for s in steps]), log_u=np.array([s["log_u"] for s in steps]),
                        halves=np.stack([s["halves"] for s in steps]), u_new_last=np.stack([s["u_new_last"] for s in steps]))
    print("goldens written to", OUT)


def config2_golden(steps=2):
    """4. oracle_config2_n32.npz : BASELINE configs[1] (n=32, n_ext=64, 90 angles, 3 layers) -- two whole oracle iterations
    (25-40 s of CPU each), so the GPU test can compare the full-size CUDA step with the oracle itself and not only with a
    cuSOLVER rendition.  Inputs are regenerated on the GPU box by the same deterministic oracle calls (seconds); the golden
    stores the expensive outputs: Phi pair, accept flag, log-ratio margin, the three half-samples, the Gibbs draw."""
    hp = o.Hyper(n_layers=3, n=32, n_ext=64, n_theta=90, beta=0.25)
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    ns = o.NoiseStream(32, 3, ft.n_ravel, H.shape[0])
    st = o.init_chain(hp, ft, [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    out = []
    for it in range(steps):
        d = o.one_step(st, ft, H, HtH, y, delta, ns.draw_iteration())
        d["halves"] = np.stack(o.current_halves(st, ft))
        d["u_new"] = np.stack([u.copy() for u in st.u_new_sym])
        print("config2 step", it, {k: v for k, v in d.items() if np.isscalar(v)}, flush=True)
        out.append(d)
    np.savez_compressed(os.path.join(OUT, "oracle_config2_n32.npz"), seed=32, beta=0.25, y_checksum=np.array([y.sum(), np.abs(y).sum()]),
                        delta=delta, accepted=np.array([s["accepted"] for s in out]), phi_cur=np.array([s["phi_cur"] for s in out]),
                        phi_new=np.array([s["phi_new"] for s in out]), log_u=np.array([s["log_u"] for s in out]),
                        log_ratio=np.array([s["log_ratio"] for s in out]), halves=np.stack([s["halves"] for s in out]),
                        u_new=np.stack([s["u_new"] for s in out]))


if __name__ == "__main__":
    if "--config2" in sys.argv:
        config2_golden()
    else:
        main()
