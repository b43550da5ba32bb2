"""CPU tests: the numpy oracle against vectors computed by the REFERENCE'S OWN Python classes.

``tests/golden/ref_step_*.npz`` were written by ``oracle/make_ref_goldens.py``, which imports
``/root/reference/mcmc/image_cupy.py`` unmodified (numpy standing in for CuPy, numba's CUDA simulator running the two
numba.cuda kernels) and runs ``Simulation.__init__``, ``Lmatrix_2D.construct_from``, ``pCN._log_ratio``, ``util.lstsq``,
``pCN.one_step``, ``pCN.adapt_step`` and ``one_step_for_sqrtBetas`` with a logged noise stream.  Two modes:

* ``alias``  (complex64 -> complex128): the reference's semantics in FP64.  Tolerance 1e-12 (measured: <= 3e-15, L and
  every table bit-identical).
* ``native`` (the reference's own mixed precision): each step restarted from the reference's previous state.  Vectors that
  only see one rounding (noise, layer-0 sample) agree to 2e-6 (measured <= 1e-7); Phi and the solved samples carry the
  reference's complex64 H / HtH / U through the conditioning of r = L^H L + HtH: measured 3e-7 at n <= 4 and 2e-4 at n = 8.
  That gap is the reference's own 32-bit rounding, which is why parity is stated against ``alias``.
"""
import json
import os

import numpy as np
import pytest

from oracle import mspde_oracle as o
from ref_golden_util import GOLDEN, RefGolden, cases, checksum, checksum_close

ALIAS_TOL = 1e-12
NATIVE_TOL = 2e-6


def rel(a, b):
    return np.abs(np.asarray(a) - np.asarray(b)).max() / max(np.abs(np.asarray(b)).max(), 1e-300)


def test_golden_files_present():
    assert {"n2", "n3", "n4", "n4k2", "n4k4", "n8", "n16"} <= set(cases("alias"))
    assert {"n2", "n3", "n4", "n4k2", "n8"} <= set(cases("native"))


@pytest.mark.parametrize("case", cases("alias"))
def test_tables_and_fixed_inputs_alias(case):
    """a1 / a14: index tables bit-exact; D_diag, prior tables, H, HtH, delta as the reference built them."""
    R = RefGolden(case, "alias")
    g, ft = R.g, R.ft
    assert np.array_equal(ft.ix, g["ix"]) and np.array_equal(ft.iy, g["iy"]) and np.array_equal(ft.Kmatrix, g["Kmatrix"])
    assert np.array_equal(ft.D_diag, g["D_diag"])
    hp = o.Hyper(n_layers=R.K, n=R.n, n_ext=R.n_ext, n_theta=R.n_theta)
    pt = o.prior_tables(hp, ft)
    assert np.array_equal(np.array(pt["sqrt_betas"]), g["sqrt_betas"])
    assert np.array_equal(pt["stdev_sym"], g["stdev_sym"])
    H, HtH = o.normalised_inputs(o.tomography_H(ft, R.n_theta), float(g["meas_std"]))
    assert rel(H[g["H_rows_idx"]], g["H_rows"]) <= 1e-14
    assert checksum_close(H, g["H_checksum"], 1e-14) and checksum_close(HtH, g["HtH_checksum"], 1e-13)
    if "H" in g.files:
        assert rel(H, g["H"]) <= 1e-14 and rel(HtH, g["HtH"]) <= 1e-14
    assert abs(o.chol_stabilizer(R.inputs()[1], float(g["chol_epsilon"])) - R.delta) <= 1e-14 * R.delta


@pytest.mark.parametrize("case", [c for c in cases("alias") if c != "n16"])
def test_construction_alias(case):
    """Simulation.__init__ / Layer.__init__ (image_cupy.py:717-766, 902-946): initial noise, samples, L, logdet."""
    R = RefGolden(case, "alias")
    g, ft, K = R.g, R.ft, R.K
    hp = o.Hyper(n_layers=K, n=R.n, n_ext=R.n_ext, n_theta=R.n_theta, beta=float(g["beta0"]))
    noise, solve = R.init_draws()
    st = o.init_chain(hp, ft, noise, solve)
    for k in range(K):
        assert rel(st.noise_cur[k], g[f"init_noise_cur_{k}"]) <= ALIAS_TOL
        assert rel(st.u_new_sym[k], g[f"init_u_new_{k}"]) <= ALIAS_TOL
        assert rel(st.u_cur_sym[k], g[f"init_u_cur_{k}"]) <= ALIAS_TOL
        if k >= 1:
            assert checksum_close(st.L_cur[k], g[f"init_L_checksum_{k}"], 1e-14)
            if f"init_L_{k}" in g.files:
                assert np.array_equal(st.L_cur[k], g[f"init_L_{k}"])          # a3 + a4: bit-identical
            assert abs(o.logdet_L(st.L_cur[k]) - float(g[f"init_logdet_{k}"])) <= 1e-11 * abs(float(g[f"init_logdet_{k}"]))


@pytest.mark.parametrize("mode,tol", [("alias", 1e-14), ("native", NATIVE_TOL)])
def test_field_tables(mode, tol):
    """a2: FourierAnalysis_2D.irfft2 -> exp -> rfft2 -> symmetrize_2D (image_cupy.py:108-149)."""
    for case in cases(mode):
        R = RefGolden(case, mode)
        g = R.g
        img = o.irfft2(o.from_u_2D_ravel_to_uHalf_2D(g["a2_u_sym"], R.n), R.ft)
        tp, tm = o.field_coeffs(g["a2_u_sym"], R.ft)
        assert rel(img, g["a2_image"]) <= tol and rel(tp, g["a2_table_plus"]) <= tol and rel(tm, g["a2_table_minus"]) <= tol


@pytest.mark.parametrize("case", [c for c in cases("alias") if c != "n16"])
def test_phi_and_lstsq_alias(case):
    """a9 (default and naive branch) and a11 on the reference's initial last-layer L."""
    R = RefGolden(case, "alias")
    g = R.g
    H, HtH = R.inputs()
    st = R.state("init")
    L = st.L_cur[R.K - 1]
    phi = o.phi_woodbury(L, H, HtH, R.y, R.delta)
    assert abs(phi - float(g["phi_init"])) <= ALIAS_TOL * abs(phi)
    if np.isfinite(float(g["phi_init_naive"])):
        # the unstabilised branch (645-657) is ill-conditioned in itself; 1e-8 is what two LAPACK runs agree to
        assert abs(o.phi_naive(L, H, R.y, 0.0) - float(g["phi_init_naive"])) <= 1e-8 * abs(phi)
    x = o.lstsq_qr(np.vstack((H, L)), g["lstsq_rhs"])
    assert rel(x, g["lstsq_x"]) <= 1e-11


def _check_step(R, st, d, s, tol, prefix="s", phi_idx=(0, 1)):
    g = R.g
    ph = g[f"{prefix}{s}_phi"]
    assert abs(d["phi_cur"] - ph[phi_idx[0]]) <= tol * abs(ph[phi_idx[0]])
    assert abs(d["phi_new"] - ph[phi_idx[1]]) <= tol * abs(ph[phi_idx[1]])


@pytest.mark.parametrize("case", cases("alias"))
def test_one_step_chain_alias(case):
    """pCN.one_step (553-617) free-running over all recorded steps, then adapt_step (528-540), then the steps with the
    sqrt-beta hyper-move (666-710).  Accept decisions identical, everything else within 1e-12."""
    R = RefGolden(case, "alias")
    g, ft, K = R.g, R.ft, R.K
    H, HtH = R.inputs()
    hp = o.Hyper(n_layers=K, n=R.n, n_ext=R.n_ext, n_theta=R.n_theta)
    st = R.state("init", R.beta_for_step(0))
    for k in range(1, K):
        assert checksum_close(st.L_cur[k], g[f"init_L_checksum_{k}"], 1e-14)
    min_margin = np.inf
    for s in range(R.n_steps):
        assert abs(st.beta - R.beta_for_step(s)) <= 1e-15
        d = o.one_step(st, ft, H, HtH, R.y, R.delta, R.step_noise(s))
        assert d["accepted"] == bool(g["accepted"][s])
        min_margin = min(min_margin, abs(d["log_ratio"] - d["log_u"]))
        _check_step(R, st, d, s, ALIAS_TOL)
        for k in range(K):
            for name, mine in (("noise_new", st.noise_new), ("noise_cur", st.noise_cur), ("u_new", st.u_new_sym), ("u_cur", st.u_cur_sym)):
                assert rel(mine[k], g[f"s{s}_{name}_{k}"]) <= 1e-11, (s, k, name)
            if k >= 1:
                assert checksum_close(st.L_latest[k], g[f"s{s}_L_latest_checksum_{k}"], 1e-10)   # exp(u) amplifies 1e-15 in u by |u|
                assert checksum_close(st.L_cur[k], g[f"s{s}_L_cur_checksum_{k}"], 1e-10)
        if "adapt_after_step" in g.files and s == int(g["adapt_after_step"]):
            o.adapt_step(st, hp)
            assert abs(st.beta - float(g["adapt_beta"])) <= 1e-15 and abs(st.betaZ - float(g["adapt_betaZ"])) <= 1e-15
    assert min_margin > 1e-6 * abs(float(g["phi_init"])) * 1e-4          # decisions are not near-ties at 1e-10 relative
    lm = [float(x) for x in g["lmat_sqrt_betas"]]
    for s in range(R.n_sb_steps):
        nz = R.step_noise(s, "b")
        d = o.one_step(st, ft, H, HtH, R.y, R.delta, nz)
        assert d["accepted"] == bool(g["b_accepted"][s])
        _check_step(R, st, d, s, ALIAS_TOL, "b")
        d2 = o.one_step_for_sqrt_betas(st, ft, H, HtH, R.y, R.delta, nz["move"], lm, float(g["pcn_step_sqrtBetas"]))
        assert d2["accepted"] == bool(g["b_accepted_sqrtbeta"][s])
        _check_step(R, st, d2, s, 1e-11, "b", (2, 3))
        assert np.abs(np.array(st.sqrt_betas) / g[f"b{s}_sqrt_betas"] - 1).max() <= 1e-13
        assert rel(st.stdev_sym, g[f"b{s}_stdev_sym"]) <= 1e-13
        for k in range(K):
            assert rel(st.u_new_sym[k], g[f"b{s}_u_new_{k}"]) <= 1e-10, (s, k)
    # a12: the recorded history rows (complex64 host arrays) are the current half-samples after each step
    hist = g[f"history_{K - 1}"]
    assert hist.dtype == np.complex64 and hist.shape[1] == ft.n_ravel


@pytest.mark.parametrize("case", cases("native"))
def test_one_step_native(case):
    """The same steps against the reference's own MIXED precision, each restarted from the reference's state."""
    R = RefGolden(case, "native")
    g, ft, K = R.g, R.ft, R.K
    H, HtH = R.inputs()
    prev = "init"
    for s in range(R.n_steps):
        st = R.state(prev, R.beta_for_step(s))
        d = o.one_step(st, ft, H, HtH, R.y, R.delta, R.step_noise(s))
        ph = g[f"s{s}_phi"]
        ref_ratio = ph[0] - ph[1]
        phi_tol = NATIVE_TOL if R.n <= 4 else 1e-3
        if abs(ref_ratio - d["log_u"]) > 10 * phi_tol * abs(ph[0]):
            assert d["accepted"] == bool(g["accepted"][s])
        _check_step(R, st, d, s, phi_tol)
        for k in range(K):
            assert rel(st.noise_new[k], g[f"s{s}_noise_new_{k}"]) <= NATIVE_TOL
            assert rel(st.u_new_sym[k], g[f"s{s}_u_new_{k}"]) <= (10 * NATIVE_TOL if R.n <= 4 else 1e-3)
        prev = f"s{s}"


def test_recorded_history_is_current_half_alias():
    """a12: Layer.record_sample stores current_sample = current_sample_sym[n_ravel-1:] after every step (792-795)."""
    R = RefGolden("n4", "alias")
    g, ft = R.g, R.ft
    for k in range(R.K):
        hist = g[f"history_{k}"]
        row = 1 if k == R.K - 1 else 0          # Simulation.__init__ never records; the driver's FBP init records row 0 of
        # the last layer only when called before run() -- here the goldens call it afterwards, so rows start at 0
        for s in range(R.n_steps):
            want = g[f"s{s}_u_cur_{k}"][ft.n_ravel - 1:].astype(np.complex64)
            assert np.array_equal(hist[s], want), (k, s)


def test_save_layout_golden_is_complete():
    """f3: the key layout util._save_object writes (util_cupy.py:553-583), captured from the reference itself."""
    lay = json.load(open(os.path.join(GOLDEN, "ref_save_layout.json")))
    for key in ("n_layers", "n_samples", "Layers 0/samples_history", "Layers 2/samples_history", "fourier/basis_number",
                "fourier/extended_basis_number", "measurement/sinogram", "measurement/theta", "measurement/stdev", "t_start", "t_end",
                "burn_percentage", "pcn/H_conj_T", "pcn/yBar", "fourier/_Umask", "fourier/D_diag", "fourier/Kmatrix"):
        assert key in lay, key
    for key in ("pcn/H", "pcn/H_t_H", "pcn/I", "pcn/In", "pcn/y", "fourier/ix", "fourier/iy", "fourier/Imatrix"):
        assert key not in lay, key              # excluded_matrix (554)
