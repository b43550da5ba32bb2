"""CPU tests: the numpy oracle against the committed golden vectors (reference kernel bodies run verbatim under the
numba CUDA simulator, reference CPU helpers) and against its own algebraic cross-checks."""
import os

import numpy as np
import pytest

from oracle import mspde_oracle as o


@pytest.mark.parametrize("n", [2, 3, 4])
def test_bttb_index_map_matches_reference_kernel_bit_exact(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"ref_construct_U_n{n}.npz"))
    il = 2 * n - 1
    r, c, mask = o.bttb_index_map(n)
    mine = np.where(mask, g["labels"][np.clip(r, 0, il - 1), np.clip(c, 0, il - 1)], 0)
    assert np.array_equal(mine, g["U"])
    # and through construct_U itself
    U = o.construct_U(g["labels"].astype(np.complex128))
    assert np.array_equal(U.real.astype(np.int32), g["U"]) and not U.imag.any()


def test_tomography_closed_form_matches_reference_kernel(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_H_tomography.npz"))
    ft = o.FourierTables(int(g["n"]), int(g["n_ext"]))
    n_theta, n_r = int(g["n_theta"]), ft.m
    Hk = g["H_kernel"] * (n_r / (2 * (n_r // 2)))
    for b in range(n_theta // 4):
        Hk[b * n_r:(b + 1) * n_r] = np.roll(Hk[b * n_r:(b + 1) * n_r], 1, axis=0)
    H = o.tomography_H(ft, n_theta)
    assert np.abs(H - Hk).max() < 1e-14
    assert np.isfinite(H).all()
    assert np.abs(H[:, ::-1].conj() - H).max() < 1e-14       # H[:, N-1-j] = conj(H[:, j])  (Appendix A.8)


def test_helpers_match_reference_cpu_helpers(golden_dir):
    g = np.load(os.path.join(golden_dir, "ref_util_helpers.npz"))
    assert np.array_equal(o.symmetrize(g["half"]), g["sym"])
    assert np.array_equal(o.symmetrize_2D(g["uh"]), g["sym2D"])
    assert np.array_equal(o.extend2D(g["sym2D"], 6), g["ext2D"])


def test_index_tables():
    ft = o.FourierTables(4, 8)
    il = 7
    temp = np.arange(-3, 4)
    assert np.array_equal(ft.ix.ravel("F"), np.repeat(temp, il))          # A.1: ix.ravel('F')[j] = temp[j // il]
    assert np.array_equal(ft.iy.ravel("F"), np.tile(temp, il))
    assert ft.D_diag[ft.n_ravel - 1] == 0.0 and ft.D_diag.max() == 0.0
    c32 = np.float32(np.float32(2) * np.float32(np.pi)) ** 2
    assert ft.D_diag[0] == -np.float64(np.float32(c32)) * 18


def test_constant_field_gives_scaled_identity():
    ft = o.FourierTables(4, 8)
    u = np.zeros(ft.N, dtype=np.complex128)                                 # u(x) = 0 -> exp(u) = 1 -> U = m I
    tp, tm = o.field_coeffs(u, ft)
    U = o.construct_U(tp)
    assert np.abs(U - ft.m * np.eye(ft.N)).max() < 1e-12
    L = o.assemble_L(u, 2.0, ft)
    assert np.abs(L - np.diag(ft.m * (ft.D_diag - 1.0) / 2.0)).max() < 1e-10


def test_bttb_is_central_convolution():
    import scipy.signal as ss
    n = 3
    ft = o.FourierTables(n, 4)
    rng = np.random.default_rng(1)
    T = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    V = rng.standard_normal((5, 5)) + 1j * rng.standard_normal((5, 5))
    U = o.construct_U(T)
    lhs = (U @ V.ravel("F")).reshape(5, 5, order="F")
    rhs = ss.convolve2d(T, V, mode="full")[n - 1:n - 1 + 5, n - 1:n - 1 + 5]
    assert np.abs(lhs - rhs).max() < 1e-12


@pytest.fixture(scope="module")
def toy():
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8)
    pb = o.synthetic_problem(hp)
    ns = o.NoiseStream(7, 3, pb["ft"].n_ravel, pb["H"].shape[0])
    st = o.init_chain(hp, pb["ft"], [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    return hp, pb, ns, st


def test_phi_three_forms_agree(toy):
    hp, pb, ns, st = toy
    L = st.L_cur[2]
    a = o.phi_woodbury(L, pb["H"], pb["HtH"], pb["y"], pb["delta"])
    b = o.phi_naive(L, pb["H"], pb["y"], pb["delta"])
    c = o.phi_direct(L, pb["H"], pb["y"], pb["delta"])
    assert abs(a - b) <= 1e-11 * abs(a) and abs(a - c) <= 1e-11 * abs(a)
    # delta = 0 on a well-conditioned layer: the reference's verbatim naive branch (645-657).  Without the
    # stabiliser I - Ht^H Ht cancels to ~1e-9, which bounds how well ANY two evaluations of the Woodbury form agree.
    L1 = st.L_cur[1]
    a = o.phi_woodbury(L1, pb["H"], pb["HtH"], pb["y"], 0.0)
    b = o.phi_naive(L1, pb["H"], pb["y"], 0.0)
    assert abs(a - b) <= 1e-7 * abs(a)


def test_hermitian_symmetry_preserved_by_solve(toy):
    hp, pb, ns, st = toy
    ft = pb["ft"]
    u = st.u_cur_sym[1]
    assert np.abs(u - u[::-1].conj()).max() <= 1e-9 * np.abs(u).max()
    assert abs(o.logdet_L(st.L_cur[1]) - np.linalg.slogdet(st.L_cur[1])[1]) < 1e-9


def test_oracle_chain_regression(golden_dir, toy):
    g = np.load(os.path.join(golden_dir, "oracle_chain_n4.npz"))
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8)
    pb = o.synthetic_problem(hp)
    assert np.array_equal(pb["y"], g["y"]) and pb["delta"] == float(g["delta"])
    ns = o.NoiseStream(7, 3, pb["ft"].n_ravel, pb["H"].shape[0])
    st = o.init_chain(hp, pb["ft"], [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    for it in range(6):
        d = o.one_step(st, pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"], ns.draw_iteration())
        assert d["accepted"] == bool(g["accepted"][it])
        assert abs(d["phi_cur"] - g["phi_cur"][it]) <= 1e-9 * abs(g["phi_cur"][it])
        assert abs(d["phi_new"] - g["phi_new"][it]) <= 1e-9 * abs(g["phi_new"][it])
        halves = np.stack(o.current_halves(st, pb["ft"]))
        assert np.abs(halves - g["halves"][it]).max() <= 1e-8 * np.abs(g["halves"][it]).max()


def test_noise_conventions():
    w = o.w_half_from_normals(np.array([1.0, 2.0]), np.array([3.0, 4.0]))
    s2 = np.float64(np.float32(1.41421356))
    assert w[0] == 2.0 / s2 and w[1] == (2.0 + 4.0j) / s2                   # DC real, doubled (A.4-2)
    sym = o.symmetrize(w)
    assert sym.shape == (3,) and sym[0] == np.conj(w[1]) and sym[1] == w[0]


def test_adapt_step_rules():
    hp = o.Hyper()
    st = o.ChainState(beta=0.5, betaZ=np.sqrt(0.75), sqrt_betas=[], stdev_sym=None, noise_cur=[], noise_new=[], u_cur_sym=[],
                      u_new_sym=[], L_cur=[], L_latest=[], accepted_total=1, iteration=10)
    o.adapt_step(st, hp)
    assert abs(st.beta - 0.5 * np.exp(2.1 * (0.1 - 0.234))) < 1e-15
    st.beta, st.accepted_total = 0.99, 10                                   # would exceed 1 -> ignored (537-540)
    o.adapt_step(st, hp)
    assert st.beta == 0.99
