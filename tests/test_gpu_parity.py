"""GPU parity tests (-m gpu): the CUDA path, called through the C-ABI, against the numpy oracle on the same seeded
inputs (small sizes), against the committed goldens, and -- at BASELINE config-2 size -- through size-independent
properties and a torch.linalg (cuBLAS/cuSOLVER) rendition used only as a checker.

Tolerances: index/byte work bit-exact; FP64 results 1e-10 relative (BASELINE.json north_star)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import mspde_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-10


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


@pytest.fixture(scope="module")
def lib():
    from mspde import _ffi
    return _ffi.lib()


def _gemm(lib, A, B, C, conj=True, alpha=1.0, beta=0.0, uplo=0, ws=None, nsplit=1):
    from mspde import _ffi
    k, m = A.shape
    n = B.shape[1]
    _ffi.check(lib.mspde_zgemm_tn(_ffi.current_stream(), int(conj), m, n, k, ctypes.c_double(alpha), _ffi.ptr(A), A.stride(0),
                                  _ffi.ptr(B), B.stride(0), ctypes.c_double(beta), _ffi.ptr(C), C.stride(0), uplo, _ffi.ptr(ws),
                                  nsplit), "zgemm_tn")


def crandn(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.complex(torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g),
                         torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g))


@pytest.fixture(params=[0, 1], ids=["4M", "3M"])
def gemm_mode(request):
    """Runs a test under both complex product schemes of the GEMM kernel (include/mspde.h: mspde_set_gemm_mode)."""
    from mspde import _ffi
    default = _ffi.get_gemm_mode()
    _ffi.set_gemm_mode(request.param)
    yield request.param
    _ffi.set_gemm_mode(default)


# ------------------------------------------------------------------------------------------------ building blocks
@pytest.mark.parametrize("m,n,k,conj,uplo,beta", [(64, 64, 16, 1, 0, 0.0), (1, 1, 1, 1, 0, 0.0), (100, 77, 35, 0, 0, 0.0),
                                                 (200, 130, 129, 1, 0, 1.0), (333, 333, 211, 1, 1, 1.0), (49, 1016, 49, 1, 0, -1.0),
                                                 (961, 961, 961, 1, 1, 0.0), (65, 63, 1, 1, 0, 0.0)])
def test_zgemm_tn_matches_cublas(lib, gemm_mode, m, n, k, conj, uplo, beta):
    A = crandn(k, m, seed=1)
    B = A if uplo else crandn(k, n, seed=2)
    C = crandn(m, n, seed=3)
    C0 = C.clone()
    alpha = -1.0 if beta else 1.0
    ref = alpha * ((A.conj().T if conj else A.T) @ B) + beta * C0
    _gemm(lib, A, B, C, conj, alpha, beta, uplo)
    torch.cuda.synchronize()
    if uplo:
        mask = torch.triu(torch.ones(m, n, dtype=torch.bool, device="cuda"))
        assert torch.equal(C[~mask], C0[~mask])                      # strictly-lower part untouched
        assert rel(C[mask].cpu(), ref[mask].cpu()) < 1e-13
    else:
        assert rel(C.cpu(), ref.cpu()) < 1e-13


def test_zgemm_tn_random_shapes_and_views(lib, gemm_mode):
    """40 random (m, n, k) including tiny and non-multiple-of-tile sizes, operands as strided sub-views (the factorizations
    call the kernel on sub-blocks), guard values around C to catch any out-of-bounds store."""
    rng = np.random.default_rng(1234)
    for trial in range(40):
        m, n, k = (int(rng.integers(1, 200)) for _ in range(3))
        if trial % 5 == 0:
            k = int(rng.integers(1, 40))
        conj, beta = bool(rng.integers(0, 2)), float(rng.choice([0.0, 1.0, -0.5]))
        pad = 3
        Abig = crandn(k + 2 * pad, m + 2 * pad + 5, seed=trial)
        Bbig = crandn(k + 2 * pad, n + 2 * pad + 1, seed=trial + 100)
        Cbig = crandn(m + 2 * pad, n + 2 * pad + 7, seed=trial + 200)
        A, B, C = Abig[pad:pad + k, pad:pad + m], Bbig[pad:pad + k, pad:pad + n], Cbig[pad:pad + m, pad:pad + n]
        C0 = Cbig.clone()
        ref = 0.75 * ((A.conj().T if conj else A.T) @ B) + beta * C
        _gemm(lib, A, B, C, conj, 0.75, beta, 0)
        torch.cuda.synchronize()
        assert rel(C.cpu(), ref.cpu()) < 1e-13, (trial, m, n, k)
        guard = torch.ones_like(Cbig, dtype=torch.bool)
        guard[pad:pad + m, pad:pad + n] = False
        assert torch.equal(Cbig[guard], C0[guard]), (trial, m, n, k)       # nothing written outside the m x n block


def test_zgemm_tn_splitk_deterministic(lib, gemm_mode):
    A, B = crandn(3001, 64, seed=4), crandn(3001, 500, seed=5)
    ws = torch.empty(8 * 64 * 500, dtype=torch.complex128, device="cuda")
    C1 = torch.zeros(64, 500, dtype=torch.complex128, device="cuda")
    C2 = torch.zeros_like(C1)
    _gemm(lib, A, B, C1, ws=ws, nsplit=8)
    _gemm(lib, A, B, C2, ws=ws, nsplit=8)
    torch.cuda.synchronize()
    assert torch.equal(C1, C2)
    assert rel(C1.cpu(), (A.conj().T @ B).cpu()) < 1e-13


@pytest.mark.parametrize("n", [1, 5, 64, 65, 200, 777])
def test_potrf_and_trsm(lib, n):
    from mspde import _ffi
    A = crandn(n + 3, n, seed=n)
    Hm = A.conj().T @ A + n * torch.eye(n, dtype=torch.complex128, device="cuda")
    Uref = torch.linalg.cholesky(Hm).conj().T
    W = torch.triu(Hm).contiguous()
    winv = torch.zeros(((n + 63) // 64) * 64 * 64, dtype=torch.complex128, device="cuda")
    info = torch.zeros(4, dtype=torch.int32, device="cuda")
    st = _ffi.current_stream()
    _ffi.check(lib.mspde_zpotrf_upper(st, _ffi.ptr(W), W.stride(0), n, _ffi.ptr(winv), _ffi.ptr(info)))
    torch.cuda.synchronize()
    assert int(info[0]) == 0
    assert rel(torch.triu(W).cpu(), Uref.cpu()) < 1e-12
    B = crandn(n, 37, seed=n + 1)
    X = B.clone()
    _ffi.check(lib.mspde_ztrsm_uh(st, _ffi.ptr(W), W.stride(0), _ffi.ptr(winv), n, _ffi.ptr(X), X.stride(0), 37))
    torch.cuda.synchronize()
    ref = torch.linalg.solve_triangular(Uref.conj().T, B, upper=False)
    assert rel(X.cpu(), ref.cpu()) < 1e-11


@pytest.mark.parametrize("n", [2, 3, 5, 8, 12])
def test_pivoted_lu_solve_and_logdet(n):
    """mspde_solve (blocked LU, partial pivoting, LAPACK pivot rule) vs cuSOLVER getrf/getrs and vs the QR path; random
    complex matrices force row interchanges in every column.  A singular matrix must surface as LinAlgError (info > 0)."""
    from mspde.core import MspdeContext
    ctx = MspdeContext(n, 64, 64, 3)
    N = ctx.N
    ctx.set_inputs(np.zeros((64, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(64), np.ones(N), np.ones(N), 0.0)
    L = crandn(N, N, seed=n)
    w = crandn(N, seed=n + 100)
    Lp = ctx.new_matrix(N, N); Lp.copy_(L)
    x, ld = ctx.solve(Lp, w, want_logdet=True)
    xq, ldq = ctx.solve(Lp, w, want_logdet=True, method="qr")
    ref = torch.linalg.solve(L, w)
    cond = float(torch.linalg.cond(L))
    assert rel(x.cpu(), ref.cpu()) < 1e-13 * cond and rel(xq.cpu(), ref.cpu()) < 1e-13 * cond
    ldr = float(torch.linalg.slogdet(L)[1])
    assert abs(ld - ldr) <= 1e-12 * abs(ldr) and abs(ldq - ldr) <= 1e-12 * abs(ldr)
    assert torch.equal(Lp, L)                                     # inputs are not modified (cp.linalg.solve semantics)
    Ls = Lp.clone()
    Ls[:, N // 2] = 0
    with pytest.raises(np.linalg.LinAlgError):
        ctx.solve(Ls, w, want_logdet=True)
    ctx.close()


def test_potrf_reports_not_positive_definite(lib):
    from mspde import _ffi
    n = 130
    Hm = torch.eye(n, dtype=torch.complex128, device="cuda")
    Hm[70, 70] = -1.0
    winv = torch.zeros(3 * 64 * 64, dtype=torch.complex128, device="cuda")
    info = torch.zeros(4, dtype=torch.int32, device="cuda")
    _ffi.check(lib.mspde_zpotrf_upper(_ffi.current_stream(), _ffi.ptr(Hm), Hm.stride(0), n, _ffi.ptr(winv), _ffi.ptr(info)))
    torch.cuda.synchronize()
    assert int(info[0]) == 71                                        # LAPACK zpotrf info convention (1-based)
    with pytest.raises(np.linalg.LinAlgError):
        _ffi.check(int(info[0]), "zpotrf")


def test_tomography_kernel_vs_reference_kernel_golden(golden_dir, lib):
    """mspde_tomography_H without post-fixes against the reference's own kernel body output (golden), and with them
    against the oracle's closed form."""
    from mspde import _ffi
    from mspde import image as im
    g = np.load(os.path.join(golden_dir, "ref_H_tomography.npz"))
    n, n_ext, n_theta = int(g["n"]), int(g["n_ext"]), int(g["n_theta"])
    ft = o.FourierTables(n, n_ext)
    sino = im.Sinogram("none.png", target_size=ft.m, n_theta=n_theta, stdev=0.2)
    dev = torch.device("cuda", 0)
    H_raw = sino.get_measurement_matrix_device(ft.ix.ravel("F"), ft.iy.ravel("F"), dev, apply_fixes=False).cpu().numpy()
    assert np.abs(H_raw - g["H_kernel"]).max() < 1e-14
    H_fix = sino.get_measurement_matrix_device(ft.ix.ravel("F"), ft.iy.ravel("F"), dev, apply_fixes=True).cpu().numpy()
    assert np.abs(H_fix - o.tomography_H(ft, n_theta)).max() < 1e-14
    assert np.abs(H_fix - sino.get_measurement_matrix(ft.ix.ravel("F"), ft.iy.ravel("F"))).max() < 1e-14
    # config-1 geometry against the numpy closed form (relative 1e-12 on entries of O(1))
    ft2 = o.FourierTables(16, 64)
    s2 = im.Sinogram("none.png", target_size=ft2.m, n_theta=45, stdev=0.2)
    Hd = s2.get_measurement_matrix_device(ft2.ix.ravel("F"), ft2.iy.ravel("F"), dev).cpu().numpy()
    Ho = o.tomography_H(ft2, 45)
    assert np.abs(Hd - Ho).max() < 1e-12 * np.abs(Ho).max()


# ------------------------------------------------------------------------------------------------ assembly
@pytest.fixture(scope="module")
def toy():
    from mspde.core import MspdeContext
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8)
    pb = o.synthetic_problem(hp)
    pt = o.prior_tables(hp, pb["ft"])
    ns = o.NoiseStream(7, 3, pb["ft"].n_ravel, pb["H"].shape[0])
    st = o.init_chain(hp, pb["ft"], [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    ctx = MspdeContext(4, 64, pb["H"].shape[0], 3)
    ctx.set_inputs(pb["H"], pb["HtH"], pb["y"], pb["ft"].D_diag, pt["stdev_sym"], pb["delta"])
    return hp, pb, pt, ns, st, ctx


@pytest.mark.parametrize("n", [2, 3, 4])
def test_assembly_kernel_index_map_bit_exact_vs_reference_kernel(golden_dir, n):
    """Label every table entry, run the CUDA assembly with D = 1, Tm = 0, sqrt_beta = 1: the output must equal the
    reference's own _construct_U output (golden, produced by the reference kernel body) bit for bit."""
    from mspde.core import MspdeContext
    g = np.load(os.path.join(golden_dir, f"ref_construct_U_n{n}.npz"))
    ctx = MspdeContext(n, 8, 16, 2)
    N = ctx.N
    ones = np.ones(N)
    z = np.zeros((4, N), dtype=np.complex128)
    ctx.set_inputs(np.zeros((16, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(16), ones, ones, 0.0)
    tp = ctx.to_dev(g["labels"].astype(np.complex128))
    tm = torch.zeros_like(tp)
    L = ctx.assemble_L(tp, tm, 1.0).cpu().numpy()
    assert np.array_equal(L.real.astype(np.int32), g["U"]) and not L.imag.any()


def test_field_coeffs_and_L_vs_oracle(toy):
    hp, pb, pt, ns, st, ctx = toy
    ft = pb["ft"]
    for k in (1, 2):
        u, sb = st.u_cur_sym[k - 1], st.sqrt_betas[k]
        tp, tm = o.field_coeffs(u, ft)
        tpd, tmd = ctx.field_coeffs(ctx.to_dev(u))
        assert rel(tpd.cpu(), tp) < RTOL and rel(tmd.cpu(), tm) < RTOL
        Lo = o.assemble_L_from_tables(tp, tm, ft.D_diag, sb)
        Ld = ctx.assemble_L(ctx.to_dev(tp), ctx.to_dev(tm), sb).cpu().numpy()
        assert np.array_equal(Ld, Lo)                                # same tables -> bit-identical Galerkin matrix
        Le = ctx.construct_L(ctx.to_dev(u), sb).cpu().numpy()
        assert rel(Le, st.L_cur[k]) < RTOL


@pytest.mark.parametrize("n,n_ext", [(4, 4), (4, 6), (5, 9), (8, 8)])
def test_assembly_other_paddings(n, n_ext):
    """extend2D is a no-op for n_ext = n and pads to m = 2 n_ext - 1 otherwise (util_cupy.py:251-267): other image sizes."""
    from mspde.core import MspdeContext
    ft = o.FourierTables(n, n_ext)
    ctx = MspdeContext(n, n_ext, 16, 2)
    N = ft.N
    ctx.set_inputs(np.zeros((16, N), np.complex128), np.zeros((N, N), np.complex128), np.zeros(16), ft.D_diag, np.ones(N), 0.0)
    rng = np.random.default_rng(n * 10 + n_ext)
    u = o.symmetrize(o.w_half_from_normals(rng.standard_normal(ft.n_ravel), rng.standard_normal(ft.n_ravel))) * 2.0
    tp, tm = o.field_coeffs(u, ft)
    tpd, tmd = ctx.field_coeffs(ctx.to_dev(u))
    assert rel(tpd.cpu(), tp) < RTOL and rel(tmd.cpu(), tm) < RTOL
    assert rel(ctx.construct_L(ctx.to_dev(u), 3.0).cpu(), o.assemble_L(u, 3.0, ft)) < RTOL
    ctx.close()


def test_four_layer_chain_vs_oracle():
    """n_layers = 4: two non-stationary middle layers (two assemblies + two dense solves before the last layer)."""
    from mspde.core import MspdeContext
    hp = o.Hyper(n_layers=4, n=4, n_ext=64, n_theta=8, beta=0.4, sigma_scaling=1e-3)   # default 0.4 overflows exp() four layers deep at this toy size
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    pt = o.prior_tables(hp, ft)
    K, M = 4, H.shape[0]
    ns = o.NoiseStream(5, K, ft.n_ravel, M)
    st = o.init_chain(hp, ft, [ns.half() for _ in range(K)], [ns.half() for _ in range(K)])
    ctx = MspdeContext(4, 64, M, K)
    ctx.set_inputs(H, HtH, y, ft.D_diag, pt["stdev_sym"], delta)
    cb = _load_chain(ctx, st)
    rec = torch.zeros((K, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
    for it in range(4):
        nz = ns.draw_iteration()
        d = o.one_step(st, ft, H, HtH, y, delta, nz)
        out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz["w_half"]], nz["log_u"], ctx.to_dev(nz["w_last"]), ctx.to_dev(nz["e"]),
                           record=rec).cpu().numpy()
        ctx.check_info("step")
        assert int(out[0]) == int(d["accepted"])
        assert abs(out[2] - d["phi_cur"]) <= RTOL * abs(d["phi_cur"]) and abs(out[3] - d["phi_new"]) <= RTOL * abs(d["phi_new"])
        for k in range(K):
            assert rel(cb.u_new[k].cpu(), st.u_new_sym[k]) < 1e-9
    ctx.close()


def test_phi_solve_lstsq_vs_oracle(toy):
    hp, pb, pt, ns, st, ctx = toy
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    N, M = ft.N, H.shape[0]
    L = st.L_cur[2]
    po, parts = o.phi_woodbury(L, H, HtH, y, delta, return_parts=True)
    Ld = ctx.new_matrix(N, N); Ld.copy_(ctx.to_dev(L))
    pg = ctx.phi(Ld, parts=True)
    assert abs(pg[0] - po) <= RTOL * abs(po)
    assert abs(pg[1] - parts["quad"]) <= RTOL * abs(parts["quad"]) and abs(pg[2] - parts["logdet"]) <= 1e-9 * abs(parts["logdet"])
    assert rel(torch.triu(ctx.buffer("r", N, N)).cpu().numpy().conj().T, parts["c"]) < RTOL
    assert rel(ctx.buffer("X", N, M).cpu(), parts["Ht"]) < RTOL
    L1 = st.L_cur[1]
    L1d = ctx.new_matrix(N, N); L1d.copy_(ctx.to_dev(L1))
    x, ld = ctx.solve(L1d, ctx.to_dev(st.noise_cur[1]), want_logdet=True)
    assert rel(x.cpu(), np.linalg.solve(L1, st.noise_cur[1])) < RTOL
    assert abs(ld - o.logdet_L(L1)) <= RTOL * abs(o.logdet_L(L1))
    rng = np.random.default_rng(0)
    rhs = rng.standard_normal(M + N) + 1j * rng.standard_normal(M + N)
    v = ctx.gibbs_lstsq(Ld, ctx.to_dev(rhs))
    assert rel(v.cpu(), o.lstsq_qr(np.vstack((H, L)), rhs)) < RTOL
    a = rng.standard_normal((300, 40)) + 1j * rng.standard_normal((300, 40))
    b = rng.standard_normal(300) + 1j * rng.standard_normal(300)
    xg = ctx.lstsq(ctx.to_dev(a), ctx.to_dev(b))
    assert rel(xg.cpu(), o.lstsq_qr(a, b)) < RTOL


def test_fast_phi_matches_woodbury(toy, full):
    """Determinant-lemma Phi (opt-in) against the oracle at toy size and against the Woodbury GPU path at full size."""
    hp, pb, pt, ns, st, ctx = toy
    N = pb["ft"].N
    for k in (1, 2):
        L = st.L_cur[k]
        po, parts = o.phi_woodbury(L, pb["H"], pb["HtH"], pb["y"], pb["delta"], return_parts=True)
        Ld = ctx.new_matrix(N, N); Ld.copy_(ctx.to_dev(L))
        pf = ctx.phi(Ld, parts=True, fast=True)
        assert abs(pf[0] - po) <= RTOL * abs(po)
        assert abs(pf[1] - parts["quad"]) <= RTOL * abs(parts["quad"])
    c2 = full.pcn.ctx
    L = full.Layers[2].LMat.current_L
    a = c2.phi(L, parts=True)
    b = c2.phi(L, parts=True, fast=True)
    assert abs(a[0] - b[0]) <= RTOL * abs(a[0]), (a, b)


def test_naive_phi_variant(toy, full):
    """f2: the reference's naive Phi branch (image_cupy.py:644-657) -- LU of L^H with M right-hand sides, multi-RHS upper
    solve, Cholesky of Z^H Z + I.  Z = solve(L^H, H^H) is checked at 1e-10 x cond(L); Phi itself only to 1e-7: without the
    stabiliser Q = I + Z^H Z is so ill-conditioned that the oracle's own three formulas agree to ~1e-9 only
    (tests/test_oracle_pins.py).  At full size the multi-right-hand-side solve is checked through its residual."""
    hp, pb, pt, ns, st, ctx = toy
    N, M = pb["ft"].N, pb["H"].shape[0]
    for k in (1, 2):
        L = st.L_cur[k]
        Ld = ctx.new_matrix(N, N); Ld.copy_(ctx.to_dev(L))
        try:
            pn = ctx.phi(Ld, parts=True, naive=True)
        except np.linalg.LinAlgError:
            # cond(L) ~ 5e9 for the last toy layer: I + Z^H Z is numerically indefinite, the reference's own Cholesky at
            # line 656 fails the same way (that is why its default is the stabilised Woodbury form)
            assert k == 2
            pn = None
        Zo = np.linalg.solve(L.conj().T, pb["H"].conj().T)
        assert rel(ctx.buffer("Z", N, M).cpu(), Zo) <= 1e-13 * np.linalg.cond(L)
        if k == 1:
            po = o.phi_naive(L, pb["H"], pb["y"], 0.0)
            assert abs(pn[0] - po) <= 1e-7 * abs(po), (pn, po)
    c2 = full.pcn.ctx
    L = full.Layers[2].LMat.current_L
    b = c2.phi(L, naive=True)       # differs from the Woodbury value by the effect of delta on the near-null space of L
    assert np.isfinite(b)
    Z = c2.buffer("Z", c2.N, c2.M)
    resid = L.contiguous().conj().T @ Z - c2.Hct
    assert float(resid.abs().max()) <= 1e-9 * float(c2.Hct.abs().max()) * 1e3


def _load_chain(ctx, st):
    from mspde.core import ChainBuffers
    K = len(st.noise_cur)
    cb = ChainBuffers(ctx, st.sqrt_betas, st.beta)
    for k in range(K):
        cb.noise_cur[k].copy_(ctx.to_dev(st.noise_cur[k])); cb.noise_new[k].copy_(ctx.to_dev(st.noise_new[k]))
        cb.u_cur[k].copy_(ctx.to_dev(st.u_cur_sym[k])); cb.u_new[k].copy_(ctx.to_dev(st.u_new_sym[k]))
        if k >= 1:
            cb.L_cur[k].copy_(ctx.to_dev(st.L_cur[k])); cb.L_latest[k].copy_(ctx.to_dev(st.L_latest[k]))
    return cb


@pytest.mark.parametrize("n,n_theta,steps,cache", [(4, 8, 6, False), (4, 8, 6, True), (8, 24, 4, False), (3, 12, 5, False),
                                                   (4, 8, 6, "fast"), (8, 24, 4, "serial")])
def test_pcn_chain_vs_oracle(n, n_theta, steps, cache):
    """Whole iterations with injected noise: accept decisions identical, Phi / samples within 1e-10."""
    from mspde.core import MspdeContext, STEP_CACHE_PHI_CUR, STEP_FAST_PHI, STEP_SERIAL
    flag = {False: 0, True: STEP_CACHE_PHI_CUR, "fast": STEP_FAST_PHI, "serial": STEP_SERIAL}[cache]
    hp = o.Hyper(n_layers=3, n=n, n_ext=64, n_theta=n_theta)
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    pt = o.prior_tables(hp, ft)
    K, M = 3, H.shape[0]
    ns = o.NoiseStream(7, K, ft.n_ravel, M)
    st = o.init_chain(hp, ft, [ns.half() for _ in range(K)], [ns.half() for _ in range(K)])
    ctx = MspdeContext(n, 64, M, K)
    ctx.set_inputs(H, HtH, y, ft.D_diag, pt["stdev_sym"], delta)
    cb = _load_chain(ctx, st)
    rec = torch.zeros((K, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
    min_margin = np.inf
    for it in range(steps):
        nz = ns.draw_iteration()
        d = o.one_step(st, ft, H, HtH, y, delta, nz)
        cb.set_step(st.beta)
        out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz["w_half"]], nz["log_u"], ctx.to_dev(nz["w_last"]), ctx.to_dev(nz["e"]),
                           flags=flag, record=rec)
        ctx.check_info("step")
        g = out.cpu().numpy()
        min_margin = min(min_margin, abs(d["log_ratio"] - d["log_u"]))
        assert int(g[0]) == int(d["accepted"]), (it, g, d)
        assert abs(g[2] - d["phi_cur"]) <= RTOL * abs(d["phi_cur"])
        assert abs(g[3] - d["phi_new"]) <= RTOL * abs(d["phi_new"])
        halves = o.current_halves(st, ft)
        for k in range(K):
            assert rel(rec[k].cpu(), halves[k]) < RTOL
            assert rel(cb.u_new[k].cpu(), st.u_new_sym[k]) < RTOL
            assert rel(cb.noise_cur[k].cpu(), st.noise_cur[k]) < 1e-14
        if (it + 1) % 2 == 0:
            o.adapt_step(st, hp)
    assert min_margin > 1e-6      # no accept decision was a near-tie, so the identical decisions are meaningful


def test_sqrt_beta_move_vs_oracle():
    """Optional hyper-parameter move (image_cupy.py:666-710) through the host mirror vs the oracle restatement."""
    from mspde import image as im
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8, beta=0.5)
    sim = im.Simulation(n_layers=3, n_samples=4, n=4, n_extended=64, beta=0.5, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                        sigma_scaling=0.4, meas_std=0.2, evaluation_interval=2, printProgress=False, seed=3, burn_percentage=25.0,
                        enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png", n_theta=8)
    sim.pcn.set_chol_epsilon(1e-6)
    pcn, ctx, Layers = sim.pcn, sim.pcn.ctx, sim.Layers
    pcn._ensure_inputs(Layers[0].stdev_sym_host)
    ft = o.FourierTables(4, 64)
    H, HtH, y = ctx.H.cpu().numpy(), ctx.HtH.cpu().numpy(), ctx.y.cpu().numpy()
    K = 3
    st = o.ChainState(beta=0.5, betaZ=float(np.sqrt(0.75)), sqrt_betas=[float(l.sqrt_beta) for l in Layers],
                      stdev_sym=Layers[0].stdev_sym_host.copy(), noise_cur=[l._noise_cur.cpu().numpy() for l in Layers],
                      noise_new=[l._noise_new.cpu().numpy() for l in Layers], u_cur_sym=[l._u_cur.cpu().numpy() for l in Layers],
                      u_new_sym=[l._u_new.cpu().numpy() for l in Layers],
                      L_cur=[None] + [l.LMat.current_L.cpu().numpy() for l in Layers[1:]],
                      L_latest=[None] + [l.LMat.latest_computed_L.cpu().numpy() for l in Layers[1:]])
    lmat_sb = [float(l.LMat.sqrt_beta) for l in Layers]
    rng = np.random.default_rng(5)
    for rep, log_u in enumerate([-1e9, 1e9, -1e9]):               # forced accept, forced reject, accept again
        nz = dict(z=rng.standard_normal(K), e=rng.standard_normal(ctx.M), log_u=log_u)
        d = o.one_step_for_sqrt_betas(st, ft, H, HtH, y, ctx.delta, nz, lmat_sb)
        acc = pcn.one_step_for_sqrtBetas(Layers, noise=nz)
        g = pcn.last_sqrt_beta_step
        assert acc == int(d["accepted"]) == (1 if log_u < 0 else 0)
        assert np.allclose(g["prop"], d["prop"], rtol=1e-14)
        assert abs(g["phi_cur"] - d["phi_cur"]) <= RTOL * abs(d["phi_cur"])
        assert abs(g["phi_new"] - d["phi_new"]) <= RTOL * abs(d["phi_new"])
        for k in range(K):
            assert rel(Layers[k]._u_new.cpu(), st.u_new_sym[k]) < RTOL
            assert abs(Layers[k].sqrt_beta - st.sqrt_betas[k]) <= 1e-14 * st.sqrt_betas[k]
        assert rel(Layers[2].LMat.current_L.cpu(), st.L_cur[2]) < RTOL
        assert rel(Layers[0].stdev_sym_host, st.stdev_sym) < 1e-15
    # and the regular step still runs afterwards with the updated hyper-parameters
    pcn.use_beta_adaptation = True
    a, b = pcn.one_step(Layers)
    assert a in (0, 1) and b in (0, 1)


def test_on_device_posterior_statistics(toy):
    """f4: Welford mean/variance of u(x) and exp(u(x)) accumulated on the device vs numpy on the oracle's fields."""
    hp, pb, pt, ns, st, ctx = toy
    ft = pb["ft"]
    rng = np.random.default_rng(9)
    stats = ctx.new_stats()
    fields = []
    for k in range(7):
        u = o.symmetrize(o.w_half_from_normals(rng.standard_normal(ft.n_ravel), rng.standard_normal(ft.n_ravel))) * (1 + k)
        fields.append(o.irfft2(o.from_u_2D_ravel_to_uHalf_2D(u, ft.n), ft))
        stats["count"] += 1
        ctx.posterior_stats_update(ctx.to_dev(u), stats["count"], stats)
    F = np.stack(fields)
    assert rel(stats["mean_u"].cpu(), F.mean(0)) < 1e-12
    assert rel((stats["m2_u"] / 7).cpu(), F.var(0)) < 1e-11
    assert rel(stats["mean_ell"].cpu(), np.exp(F).mean(0)) < 1e-12
    assert rel((stats["m2_ell"] / 6).cpu(), np.exp(F).var(0, ddof=1)) < 1e-10


def test_config1_two_layer_chain_vs_oracle():
    """BASELINE configs[0] geometry (2 layers, n=16, 45 angles: N=961, M=5715) -- the reference's CPU-runnable case:
    two whole iterations against the oracle with injected noise (accept decisions identical, 1e-10 on Phi and samples)."""
    from mspde.core import MspdeContext
    hp = o.Hyper(n_layers=2, n=16, n_ext=64, n_theta=45, beta=0.2)
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    pt = o.prior_tables(hp, ft)
    K, M = 2, H.shape[0]
    ns = o.NoiseStream(21, K, ft.n_ravel, M)
    st = o.init_chain(hp, ft, [ns.half() for _ in range(K)], [ns.half() for _ in range(K)])
    ctx = MspdeContext(16, 64, M, K)
    ctx.set_inputs(H, HtH, y, ft.D_diag, pt["stdev_sym"], delta)
    cb = _load_chain(ctx, st)
    rec = torch.zeros((K, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
    for it in range(2):
        nz = ns.draw_iteration()
        d = o.one_step(st, ft, H, HtH, y, delta, nz)
        out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz["w_half"]], nz["log_u"], ctx.to_dev(nz["w_last"]), ctx.to_dev(nz["e"]),
                           record=rec)
        ctx.check_info("step")
        g = out.cpu().numpy()
        assert int(g[0]) == int(d["accepted"]) and abs(d["log_ratio"] - d["log_u"]) > 1e-6
        assert abs(g[2] - d["phi_cur"]) <= RTOL * abs(d["phi_cur"]) and abs(g[3] - d["phi_new"]) <= RTOL * abs(d["phi_new"])
        for k in range(K):
            assert rel(cb.u_new[k].cpu(), st.u_new_sym[k]) < RTOL
            assert rel(rec[k].cpu(), o.current_halves(st, ft)[k]) < RTOL
    ctx.close()


def test_chain_against_committed_golden(golden_dir, toy):
    """Same toy chain as tests/golden/oracle_chain_n4.npz (generated by oracle/make_golden.py in the build container)."""
    hp, pb, pt, ns_unused, st_unused, ctx = toy
    g = np.load(os.path.join(golden_dir, "oracle_chain_n4.npz"))
    ft = pb["ft"]
    ns = o.NoiseStream(7, 3, ft.n_ravel, pb["H"].shape[0])
    st = o.init_chain(hp, ft, [ns.half() for _ in range(3)], [ns.half() for _ in range(3)])
    cb = _load_chain(ctx, st)
    rec = torch.zeros((3, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
    for it in range(6):
        nz = ns.draw_iteration()
        out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz["w_half"]], nz["log_u"], ctx.to_dev(nz["w_last"]), ctx.to_dev(nz["e"]),
                           record=rec).cpu().numpy()
        assert int(out[0]) == int(g["accepted"][it])
        assert abs(out[2] - g["phi_cur"][it]) <= RTOL * abs(g["phi_cur"][it])
        assert abs(out[3] - g["phi_new"][it]) <= RTOL * abs(g["phi_new"][it])
        assert rel(rec.cpu(), g["halves"][it]) < RTOL
        assert rel(cb.u_new[2].cpu(), g["u_new_last"][it]) < RTOL


# ------------------------------------------------------------------------------------------------ host mirror, full size
def test_simulation_surface_small():
    """The reference-facing classes run end to end and honour the reference's error behaviour."""
    from mspde import image as im
    sim = im.Simulation(n_layers=3, n_samples=6, n=4, n_extended=64, beta=0.5, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                        sigma_scaling=0.4, meas_std=0.2, evaluation_interval=2, printProgress=False, seed=3, burn_percentage=25.0,
                        enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png", n_theta=8)
    sim.pcn.set_chol_epsilon(1e-6)
    sim.run()
    assert sim.Layers[0].i_record == 6 and np.isfinite(sim.Layers[2].samples_history).all()
    assert 0.0 <= sim.acceptancePercentage <= 1.0
    L = sim.Layers[2].LMat.current_L
    assert abs(sim.pcn._log_ratio(L) - sim.pcn.last_step["phi_cur"]) >= 0.0     # callable on its own
    # f3/f4: statistics, result file, chain continuation, FBP initialisation
    from mspde import io
    sim.pcn.accumulate_stats = True
    sim.pcn.one_step(sim.Layers)
    fs = sim.Layers[1].field_statistics()
    assert fs["count"] == 1 and torch.isfinite(fs["u_mean"]).all()
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        fn = sim.save(os.path.join(d, "result_0.hdf5"))
        last = io.load_last_samples(fn)
        # the file holds complex64 histories, as the reference's does (image_cupy.py:729, 939)
        assert len(last) == 3 and np.array_equal(last[2], sim.Layers[2].samples_history[sim.Layers[2].i_record - 1].astype(np.complex64))
        sim2 = im.Simulation(n_layers=3, n_samples=3, n=4, n_extended=64, beta=0.5, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                             sigma_scaling=0.4, meas_std=0.2, evaluation_interval=2, printProgress=False, seed=4,
                             burn_percentage=25.0, enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png",
                             n_theta=8)
        io.initialize_from_file(sim2, fn)
        assert rel(sim2.Layers[2].current_sample.cpu(), last[2].astype(np.complex128)) < 1e-15
        img = io.initialize_using_FBP(sim2)
        assert img.shape == (127, 127) and np.isfinite(img).all()
        sim2.pcn.set_chol_epsilon(1e-6)
        sim2.run()
    # a non-positive-definite system must surface as numpy.linalg.LinAlgError (Simulation.run's only except clause)
    sim.pcn.ctx.update_scalars(delta=-1e30)
    with pytest.raises(np.linalg.LinAlgError):
        sim.pcn._log_ratio(L)


@pytest.fixture(scope="module")
def full():
    """BASELINE configs[1] geometry: n=32, n_ext=64, 90 angles, 3 layers (N=3969, M=11430)."""
    from mspde import image as im
    sim = im.Simulation(n_layers=3, n_samples=8, n=32, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                        sigma_scaling=0.4, meas_std=0.2, evaluation_interval=10, printProgress=False, seed=1, burn_percentage=25.0,
                        enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png", n_theta=90)
    sim.pcn.set_chol_epsilon(1e-6)
    sim.pcn._ensure_inputs(sim.Layers[0].stdev_sym_host)
    return sim


def test_full_size_constant_field_is_scaled_identity(full):
    ctx = full.pcn.ctx
    u = torch.zeros(ctx.N, dtype=torch.complex128, device=ctx.device)
    L = ctx.construct_L(u, 2.0)
    d = torch.diagonal(L).clone()
    expect = ctx.m * (ctx.D_diag - 1.0) / 2.0
    assert float((d.real - expect).abs().max()) <= 1e-9 * float(expect.abs().max())
    L.diagonal().zero_()
    assert float(L.abs().max()) <= 1e-9 * float(expect.abs().max())


def test_full_size_phi_matches_cusolver_rendition(full):
    """Phi at N=3969, M=11430 against a torch.linalg rendition of image_cupy.py:634-643 (checker only)."""
    pcn, ctx = full.pcn, full.pcn.ctx
    L = full.Layers[2].LMat.current_L
    phi = ctx.phi(L, parts=True)
    Lc = L.contiguous()
    r = Lc.conj().T @ Lc + ctx.HtH + ctx.delta * torch.eye(ctx.N, dtype=torch.complex128, device=ctx.device)
    c = torch.linalg.cholesky(r)
    Ht = torch.linalg.solve_triangular(c, ctx.Hct.contiguous(), upper=False)
    Qi = torch.eye(ctx.M, dtype=torch.complex128, device=ctx.device) - Ht.conj().T @ Ht
    C = torch.linalg.cholesky(Qi)
    quad = 0.5 * float(torch.linalg.norm(ctx.y.to(torch.complex128) @ C) ** 2)
    logdet = float(torch.log(torch.diagonal(C).real).sum())
    ref = quad - logdet
    assert abs(phi[0] - ref) <= RTOL * abs(ref), (phi, ref)
    assert abs(phi[1] - quad) <= RTOL * abs(quad) and abs(phi[2] - logdet) <= 1e-9 * abs(logdet)
    U = torch.triu(ctx.buffer("r", ctx.N, ctx.N))
    assert float((U.conj().T - c).abs().max()) <= 1e-9 * float(c.abs().max())


def test_full_size_solve_and_lstsq_properties(full):
    ctx = full.pcn.ctx
    L = full.Layers[1].LMat.current_L
    w = full.Layers[1].current_noise_sample
    x, ld = ctx.solve(L, w, want_logdet=True)
    Lc = L.contiguous()
    resid = float(torch.linalg.norm(Lc @ x - w) / (torch.linalg.norm(Lc) * torch.linalg.norm(x)))
    assert resid < 1e-14                                             # backward error of the dense solve
    assert abs(ld - float(torch.linalg.slogdet(Lc)[1])) <= RTOL * abs(ld)
    L2 = full.Layers[2].LMat.current_L
    g = torch.Generator(device="cuda").manual_seed(5)
    rhs = torch.complex(torch.randn(ctx.M + ctx.N, dtype=torch.float64, device="cuda", generator=g),
                        torch.randn(ctx.M + ctx.N, dtype=torch.float64, device="cuda", generator=g))
    v = ctx.gibbs_lstsq(L2, rhs)
    A = torch.cat((ctx.H.contiguous(), L2.contiguous()), 0)
    grad = A.conj().T @ (A @ v - rhs)                                # normal-equation residual of the least-squares solution
    assert float(torch.linalg.norm(grad)) <= 1e-11 * float(torch.linalg.norm(A) ** 2 * torch.linalg.norm(v))
    vref = torch.linalg.lstsq(A, rhs.unsqueeze(1)).solution.squeeze(1)
    assert float((v - vref).abs().max()) <= 1e-9 * float(vref.abs().max())


def _torch_assemble(fourier, u_sym, sqrt_beta, D):
    """torch rendition (checker only) of Lmatrix_2D.construct_from: image_cupy.py:108-149, 171-179 + the BTTB gather."""
    from mspde import util
    n = fourier.basis_number
    il = 2 * n - 1
    uh = util.from_u_2D_ravel_to_uHalf_2D(u_sym, n)
    img = fourier.irfft2(uh)
    tp = util.symmetrize_2D(fourier.rfft2(torch.exp(img)))
    tm = util.symmetrize_2D(fourier.rfft2(torch.exp(-img)))
    idx = torch.arange(il * il, device=u_sym.device)
    blk, inn = idx // il, idx % il
    dij = blk[:, None] - blk[None, :]
    dkl = inn[:, None] - inn[None, :]
    mask = (dij.abs() <= n - 1) & (dkl.abs() <= n - 1)
    r = (dkl + n - 1).clamp(0, il - 1)
    c = (dij + n - 1).clamp(0, il - 1)
    Up = torch.where(mask, tp[r, c], torch.zeros((), dtype=tp.dtype, device=tp.device))
    Um = torch.where(mask, tm[r, c], torch.zeros((), dtype=tp.dtype, device=tp.device))
    return (Up * D - Um) / sqrt_beta


def test_full_size_step_vs_torch_rendition(full):
    """One whole pCN iteration at BASELINE configs[1] size against a torch.linalg (cuFFT/cuBLAS/cuSOLVER) rendition of
    pCN.one_step (image_cupy.py:553-617) fed the same injected noise: accept decision, both Phi values and every layer's new
    sample.  torch is the checker here, never the product path."""
    from mspde import util
    pcn, ctx, Layers, f = full.pcn, full.pcn.ctx, full.Layers, full.fourier
    K, N, M, nr = 3, ctx.N, ctx.M, ctx.n_ravel
    rng = np.random.Generator(np.random.PCG64(17))
    wh = [ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))) for _ in range(K - 1)]
    wl = ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr)))
    e = ctx.to_dev(rng.standard_normal(M))
    log_u = -0.7
    beta, betaZ = pcn.beta, pcn.betaZ
    # ---- rendition, from the state before the step
    noise_cur = [l._noise_cur.clone() for l in Layers]
    sd = ctx.stdev_sym
    I_N = torch.eye(N, dtype=torch.complex128, device=ctx.device)

    def phi_t(L):
        r = L.conj().T @ L + ctx.HtH + ctx.delta * I_N
        c = torch.linalg.cholesky(r)
        Ht = torch.linalg.solve_triangular(c, ctx.Hct.contiguous(), upper=False)
        Qi = torch.eye(M, dtype=torch.complex128, device=ctx.device) - Ht.conj().T @ Ht
        C = torch.linalg.cholesky(Qi)
        return 0.5 * float(torch.linalg.norm(ctx.y.to(torch.complex128) @ C) ** 2) - float(torch.log(torch.diagonal(C).real).sum())

    n_new0 = betaZ * noise_cur[0] + beta * util.symmetrize(wh[0])
    u0 = sd * n_new0
    n_new1 = betaZ * noise_cur[1] + beta * util.symmetrize(wh[1])
    L1 = _torch_assemble(f, u0, Layers[1].sqrt_beta, ctx.D_diag)
    u1 = torch.linalg.solve(L1, n_new1)
    del L1
    phi_cur = phi_t(Layers[2].LMat.current_L.contiguous())
    L2 = _torch_assemble(f, u1, Layers[2].sqrt_beta, ctx.D_diag)
    phi_new = phi_t(L2)
    accepted = (phi_cur - phi_new) > log_u
    base = Layers[2]._noise_new.clone() if accepted else noise_cur[2]
    n_new2 = betaZ * base + beta * util.symmetrize(wl)
    rhs = torch.cat(((ctx.y - e).to(torch.complex128), -n_new2))
    A = torch.cat((ctx.H.contiguous(), L2), 0)
    v = torch.linalg.lstsq(A, rhs.unsqueeze(1)).solution.squeeze(1)
    del A, L2
    # ---- product path
    cb = pcn._bind(Layers)
    out = ctx.pcn_step(cb, wh, log_u, wl, e).cpu().numpy()
    ctx.check_info("step")
    assert int(out[0]) == int(accepted) and abs((phi_cur - phi_new) - log_u) > 1e-6
    assert abs(out[2] - phi_cur) <= RTOL * abs(phi_cur) and abs(out[3] - phi_new) <= RTOL * abs(phi_new)
    tol = 1e-9        # the dense solves amplify rounding by cond(L) ~ 1e4-1e5 on both sides
    assert float((Layers[0]._u_new - u0).abs().max()) <= 1e-14 * float(u0.abs().max())
    assert float((Layers[1]._u_new - u1).abs().max()) <= tol * float(u1.abs().max())
    assert float((Layers[2]._noise_new - n_new2).abs().max()) <= 1e-14 * float(n_new2.abs().max())
    assert float((Layers[2]._u_new - v).abs().max()) <= tol * float(v.abs().max())


def test_full_size_steps_are_reproducible(full):
    """Two identical chains fed identical noise produce bit-identical results (deterministic reductions)."""
    from mspde import util
    pcn, ctx, Layers = full.pcn, full.pcn.ctx, full.Layers
    rng = np.random.Generator(np.random.PCG64(3))
    nr, M, K = ctx.n_ravel, ctx.M, 3
    w = [ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))) for _ in range(K - 1)]
    wl = ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr)))
    e = ctx.to_dev(rng.standard_normal(M))
    snap = [(l._noise_cur.clone(), l._u_cur.clone()) for l in Layers]
    res = []
    for rep in range(2):
        for l, (a, b) in zip(Layers, snap):
            l._noise_cur.copy_(a); l._u_cur.copy_(b)
        cb = pcn._bind(Layers)
        out = ctx.pcn_step(cb, w, -0.3, wl, e).clone()
        ctx.check_info("step")
        res.append((out.cpu(), Layers[2]._u_new.clone().cpu(), Layers[1]._u_new.clone().cpu()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert torch.isfinite(res[0][1].abs()).all()


def test_full_size_step_3m_vs_4m(full):
    """The same full-size pCN iteration under the two complex product schemes of the GEMM kernel (6-flop 3M, the default,
    and 8-flop 4M): same accept decision, Phi values within 1e-10 relative, samples within the conditioning of the solves."""
    from mspde import util, _ffi
    pcn, ctx, Layers = full.pcn, full.pcn.ctx, full.Layers
    rng = np.random.Generator(np.random.PCG64(5))
    nr, M, K = ctx.n_ravel, ctx.M, 3
    w = [ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr))) for _ in range(K - 1)]
    wl = ctx.to_dev(util.w_half_from_normals(rng.standard_normal(nr), rng.standard_normal(nr)))
    e = ctx.to_dev(rng.standard_normal(M))
    snap = [(l._noise_cur.clone(), l._u_cur.clone()) for l in Layers]
    default = _ffi.get_gemm_mode()
    res = []
    try:
        for mode in (_ffi.GEMM_4M, _ffi.GEMM_3M):
            _ffi.set_gemm_mode(mode)
            for l, (a, b) in zip(Layers, snap):
                l._noise_cur.copy_(a); l._u_cur.copy_(b)
            cb = pcn._bind(Layers)
            out = ctx.pcn_step(cb, w, -0.3, wl, e).clone()
            ctx.check_info("step")
            res.append((out.cpu().numpy(), Layers[2]._u_new.clone(), Layers[1]._u_new.clone()))
    finally:
        _ffi.set_gemm_mode(default)
        for l, (a, b) in zip(Layers, snap):
            l._noise_cur.copy_(a); l._u_cur.copy_(b)
    (o4, v4, u4), (o3, v3, u3) = res
    assert int(o4[0]) == int(o3[0])
    assert abs(o4[2] - o3[2]) <= RTOL * abs(o4[2]) and abs(o4[3] - o3[3]) <= RTOL * abs(o4[3])
    assert float((u3 - u4).abs().max()) <= 1e-9 * float(u4.abs().max())
    assert float((v3 - v4).abs().max()) <= 1e-9 * float(v4.abs().max())


def test_config2_full_size_chain_vs_oracle_golden(golden_dir, gemm_mode):
    """BASELINE configs[1] (n=32, 90 angles, 3 layers) against the ORACLE itself: tests/golden/oracle_config2_n32.npz holds two
    whole oracle iterations (oracle/make_golden.py --config2; ~35 s of CPU each, which is why they are precomputed).  Inputs
    and the initial chain state are regenerated here by the same deterministic oracle calls (seconds); the CUDA step runs under
    both complex product schemes.  Accept flags identical; Phi and all three half-samples within 1e-10 / 1e-9."""
    from mspde.core import MspdeContext
    g = np.load(os.path.join(golden_dir, "oracle_config2_n32.npz"))
    hp = o.Hyper(n_layers=3, n=32, n_ext=64, n_theta=90, beta=float(g["beta"]))
    pb = o.synthetic_problem(hp)
    ft, H, HtH, y, delta = pb["ft"], pb["H"], pb["HtH"], pb["y"], pb["delta"]
    assert np.allclose([y.sum(), np.abs(y).sum()], g["y_checksum"], rtol=1e-12) and abs(delta - float(g["delta"])) <= 1e-12 * delta
    pt = o.prior_tables(hp, ft)
    K, M = 3, H.shape[0]
    ns = o.NoiseStream(int(g["seed"]), K, ft.n_ravel, M)
    st = o.init_chain(hp, ft, [ns.half() for _ in range(K)], [ns.half() for _ in range(K)])
    ctx = MspdeContext(32, 64, M, K)
    ctx.set_inputs(H, HtH, y, ft.D_diag, pt["stdev_sym"], delta)
    cb = _load_chain(ctx, st)
    rec = torch.zeros((K, ft.n_ravel), dtype=torch.complex128, device=ctx.device)
    for it in range(len(g["accepted"])):
        nz = ns.draw_iteration()
        assert abs(nz["log_u"] - float(g["log_u"][it])) < 1e-15
        out = ctx.pcn_step(cb, [ctx.to_dev(w) for w in nz["w_half"]], nz["log_u"], ctx.to_dev(nz["w_last"]), ctx.to_dev(nz["e"]),
                           record=rec).cpu().numpy()
        ctx.check_info("step")
        assert int(out[0]) == int(g["accepted"][it])
        assert abs(out[2] - g["phi_cur"][it]) <= RTOL * abs(g["phi_cur"][it]), (it, out, g["phi_cur"][it])
        assert abs(out[3] - g["phi_new"][it]) <= RTOL * abs(g["phi_new"][it]), (it, out, g["phi_new"][it])
        assert abs(out[1] - g["log_ratio"][it]) <= RTOL * max(abs(g["phi_cur"][it]), abs(g["phi_new"][it]))
        assert rel(rec.cpu(), g["halves"][it]) < 1e-9                   # solves amplify rounding by cond(L) ~ 1e4-1e5
        for k in range(K):
            assert rel(cb.u_new[k].cpu(), g["u_new"][it][k]) < 1e-9, (it, k)
    ctx.close()


def test_block_distributed_phi_matches_single_gpu(toy):
    """mspde.dist.DistPhi (block-cyclic row layout, panel broadcasts) in a 1-rank NCCL group, several panels (nb=64),
    against the single-GPU Phi.  The multi-rank equality and the n=96 run are in profiles/r01_dist_phi_*.log."""
    import torch.distributed as dist
    from mspde.dist import BlockRows, DistPhi
    hp, pb, pt, ns, st, ctx0 = toy
    from mspde.core import MspdeContext
    hp8 = o.Hyper(n_layers=3, n=8, n_ext=64, n_theta=8)
    pb8 = o.synthetic_problem(hp8)
    ft, H, HtH, y, delta = pb8["ft"], pb8["H"], pb8["HtH"], pb8["y"], pb8["delta"]
    ctx = MspdeContext(8, 64, H.shape[0], 3)
    ctx.set_inputs(H, HtH, y, ft.D_diag, o.prior_tables(hp8, ft)["stdev_sym"], delta)
    rng = np.random.default_rng(2)
    u = o.symmetrize(o.w_half_from_normals(rng.standard_normal(ft.n_ravel), rng.standard_normal(ft.n_ravel))) * 3.0
    L = ctx.construct_L(ctx.to_dev(u), 1.0e3)
    ref = ctx.phi(L)
    created = False
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", str(29600 + os.getpid() % 300))
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda", 0))
        created = True
    try:
        dp = DistPhi(ctx.N, ctx.M, nb=64, device=ctx.device)
        hr = BlockRows(ctx.N, 64, 1, 0, ctx.device)
        for i in hr.owned:
            hr.block(i).copy_(ctx.HtH[i * 64:i * 64 + hr.rows(i), :])
        dp.setup(ctx.H, ctx.y, ctx.delta, HtH_rows=hr)
        val = dp.phi(L)
        assert abs(val - ref) <= RTOL * abs(ref), (val, ref)
        dp2 = DistPhi(ctx.N, ctx.M, nb=128, device=ctx.device)
        dp2.setup(ctx.H, ctx.y, ctx.delta)                 # HtH formed by the distributed Gram kernel
        assert abs(dp2.phi(L) - ref) <= 1e-9 * abs(ref)
        # block-distributed pivoted LU and Householder QR drivers (panel factor on the owner, broadcast, local apply)
        from mspde.dist import DistLU, DistQR
        w = crandn(ctx.N, seed=11)
        x1, ld1 = ctx.solve(L, w, want_logdet=True)
        x2, ld2 = DistLU(ctx.device).solve(L, w)
        assert torch.equal(x1, x2) and ld1 == ld2          # same kernels, same order: bit-identical
        rhs = crandn(ctx.M + ctx.N, seed=12)
        v1 = ctx.gibbs_lstsq(L, rhs)
        v2 = DistQR(ctx.device).lstsq([ctx.H, L], rhs)
        assert rel(v2.cpu(), v1.cpu()) < 1e-12             # same panel kernel; 64-wide updates there, rank-128 pairs on one GPU
    finally:
        if created:
            dist.destroy_process_group()


def test_drivers_run_end_to_end(tmp_path):
    """twoDsim.py (reference flag surface) and run_chains.py (config-4 style chain sharding + gather), single process."""
    import importlib.util
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mcmc-multispde_b200")

    def load(name):
        spec = importlib.util.spec_from_file_location(name, os.path.join(root, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    load("twoDsim").main(["--n-layers", "3", "--n", "4", "--n-theta", "8", "--n-samples", "6", "--beta", "0.5",
                          "--noprint-progress", "--adapt-sqrtbeta", "--result-dir", str(tmp_path)])
    files = [f for d, _, fs in os.walk(tmp_path) for f in fs]
    assert any(f.startswith("result_0") for f in files)
    out = str(tmp_path / "chains.npz")
    load("run_chains").main(["--chains", "3", "--n", "4", "--n-theta", "8", "--n-samples", "5", "--beta", "0.5", "--out", out])
    d = np.load(out)
    assert d["samples"].shape == (3, 3, 5, 25) and np.isfinite(d["samples"]).all()
    assert not np.array_equal(d["samples"][0], d["samples"][1])          # different seeds -> different chains


def test_config3_size_phi_solve_and_lstsq():
    """BASELINE configs[2] geometry (n=64: N=16129, M=11430): Phi against the torch.linalg rendition, the pivoted-LU solve
    against cuSOLVER, and the optimality of the Gibbs least-squares solution (the oracle is out of reach at this size:
    ~12 min per Phi on 8 host cores)."""
    from mspde import image as im
    sim = im.Simulation(n_layers=3, n_samples=2, n=64, n_extended=64, beta=0.3, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                        sigma_scaling=0.4, meas_std=0.2, evaluation_interval=10, printProgress=False, seed=2, burn_percentage=25.0,
                        enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png", n_theta=90)
    sim.pcn.set_chol_epsilon(1e-6)
    pcn, ctx = sim.pcn, sim.pcn.ctx
    pcn._ensure_inputs(sim.Layers[0].stdev_sym_host)
    N, M = ctx.N, ctx.M
    assert (N, M) == (16129, 11430)
    L = sim.Layers[2].LMat.current_L
    phi = ctx.phi(L, parts=True)
    Lc = L.contiguous()
    r = Lc.conj().T @ Lc
    r += ctx.HtH
    r.diagonal().add_(ctx.delta)
    c = torch.linalg.cholesky(r)
    del r
    Ht = torch.linalg.solve_triangular(c, ctx.Hct.contiguous(), upper=False)
    del c
    Qi = -(Ht.conj().T @ Ht)
    Qi.diagonal().add_(1.0)
    del Ht
    C = torch.linalg.cholesky(Qi)
    del Qi
    quad = 0.5 * float(torch.linalg.norm(ctx.y.to(torch.complex128) @ C) ** 2)
    logdet = float(torch.log(torch.diagonal(C).real).sum())
    del C
    ref = quad - logdet
    assert abs(phi[0] - ref) <= RTOL * abs(ref), (phi, ref)
    # dense solve of the middle layer
    L1 = sim.Layers[1].LMat.current_L
    w = sim.Layers[1].current_noise_sample
    x, ld = ctx.solve(L1, w, want_logdet=True)
    L1c = L1.contiguous()
    xr = torch.linalg.solve(L1c, w)
    assert float((x - xr).abs().max()) <= 1e-9 * float(xr.abs().max())
    assert abs(ld - float(torch.linalg.slogdet(L1c)[1])) <= RTOL * abs(ld)
    del L1c, xr
    # Gibbs least squares: normal-equation residual
    g = torch.Generator(device="cuda").manual_seed(9)
    rhs = torch.complex(torch.randn(M + N, dtype=torch.float64, device="cuda", generator=g),
                        torch.randn(M + N, dtype=torch.float64, device="cuda", generator=g))
    v = ctx.gibbs_lstsq(L, rhs)
    res_top = ctx.H @ v - rhs[:M]
    res_bot = Lc @ v - rhs[M:]
    grad = ctx.H.conj().T @ res_top + Lc.conj().T @ res_bot
    scale = (float(torch.linalg.norm(ctx.H)) ** 2 + float(torch.linalg.norm(Lc)) ** 2) * float(torch.linalg.norm(v))
    assert float(torch.linalg.norm(grad)) <= 1e-11 * scale
    ctx.close()


def test_presummed_3m_kernel_is_bit_identical(lib):
    """The packed-operand 3M kernel (mspde_set_gemm_pre_min: both operands repacked into the kernel's stage-tile image with
    the T3 operand planes formed once per matrix; the default for large products) performs the same arithmetic in the same
    order as the in-loop kernel: identical bits on edge shapes (odd sizes, k tails, beta != 0, upper) and on the k >= 512
    shapes it is used for."""
    from mspde import _ffi
    mode0, pre0 = _ffi.get_gemm_mode(), _ffi.get_gemm_pre_min()
    try:
        _ffi.set_gemm_mode(_ffi.GEMM_3M)
        for (m, n, k, conj, uplo, beta) in [(65, 33, 17, 1, 0, 0.0), (127, 63, 35, 0, 0, 1.0), (333, 333, 211, 1, 1, 1.0),
                                            (200, 1031, 129, 1, 0, -0.5), (1, 1, 1, 1, 0, 0.0), (961, 961, 961, 1, 1, 0.0),
                                            (700, 333, 513, 1, 0, 1.0), (1025, 1025, 1030, 0, 1, 0.0), (65, 2100, 777, 1, 0, -1.0)]:
            A = crandn(k, m, seed=1)
            B = A if uplo else crandn(k, n, seed=2)
            C0 = crandn(m, n, seed=3)
            outs = []
            for pre in (0, 1):
                _ffi.set_gemm_pre_min(pre)
                C = C0.clone()
                _gemm(lib, A, B, C, bool(conj), 0.75, beta, uplo)
                torch.cuda.synchronize()
                outs.append(C)
            assert torch.equal(outs[0], outs[1]), (m, n, k)
    finally:
        _ffi.set_gemm_mode(mode0)
        _ffi.set_gemm_pre_min(pre0)


@pytest.mark.parametrize("rows,cols", [(300, 129), (1100, 200), (2000, 449), (700, 700)])
def test_qr_block_reflector_widths_agree(rows, cols):
    """util.lstsq through the Householder QR with rank-128 block reflectors (two panels per trailing update, the default) and
    with rank-64 ones (every panel on its own): same panel kernel, solutions agree to rounding and match cuSOLVER's lstsq.
    Sizes cover an odd number of panels, a last panel of one column, and a square system."""
    from mspde import _ffi
    from mspde.core import MspdeContext
    n = int(np.ceil((np.sqrt(cols) + 1) / 2))
    while (2 * n - 1) ** 2 < cols:
        n += 1
    N = (2 * n - 1) ** 2
    ctx = MspdeContext(n, 64, max(rows - N, 1) + 64, 3)
    assert ctx.M + ctx.N >= rows and ctx.N >= cols
    a = crandn(rows, cols, seed=rows) + 0.5 * torch.eye(rows, cols, dtype=torch.complex128, device="cuda")
    b = crandn(rows, seed=cols)
    default = _ffi.get_qr_block()
    sols = {}
    try:
        for width in (64, 128):
            _ffi.set_qr_block(width)
            sols[width] = ctx.lstsq(a, b).clone()
    finally:
        _ffi.set_qr_block(default)
    ref = torch.linalg.lstsq(a, b.unsqueeze(1)).solution.squeeze(1)
    assert rel(sols[128].cpu(), ref.cpu()) < 1e-11 and rel(sols[64].cpu(), ref.cpu()) < 1e-11
    assert rel(sols[128].cpu(), sols[64].cpu()) < 1e-12
    ctx.close()


def test_multichain_runner_matches_chains_run_alone():
    """parallel.MultiChain (configs[3]: many chains per GPU through ONE shared context) against the same chains run alone:
    with injected noise every chain's accept flags, Phi values and samples are bit-identical, and the pooled Welford field
    statistics equal the statistics of all chains' samples taken together."""
    from mspde import image as im, parallel

    def make(seed, n_samples=5):
        sim = im.Simulation(n_layers=3, n_samples=n_samples, n=4, n_extended=64, beta=0.5, kappa=5e9, sigma_0=4e7, sigma_v=3.2e4,
                            sigma_scaling=0.4, meas_std=0.2, evaluation_interval=2, printProgress=False, seed=seed, burn_percentage=25.0,
                            enable_step_adaptation=True, pcn_variant="dunlop", phantom_name="shepp.png", n_theta=8)
        sim.pcn.set_chol_epsilon(1e-6)
        return sim
    ids, base = [0, 1, 2], 11
    sim = make(parallel.chain_seed(base, ids[0]))
    mc = parallel.MultiChain(sim, ids, base_seed=base, accumulate_stats=True)
    hp = o.Hyper(n_layers=3, n=4, n_ext=64, n_theta=8)
    ft = o.FourierTables(4, 64)
    streams = [o.NoiseStream(100 + c, 3, ft.n_ravel, sim.pcn.ctx.M) for c in ids]
    to = sim.pcn.ctx.to_dev
    # snapshot every chain's initial state, then step all chains together
    snaps = [[(l._noise_cur.clone(), l._u_cur.clone(), l._u_new.clone(), l.LMat.current_L.clone() if k else None)
              for k, l in enumerate(ch["Layers"])] for ch in mc.chains]
    noises = [[s.draw_iteration() for _ in range(4)] for s in streams]
    log = [[] for _ in ids]
    for it in range(4):
        nz = [([to(w) for w in noises[c][it]["w_half"]], noises[c][it]["log_u"], to(noises[c][it]["w_last"]), to(noises[c][it]["e"]))
              for c in range(len(ids))]
        betas = [ch["beta"] for ch in mc.chains]
        mc.step_all(noise=nz)
        for c, ch in enumerate(mc.chains):
            log[c].append((betas[c], [l._u_new.clone() for l in ch["Layers"]], [l._u_cur.clone() for l in ch["Layers"]]))
    acc_multi, steps_multi = mc.accepted_and_steps()
    assert steps_multi == 12
    # now every chain alone, in a fresh context, from the same initial state with the same noise and step sizes
    bad = []
    for c, cid in enumerate(ids):
        solo = make(parallel.chain_seed(base, ids[0]))      # same seed -> same synthetic data y as the shared context
        for k, l in enumerate(solo.Layers):
            nc, uc, un, Lc = snaps[c][k]
            l._noise_cur.copy_(nc); l._noise_new.copy_(nc); l._u_cur.copy_(uc); l._u_new.copy_(un)
            if k:
                l.LMat.current_L.copy_(Lc)
        for it in range(4):
            solo.pcn.beta, solo.pcn.betaZ = log[c][it][0], float(np.sqrt(max(0.0, 1 - log[c][it][0] ** 2)))
            n_ = noises[c][it]
            solo.pcn.one_step(solo.Layers, noise=([to(w) for w in n_["w_half"]], n_["log_u"], to(n_["w_last"]), to(n_["e"])))
            for k in range(3):
                if not torch.equal(solo.Layers[k]._u_new, log[c][it][1][k]):
                    bad.append((cid, it, k, "new", rel(solo.Layers[k]._u_new.cpu(), log[c][it][1][k].cpu())))
                if not torch.equal(solo.Layers[k]._u_cur, log[c][it][2][k]):
                    bad.append((cid, it, k, "cur", rel(solo.Layers[k]._u_cur.cpu(), log[c][it][2][k].cpu())))
        solo.pcn.ctx.close()
    assert not bad, bad
    # pooled statistics == statistics over all recorded samples of all chains
    pooled = mc.pooled_field_statistics()
    assert pooled[0]["count"] == 12
    fields = []
    for ch in mc.chains:
        lay = ch["Layers"][1]
        for r in range(lay.i_record):
            u = o.symmetrize(lay.samples_history[r])
            fields.append(o.irfft2(o.from_u_2D_ravel_to_uHalf_2D(u, 4), ft))
    F = np.stack(fields)
    assert rel(pooled[1]["u_mean"].cpu(), F.mean(0)) < 1e-11
    assert rel(pooled[1]["u_var"].cpu(), F.var(0)) < 1e-10
    sim.pcn.ctx.close()
