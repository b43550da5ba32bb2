"""CPU oracle for the pCN hot path of MCMC-MultiSPDE  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module.  The product path (``mcmc-multispde_b200/mspde``) never does, and
it fails loudly when the CUDA library is missing.

What this is
------------
A plain numpy/scipy FP64 restatement of the reference's per-iteration path (``pCN.one_step`` and what
it calls).  Every function cites the reference ``file:line`` it follows (paths relative to
``/root/reference``).  The reference itself is mixed precision (complex64 fields, complex128 solves);
this oracle restates its *semantics* in complex128/float64 throughout, which is the only meaningful
anchor for the 1e-10 parity target.

Pinning status: **parity unpinned by the reference** -- the reference tree holds no tests, golden
vectors or fixtures for this path, and its CuPy modules cannot be imported in this image (cupy,
skimage, h5py absent).  What *is* pinned, by ``oracle/make_golden.py`` + ``tests/test_oracle_pins.py``:
  * the BTTB index map of ``_construct_U`` and the entries of ``_calculate_H_Tomography`` are checked
    against the reference kernel bodies themselves, AST-extracted from
    ``/root/reference/mcmc/util_cupy.py`` and executed under ``NUMBA_ENABLE_CUDASIM=1`` (golden
    fixtures committed under ``tests/golden/``);
  * ``symmetrize``, ``symmetrize_2D``, ``extend2D`` are checked against ``mcmc/util_2D.py`` / ``util.py``;
  * Phi is checked through three algebraically independent formulas (Woodbury / naive / direct Q).
Third-party arithmetic behind the reference path (CuPy 7.0.0bc3 on CUDA 10.1: cuBLAS, cuSOLVER, cuFFT,
cuRAND) is restated with numpy.fft and LAPACK (OpenBLAS through numpy/scipy).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

ORDER = "F"                      # mcmc/image_cupy.py:24
PI32 = np.float32(np.pi)         # mcmc/util_cupy.py:22  (PI = cp.float32(cp.pi))
SQRT2_32 = np.float32(1.41421356)  # mcmc/util_cupy.py:21


# --------------------------------------------------------------------------------------------------
# a1: Fourier index tables and half/sym/2-D maps
# --------------------------------------------------------------------------------------------------
@dataclass
class FourierTables:
    """Index tables of ``FourierAnalysis_2D.__init__`` (mcmc/image_cupy.py:33-85)."""

    n: int
    n_ext: int
    il: int = 0
    N: int = 0
    n_ravel: int = 0
    m: int = 0
    ix: np.ndarray = field(default=None, repr=False)
    iy: np.ndarray = field(default=None, repr=False)
    D_diag: np.ndarray = field(default=None, repr=False)
    Kmatrix: np.ndarray = field(default=None, repr=False)

    def __post_init__(self):
        n = self.n
        self.il = 2 * n - 1                       # image_cupy.py:39
        self.N = self.il * self.il                # basis_number_2D_sym, image_cupy.py:39
        self.n_ravel = 2 * n * n - 2 * n + 1      # basis_number_2D_ravel, image_cupy.py:38
        self.m = 2 * self.n_ext - 1               # image side, image_cupy.py:883
        temp = np.arange(-(n - 1), n, dtype=np.int32)          # image_cupy.py:52
        self.ix, self.iy = np.meshgrid(temp, temp)             # image_cupy.py:57
        # image_cupy.py:63 -- (2*PI)**2 is evaluated in float32 (PI is a float32 scalar) and only then
        # promoted to float64 by the int32 array.
        c32 = np.float32(2) * PI32
        c32 = np.float32(c32 * c32)
        k2 = (self.ix.ravel(ORDER).astype(np.int64) ** 2 + self.iy.ravel(ORDER).astype(np.int64) ** 2)
        self.D_diag = -np.float64(c32) * k2.astype(np.float64)
        ones = np.ones_like(temp)
        self.Kmatrix = np.vstack((np.kron(temp, ones), np.kron(ones, temp)))   # image_cupy.py:73-74


def symmetrize(w_half: np.ndarray) -> np.ndarray:
    """mcmc/util_cupy.py:201-204 -- ``concat(conj(half[:0:-1]), half)``."""
    return np.concatenate((w_half[:0:-1].conj(), w_half))


def symmetrize_2D(uh: np.ndarray) -> np.ndarray:
    """mcmc/util_cupy.py:234-238."""
    w = uh[:, 1:]
    return np.hstack((w[::-1, :][:, ::-1].conj(), uh))


def from_u_2D_ravel_to_uHalf_2D(u_sym: np.ndarray, n: int) -> np.ndarray:
    """mcmc/util_cupy.py:245-246 -- ``reshape(il, il, 'F')[:, n-1:]``."""
    return u_sym.reshape(2 * n - 1, 2 * n - 1, order=ORDER)[:, n - 1:]


def extend2D(u_in: np.ndarray, num: int) -> np.ndarray:
    """mcmc/util_cupy.py:251-267 -- centred zero-pad to (2num-1)^2 (no-op when num <= n)."""
    if u_in.shape[1] != u_in.shape[0]:
        n = u_in.shape[1]
        if num > n:
            z = np.zeros((2 * num - 1, num), dtype=np.complex128)
            z[(num - 1) - (n - 1):(num - 1) + n, :n] = u_in
            return z
        return u_in
    n = (u_in.shape[0] + 1) // 2
    if num > n:
        z = np.zeros((2 * num - 1, 2 * num - 1), dtype=np.complex128)
        z[(num - 1) - (n - 1):(num - 1) + n, (num - 1) - (n - 1):(num - 1) + n] = u_in
        return z
    return u_in


# --------------------------------------------------------------------------------------------------
# a2: field synthesis (FFT conventions of image_cupy.py:108-129) and kappa functions
# --------------------------------------------------------------------------------------------------
def irfft2(uh: np.ndarray, ft: FourierTables) -> np.ndarray:
    """mcmc/image_cupy.py:116-129 -- ``Re(ifft2(ifftshift(pad(sym2D)))) * m``."""
    u2d = extend2D(symmetrize_2D(uh), ft.n_ext)
    u2d = np.fft.ifftshift(u2d)
    return np.fft.ifft2(u2d).real * (2 * ft.n_ext - 1)


def rfft2(z: np.ndarray, ft: FourierTables) -> np.ndarray:
    """mcmc/image_cupy.py:108-114 -- ``fftshift(fft2(z), axes=0)[m//2-(n-1):m//2+n, :n] / m``."""
    m = z.shape[0]
    t = np.fft.fft2(z.astype(np.complex128))
    t = np.fft.fftshift(t, axes=0)
    return t[m // 2 - (ft.n - 1):m // 2 + ft.n, :ft.n] / (2 * ft.n_ext - 1)


def kappa_pow_min_nu(u):
    """mcmc/util_cupy.py:270-274 (d=2, nu=1): exp(u)."""
    return np.exp(u)


def kappa_pow_d_per_2(u):
    """mcmc/util_cupy.py:277-281: exp(-u)."""
    return np.exp(-u)


def field_coeffs(u_sym: np.ndarray, ft: FourierTables):
    """The two symmetrised (2n-1)^2 coefficient tables used by the assembly:
    ``sym2D(rfft2(exp(+u(x))))`` and ``sym2D(rfft2(exp(-u(x))))``
    (mcmc/image_cupy.py:139-145 called from 171-175)."""
    uh = from_u_2D_ravel_to_uHalf_2D(u_sym, ft.n)
    img = irfft2(uh, ft)
    tp = symmetrize_2D(rfft2(kappa_pow_min_nu(img), ft))
    tm = symmetrize_2D(rfft2(kappa_pow_d_per_2(img), ft))
    return tp, tm


# --------------------------------------------------------------------------------------------------
# a3/a4: BTTB scatter and the Galerkin operator L(u)
# --------------------------------------------------------------------------------------------------
def bttb_index_map(n: int):
    """Closed form of the ``_construct_U`` scatter (mcmc/util_cupy.py:426-442).

    Returns ``(row_idx, col_idx, mask)`` of shape (N, N): ``U[m, n'] = uSym2D[row_idx, col_idx]`` where
    ``mask`` is true, 0 elsewhere.  ``row_idx = (k-l)+(n-1)``, ``col_idx = (i-j)+(n-1)`` with
    ``m = i*il+k`` and ``n' = j*il+l``.  Integer work: must be bit-exact."""
    il = 2 * n - 1
    idx = np.arange(il * il, dtype=np.int32)
    blk, inn = idx // il, idx % il
    d_ij = blk[:, None] - blk[None, :]
    d_kl = inn[:, None] - inn[None, :]
    mask = (d_ij >= -(n - 1)) & (d_ij < n) & (d_kl >= -(n - 1)) & (d_kl < n)
    row_idx = (d_kl + (n - 1)).astype(np.int32)
    col_idx = (d_ij + (n - 1)).astype(np.int32)
    return row_idx, col_idx, mask


def construct_U(u_sym2d: np.ndarray) -> np.ndarray:
    """mcmc/image_cupy.py:131-136 -> mcmc/util_cupy.py:323-330 (dense N x N, zero outside the band)."""
    il = u_sym2d.shape[0]
    n = (il + 1) // 2
    r, c, mask = bttb_index_map(n)
    U = np.zeros((il * il, il * il), dtype=np.complex128)
    U[mask] = u_sym2d[r[mask], c[mask]]
    return U


def assemble_L_from_tables(tp: np.ndarray, tm: np.ndarray, D_diag: np.ndarray, sqrt_beta: float) -> np.ndarray:
    """mcmc/image_cupy.py:179 -- ``L = (U[e^u] * D_diag - U[e^-u]) / sqrt_beta`` (column scaling;
    multiply, subtract, divide in that order -- SURVEY Appendix A.2)."""
    return (construct_U(tp) * D_diag - construct_U(tm)) / sqrt_beta


def assemble_L(u_sym: np.ndarray, sqrt_beta: float, ft: FourierTables) -> np.ndarray:
    """``Lmatrix_2D.construct_from`` (mcmc/image_cupy.py:231-233 -> 171-196)."""
    tp, tm = field_coeffs(u_sym, ft)
    return assemble_L_from_tables(tp, tm, ft.D_diag, float(sqrt_beta))


def logdet_L(L: np.ndarray) -> float:
    """``Lmatrix_2D.logDet`` (mcmc/image_cupy.py:268-276) through the patched complex slogdet
    (patch/norms.py:236-293): ``sum(log(abs(diag(LU))))`` after getrf; -inf on info != 0."""
    lu, piv, info = sla.lapack.zgetrf(np.asfortranarray(L))
    if info != 0:
        return -np.inf
    return float(np.sum(np.log(np.abs(np.diag(lu)))))


# --------------------------------------------------------------------------------------------------
# a14 / f1: tomography matrix H (Appendix A.8) and the fixed inputs
# --------------------------------------------------------------------------------------------------
def tomography_H(ft: FourierTables, n_theta: int) -> np.ndarray:
    """Un-normalised H, FP64.  ``Sinogram.get_measurement_matrix`` (mcmc/image_cupy.py:364-396) with the
    body of ``_calculate_H_Tomography`` (mcmc/util_cupy.py:445-470) in closed form."""
    n_r = ft.m
    theta = np.linspace(0.0, 180.0, n_theta, endpoint=False)            # image_cupy.py:354
    shifted = theta + theta[1]                                          # 365
    r = np.linspace(-0.5, 0.5, n_r, endpoint=False)[::-1]               # 358-361
    tg, rg = np.meshgrid(shifted * np.float64(PI32) / 180.0, r)         # 366 (util.PI is float32)
    rv, tv = rg.ravel(ORDER), tg.ravel(ORDER)
    M = rv.size
    kx = ft.ix.ravel(ORDER).astype(np.float64)
    ky = ft.iy.ravel(ORDER).astype(np.float64)
    th = tv[::-1] + 0.5 * math.pi                                       # util_cupy.py:457 (theta[-(m+1)])
    s, c = np.sin(th)[:, None], np.cos(th)[:, None]
    ku = kx[None, :] * c + ky[None, :] * s                              # 463
    kv = -kx[None, :] * s + ky[None, :] * c                             # 464
    ell = np.sqrt(0.25 - rv * rv)[:, None]                              # 465
    phase = np.exp(1j * math.pi * ((kx + ky)[None, :] - 2.0 * kv * rv[:, None]))
    with np.errstate(divide="ignore", invalid="ignore"):
        val = np.sin(2.0 * math.pi * ku * ell) / (math.pi * ku)        # 467
    H = np.where(kv * kv > 0.0, phase * val, phase * (2.0 * ell))       # 466-470
    H = H * (n_r / (2 * (n_r // 2)))                                    # image_cupy.py:377-378
    for b in range(n_theta // 4):                                       # 381-385
        H[b * n_r:(b + 1) * n_r, :] = np.roll(H[b * n_r:(b + 1) * n_r, :], 1, axis=0)
    assert M == n_r * n_theta
    return H


def normalised_inputs(H0: np.ndarray, meas_std: float):
    """``pCN.__init__`` normalisation (mcmc/image_cupy.py:437-442) and ``create_H_matrix`` (505-506):
    returns ``H = H0/sigma`` and ``HtH = 0.5*(H0^H H0 + (H0^H H0)^H)/sigma^2``."""
    t = H0.conj().T @ H0
    HtH0 = 0.5 * (t + t.conj().T)
    return H0 / meas_std, HtH0 / (meas_std ** 2)


def chol_stabilizer(HtH: np.ndarray, chol_epsilon: float) -> float:
    """``set_chol_epsilon`` -> ``reset_chol_stabilizer`` (mcmc/image_cupy.py:473-478): eps*||HtH||_F."""
    return float(chol_epsilon * np.linalg.norm(HtH))


# --------------------------------------------------------------------------------------------------
# a9: Phi(L) -- three algebraically independent forms
# --------------------------------------------------------------------------------------------------
def phi_woodbury(L, H, HtH, y, delta, return_parts=False):
    """Default ``pCN._log_ratio`` (mcmc/image_cupy.py:633-643)."""
    N = L.shape[0]
    temp = L.conj().T @ L                                               # 634
    r = temp + HtH + delta * np.eye(N)                                  # 636
    c = np.linalg.cholesky(r)                                           # 637
    Ht = sla.solve_triangular(c, H.conj().T, lower=True)                # 638 (general solve of a triangular c)
    Q_inv = np.eye(H.shape[0]) - Ht.conj().T @ Ht                       # 639
    C = np.linalg.cholesky(Q_inv)                                       # 642
    quad = 0.5 * np.linalg.norm(y @ C) ** 2
    logdet = np.sum(np.log(np.diag(C.real)))
    phi = quad - logdet                                                 # 643
    if return_parts:
        return float(phi), dict(r=r, c=c, Ht=Ht, Q_inv=Q_inv, C=C, quad=float(quad), logdet=float(logdet))
    return float(phi)


def phi_naive(L, H, y, delta=0.0):
    """``use_naive_logRatio_implementation`` branch (mcmc/image_cupy.py:644-657).  With ``delta=0`` this is the
    reference's formula verbatim (``Z = L^-H H^H``).  With ``delta>0`` the same algebra is applied to the
    stabilised prior precision ``A = L^H L + delta I`` (``Z = a^-1 H^H``, ``A = a a^H``) so that it stays an
    independent check of the stabilised Woodbury form: ``(I + Z^H Z)^-1 = I - H (A + H^H H)^-1 H^H``."""
    if delta == 0.0:
        Z = np.linalg.solve(L.conj().T, H.conj().T)                     # 645
    else:
        a = np.linalg.cholesky(L.conj().T @ L + delta * np.eye(L.shape[0]))
        Z = sla.solve_triangular(a, H.conj().T, lower=True)
    C = np.linalg.cholesky(Z.conj().T @ Z + np.eye(H.shape[0]))         # 656
    x = sla.solve_triangular(C, y.astype(np.complex128), lower=True)
    return float(0.5 * np.linalg.norm(x) ** 2 + np.sum(np.log(np.diag(C).real)))   # 657


def phi_direct(L, H, y, delta=0.0):
    """Direct ``Q = I + (L^-H H^H)^H (L^-H H^H)`` form of Legacy/image_cupy_N.py:527-540 (independent check),
    with the same optional stabiliser as :func:`phi_naive`."""
    A = L.conj().T @ L + delta * np.eye(L.shape[0])
    Q = np.eye(H.shape[0]) + H @ np.linalg.solve(A, H.conj().T)
    Q = 0.5 * (Q + Q.conj().T)
    sign, logdet = np.linalg.slogdet(Q)
    return float(0.5 * np.real(y @ np.linalg.solve(Q, y.astype(np.complex128))) + 0.5 * logdet)


# --------------------------------------------------------------------------------------------------
# a11: Gibbs least squares (util.lstsq, mcmc/util_cupy.py:590-609 + extra_linalg.solve_triangular)
# --------------------------------------------------------------------------------------------------
def lstsq_qr(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Reduced QR -> ``q^H b`` -> upper triangular back-substitution (mcmc/util_cupy.py:599-605,
    mcmc/extra_linalg.py:15-106 = cuBLAS ztrsm, upper, non-transposed, one rhs)."""
    q, r = np.linalg.qr(a)
    d = q.conj().T @ b
    return sla.solve_triangular(r, d, lower=False)


# --------------------------------------------------------------------------------------------------
# a6: noise
# --------------------------------------------------------------------------------------------------
def w_half_from_normals(re: np.ndarray, im: np.ndarray) -> np.ndarray:
    """``util.construct_w_Half`` (mcmc/util_cupy.py:34-39) with the normals injected:
    ``w = re + i im``; ``w[0] = 2*Re(w[0])``; ``w / sqrt(2)`` (SQRT2 is the float32 constant 1.41421356)."""
    w = re.astype(np.float64) + 1j * im.astype(np.float64)
    w[0] = 2.0 * w[0].real
    return w / np.float64(SQRT2_32)


class NoiseStream:
    """Injected proposal noise.  Per iteration, in the reference's RNG call order
    (mcmc/image_cupy.py:558, 573, 583, 585): (n_layers-1) half vectors, 1 uniform, 1 half vector, M normals."""

    def __init__(self, seed: int, n_layers: int, n_ravel: int, M: int):
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.n_layers, self.n_ravel, self.M = n_layers, n_ravel, M

    def half(self) -> np.ndarray:
        re = self.rng.standard_normal(self.n_ravel)
        im = self.rng.standard_normal(self.n_ravel)
        return w_half_from_normals(re, im)

    def draw_iteration(self) -> dict:
        w = [self.half() for _ in range(self.n_layers - 1)]
        log_u = float(np.log(self.rng.random()))
        w_last = self.half()
        e = self.rng.standard_normal(self.M)
        return dict(w_half=w, log_u=log_u, w_last=w_last, e=e)


# --------------------------------------------------------------------------------------------------
# Construction (Simulation.__init__, mcmc/image_cupy.py:839-946)
# --------------------------------------------------------------------------------------------------
@dataclass
class Hyper:
    n_layers: int = 3
    n: int = 32
    n_ext: int = 64
    n_theta: int = 90
    kappa: float = 5e9
    sigma_0: float = 4e7
    sigma_v: float = 3.2e4
    sigma_scaling: float = 0.4
    meas_std: float = 0.2
    beta: float = 1.0
    chol_epsilon: float = 1e-6
    evaluation_interval: int = 10
    target_acceptance_rate: float = 0.234
    beta_feedback_gain: float = 2.1


def prior_tables(hp: Hyper, ft: FourierTables):
    """sqrt-beta scalars and the layer-0 diagonal prior (mcmc/image_cupy.py:839-847, 869, 878-880)."""
    d = 2
    nu = 2 - d / 2
    alpha = nu + d / 2
    pi = np.float64(PI32)                               # util.PI is float32 (844)
    beta_u = (hp.sigma_0 ** 2) * (2 ** d * pi ** (d / 2) * math.gamma(alpha)) / math.gamma(nu)
    beta_v = beta_u * (hp.sigma_v / hp.sigma_0) ** 2
    sqrtBeta_v = np.float32(np.sqrt(beta_v))            # 846
    sqrtBeta_0 = np.float32(np.sqrt(beta_u))            # 847
    Lu_diag = ((ft.D_diag * hp.kappa ** (-nu) - hp.kappa ** (2 - nu) * np.ones(ft.N, np.float32))
               / np.float64(sqrtBeta_0)).astype(np.float32)            # 869
    stdev_sym = (-1.0 / Lu_diag.astype(np.float64))     # 878
    stdev_sym = stdev_sym.astype(np.float32).astype(np.float64)
    stdev_sym[ft.n_ravel - 1] /= 2                      # 879-880 (aliasing quirk A.4-1)
    sqrt_betas = [float(sqrtBeta_0)]
    for i in range(1, hp.n_layers):
        if i == hp.n_layers - 1:
            sqrt_betas.append(float(sqrtBeta_v))                                   # 913
        else:
            sqrt_betas.append(float(np.float64(sqrtBeta_v) * np.sqrt(hp.sigma_scaling)))   # 934
    return dict(sqrt_betas=sqrt_betas, stdev_sym=stdev_sym)


@dataclass
class ChainState:
    """Per-chain state of A.5 (all FP64 / complex128)."""

    beta: float
    betaZ: float
    sqrt_betas: list
    stdev_sym: np.ndarray
    noise_cur: list
    noise_new: list
    u_cur_sym: list
    u_new_sym: list
    L_cur: list
    L_latest: list
    accepted_total: int = 0
    iteration: int = 0
    # Lmatrix_2D.sqrt_beta per layer: fixed at construction (image_cupy.py:161-163) and the value construct_from divides by
    # (179), even after an accepted sqrt-beta move has changed Layer.sqrt_beta (703-704).  None = same as sqrt_betas.
    lmat_sqrt_betas: list = None


def init_chain(hp: Hyper, ft: FourierTables, init_w_half: list, init_solve_w_half: list) -> ChainState:
    """Initial state (A.4-11, A.6): layer 0 ``u = stdev_sym*w``; every other layer ``L = assemble(prev.u_cur)``
    and ``u = solve(L, w')`` (mcmc/image_cupy.py:735, 753-756, 905, 937).  ``init_w_half[k]`` is the draw that
    becomes ``current_noise_sample`` (735); ``init_solve_w_half[k]`` the separate draw used at 755/905."""
    pt = prior_tables(hp, ft)
    K = hp.n_layers
    st = ChainState(beta=float(hp.beta), betaZ=float(np.sqrt(max(0.0, 1 - hp.beta ** 2))),
                    sqrt_betas=pt["sqrt_betas"], stdev_sym=pt["stdev_sym"],
                    noise_cur=[None] * K, noise_new=[None] * K, u_cur_sym=[None] * K, u_new_sym=[None] * K,
                    L_cur=[None] * K, L_latest=[None] * K)
    for k in range(K):
        st.noise_cur[k] = symmetrize(init_w_half[k])
        st.noise_new[k] = st.noise_cur[k].copy()
        if k == 0:
            st.u_cur_sym[0] = st.stdev_sym * symmetrize(init_solve_w_half[0])
        else:
            st.L_cur[k] = assemble_L(st.u_cur_sym[k - 1], st.sqrt_betas[k], ft)
            st.L_latest[k] = st.L_cur[k]
            st.u_cur_sym[k] = np.linalg.solve(st.L_cur[k], symmetrize(init_solve_w_half[k]))
        st.u_new_sym[k] = st.u_cur_sym[k].copy()
    return st


# --------------------------------------------------------------------------------------------------
# a6-a13: one pCN step (pCN.one_step, mcmc/image_cupy.py:553-617)
# --------------------------------------------------------------------------------------------------
def one_step(st: ChainState, ft: FourierTables, H, HtH, y, delta, noise: dict, timers: dict | None = None):
    """One iteration with injected noise.  Returns a dict of per-step diagnostics."""
    import time
    K = len(st.noise_cur)
    N, M = ft.N, H.shape[0]
    lsb = st.sqrt_betas if st.lmat_sqrt_betas is None else st.lmat_sqrt_betas           # LMat.sqrt_beta (construct_from, 179)
    t0 = time.perf_counter()
    for k in range(K - 1):                                                               # 557
        st.noise_new[k] = st.betaZ * st.noise_cur[k] + st.beta * symmetrize(noise["w_half"][k])   # 558 -> 790
        if k == 0:
            st.u_new_sym[0] = st.stdev_sym * st.noise_new[0]                              # 563
        else:
            st.L_latest[k] = assemble_L(st.u_new_sym[k - 1], lsb[k], ft)                  # 560
            st.u_new_sym[k] = np.linalg.solve(st.L_latest[k], st.noise_new[k])            # 561
    t1 = time.perf_counter()
    phi_cur = phi_woodbury(st.L_cur[K - 1], H, HtH, y, delta)                             # 567-568
    t2 = time.perf_counter()
    st.L_latest[K - 1] = assemble_L(st.u_new_sym[K - 2], lsb[K - 1], ft)                  # 570
    t3 = time.perf_counter()
    phi_new = phi_woodbury(st.L_latest[K - 1], H, HtH, y, delta)                          # 571
    t4 = time.perf_counter()
    log_ratio = phi_cur - phi_new
    accepted = bool(log_ratio > noise["log_u"])                                           # 573 (strict)
    if accepted:
        for k in range(K):                                                                # 575-579
            st.u_cur_sym[k] = st.u_new_sym[k].copy()                                      # 803-804
            st.noise_cur[k] = st.noise_new[k].copy()                                      # 807
            if k >= 1:
                st.L_cur[k] = st.L_latest[k]                                              # 279-280
    # Gibbs part (583-599): uses noise_cur AFTER the accept copy and L_latest even after a rejection (A.4-3)
    st.noise_new[K - 1] = st.betaZ * st.noise_cur[K - 1] + st.beta * symmetrize(noise["w_last"])   # 583
    wBar = np.concatenate((noise["e"].astype(np.complex128), st.noise_new[K - 1]))        # 585-586
    yBar = np.concatenate((y, np.zeros(N)))                                               # 449
    LBar = np.vstack((H, st.L_latest[K - 1]))                                             # 596
    st.u_new_sym[K - 1] = lstsq_qr(LBar, yBar - wBar)                                     # 597-598
    t5 = time.perf_counter()
    st.accepted_total += int(accepted)
    st.iteration += 1
    if timers is not None:
        timers["layers"] = timers.get("layers", 0.0) + (t1 - t0)
        timers["phi"] = timers.get("phi", 0.0) + (t2 - t1) + (t4 - t3)
        timers["assembly_last"] = timers.get("assembly_last", 0.0) + (t3 - t2)
        timers["gibbs"] = timers.get("gibbs", 0.0) + (t5 - t4)
    return dict(accepted=accepted, log_ratio=float(log_ratio), phi_cur=float(phi_cur), phi_new=float(phi_new),
                log_u=float(noise["log_u"]))


def one_step_for_sqrt_betas(st: ChainState, ft: FourierTables, H, HtH, y, delta, noise: dict, lmat_sqrt_betas: list,
                            pcn_step_sqrt_betas: float = 0.5, stdev_sqrt_betas=None):
    """Optional hyper-parameter move ``pCN.one_step_for_sqrtBetas`` (mcmc/image_cupy.py:666-710), restated with its quirks:
    line 673 uses ``sqrt(1 - s^2)`` (not the stored ``pcn_step_sqrtBetas_Z``); line 695 re-assembles the last layer with the
    Lmatrix's own ``sqrt_beta`` (``lmat_sqrt_betas``: fixed at construction, never updated by an acceptance); an acceptance
    (699-708) updates ``sqrt_beta``/``current_L``/``stdev_sym`` but NOT the current samples.
    noise: dict(z = n_layers normals, e = M normals, log_u)."""
    K, N = len(st.noise_cur), ft.N
    if st.lmat_sqrt_betas is None:
        st.lmat_sqrt_betas = list(lmat_sqrt_betas)       # from now on plain steps keep assembling with these (quirk, 179 vs 704)
    s = pcn_step_sqrt_betas
    sd = np.ones(K) if stdev_sqrt_betas is None else np.asarray(stdev_sqrt_betas)
    z = sd * noise["z"]                                                                            # 668
    prop = np.zeros(K)
    stdev_sym_temp = None
    for i in range(K):
        prop[i] = np.exp(np.sqrt(1 - s ** 2) * np.log(st.sqrt_betas[i]) + s * z[i])                # 673-674
        if i == 0:
            stdev_sym_temp = (prop[0] / st.sqrt_betas[0]) * st.stdev_sym                           # 676
            st.u_new_sym[0] = stdev_sym_temp * st.noise_cur[0]                                     # 677
        else:
            st.L_latest[i] = assemble_L(st.u_new_sym[i - 1], prop[i], ft)                          # 679
            if i < K - 1:
                st.u_new_sym[i] = np.linalg.solve(st.L_latest[i], st.noise_cur[i])                 # 681
            else:
                wBar = np.concatenate((noise["e"].astype(np.complex128), st.noise_cur[K - 1]))     # 683-685
                yBar = np.concatenate((y, np.zeros(N)))
                st.u_new_sym[K - 1] = lstsq_qr(np.vstack((H, st.L_latest[K - 1])), yBar - wBar)    # 686-688
    phi_cur = phi_woodbury(st.L_cur[K - 1], H, HtH, y, delta)                                      # 692-693
    st.L_latest[K - 1] = assemble_L(st.u_new_sym[K - 2], lmat_sqrt_betas[K - 1], ft)               # 695 (LMat's own sqrt_beta)
    phi_new = phi_woodbury(st.L_latest[K - 1], H, HtH, y, delta)                                   # 696
    log_ratio = phi_cur - phi_new
    accepted = bool(log_ratio > noise["log_u"])                                                    # 699
    if accepted:
        for i in range(K):                                                                         # 703-708
            st.sqrt_betas[i] = float(prop[i])
            if i >= 1:
                st.L_cur[i] = st.L_latest[i]
        st.stdev_sym = stdev_sym_temp
    return dict(accepted=accepted, log_ratio=float(log_ratio), phi_cur=float(phi_cur), phi_new=float(phi_new), prop=prop)


def adapt_step(st: ChainState, hp: Hyper):
    """``Simulation.run`` step adaptation (mcmc/image_cupy.py:975-982) -> ``pCN.adapt_step/set_step`` (528-540):
    cumulative acceptance rate; update ignored outside (1e-7, 1)."""
    acc = st.accepted_total / st.iteration
    b = st.beta * math.exp(hp.beta_feedback_gain * (acc - hp.target_acceptance_rate))
    if 1e-7 < b < 1:
        st.beta = b
        st.betaZ = math.sqrt(1 - b * b)


def current_halves(st: ChainState, ft: FourierTables):
    """``Layer.record_sample`` payload (mcmc/image_cupy.py:792-795): ``current_sample = sym[n_ravel-1:]``."""
    return [u[ft.n_ravel - 1:].copy() for u in st.u_cur_sym]


# --------------------------------------------------------------------------------------------------
# Synthetic problem: H by formula, y = Re(H sym(v_true)) + N(0,1)
# --------------------------------------------------------------------------------------------------
def synthetic_phantom(m: int) -> np.ndarray:
    """Analytic Shepp-Logan-like phantom on an m x m grid (ellipses); used when shepp.png is not available
    (the GPU box has no /root/reference).  Any real image exercises identical arithmetic."""
    yy, xx = np.mgrid[-1:1:m * 1j, -1:1:m * 1j]
    img = np.zeros((m, m))
    ell = [(1.0, .69, .92, 0, 0, 0), (-.8, .6624, .8740, 0, -.0184, 0), (-.2, .11, .31, .22, 0, -18),
           (-.2, .16, .41, -.22, 0, 18), (.1, .21, .25, 0, .35, 0), (.1, .046, .046, 0, .1, 0),
           (.1, .046, .046, 0, -.1, 0), (.1, .046, .023, -.08, -.605, 0), (.1, .023, .023, 0, -.606, 0),
           (.1, .023, .046, .06, -.605, 0)]
    for a, rx, ry, x0, y0, ang in ell:
        t = np.deg2rad(ang)
        xr = (xx - x0) * np.cos(t) + (yy - y0) * np.sin(t)
        yr = -(xx - x0) * np.sin(t) + (yy - y0) * np.cos(t)
        img[(xr / rx) ** 2 + (yr / ry) ** 2 <= 1] += a
    return img[::-1]


def synthetic_problem(hp: Hyper, seed: int = 1, phantom: np.ndarray | None = None):
    """Fixed inputs shared by the oracle and the CUDA path: tables, H, HtH, y, delta."""
    ft = FourierTables(hp.n, hp.n_ext)
    H0 = tomography_H(ft, hp.n_theta)
    H, HtH = normalised_inputs(H0, hp.meas_std)
    delta = chol_stabilizer(HtH, hp.chol_epsilon)
    if phantom is None:
        phantom = synthetic_phantom(ft.m)
    v_half = rfft2(phantom, ft)
    v_true = symmetrize_2D(v_half).ravel(ORDER)
    rng = np.random.Generator(np.random.PCG64(seed))
    y = (H @ v_true).real + rng.standard_normal(H.shape[0])
    return dict(ft=ft, H=H, HtH=HtH, y=y, delta=delta, v_true=v_true)
