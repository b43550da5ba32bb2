"""Numerical behaviour of the 3M complex product scheme used by the default GEMM kernel (csrc/zgemm_tn.cu), restated in
numpy so it can be checked without a GPU:  T1 = Ar^T Br, T2 = Ai^T Bi, T3 = (Ar + Ai)^T (Br + Bi),
C = (T1 - T2) + i (T3 - T1 - T2).  Reference values come from extended precision (numpy longdouble).

What is asserted is what DESIGN.md claims: the normwise error of 3M is of the same order as that of the 4M scheme, for
well-scaled and for badly scaled (|imag| << |real|) operands; and the one thing 3M gives up -- the componentwise bound on
the imaginary part -- is visible in the badly scaled case."""
import numpy as np


def _products(A, B):
    Ar, Ai, Br, Bi = A.real, A.imag, B.real, B.imag
    c4 = (Ar.T @ Br - Ai.T @ Bi) + 1j * (Ar.T @ Bi + Ai.T @ Br)
    t1, t2, t3 = Ar.T @ Br, Ai.T @ Bi, (Ar + Ai).T @ (Br + Bi)
    c3 = (t1 - t2) + 1j * ((t3 - t1) - t2)
    L = np.longdouble
    Arl, Ail, Brl, Bil = Ar.astype(L), Ai.astype(L), Br.astype(L), Bi.astype(L)
    ref_r, ref_i = Arl.T @ Brl - Ail.T @ Bil, Arl.T @ Bil + Ail.T @ Brl
    return c3, c4, ref_r, ref_i


def _normwise(c, ref_r, ref_i, scale):
    err = np.sqrt(((c.real - ref_r) ** 2 + (c.imag - ref_i) ** 2).astype(np.float64))
    return float(err.max() / scale)


def test_3m_normwise_error_matches_4m_on_random_operands():
    rng = np.random.default_rng(0)
    k, m, n = 512, 48, 40
    A = rng.standard_normal((k, m)) + 1j * rng.standard_normal((k, m))
    B = rng.standard_normal((k, n)) + 1j * rng.standard_normal((k, n))
    c3, c4, rr, ri = _products(A, B)
    scale = float((np.abs(A).T @ np.abs(B)).max())              # the sum |a||b| of the bound
    e3, e4 = _normwise(c3, rr, ri, scale), _normwise(c4, rr, ri, scale)
    assert e4 < 1e-15 and e3 < 1e-15                              # both far below k*u
    assert e3 < 6 * e4 + 1e-17


def test_3m_keeps_normwise_accuracy_but_not_componentwise_on_badly_scaled_operands():
    rng = np.random.default_rng(1)
    k, m, n = 256, 32, 32
    A = rng.standard_normal((k, m)) + 1e-9j * rng.standard_normal((k, m))      # nearly real operands
    B = rng.standard_normal((k, n)) + 1e-9j * rng.standard_normal((k, n))
    c3, c4, rr, ri = _products(A, B)
    scale = float((np.abs(A).T @ np.abs(B)).max())
    assert _normwise(c3, rr, ri, scale) < 1e-15 and _normwise(c4, rr, ri, scale) < 1e-15
    # imaginary parts are O(1e-9 * sqrt(k)): 4M resolves them to ~1e-16 relative, 3M only to u * |real part| absolute
    rel3 = float(np.abs((c3.imag - ri).astype(np.float64)).max() / np.abs(ri.astype(np.float64)).max())
    rel4 = float(np.abs((c4.imag - ri).astype(np.float64)).max() / np.abs(ri.astype(np.float64)).max())
    assert rel4 < 1e-14
    assert 1e-9 < rel3 < 1e-4                                     # the documented loss; still u relative to the product's norm
