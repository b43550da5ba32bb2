"""GPU test (-m gpu): the packed-operand GEMM workspaces are returned on request (include/mspde.h:
mspde_release_gemm_workspaces) and the next large product allocates afresh with identical results."""
import ctypes
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200"))

pytestmark = pytest.mark.gpu


def crandn(*shape, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return torch.complex(torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g),
                         torch.randn(*shape, dtype=torch.float64, device="cuda", generator=g))


def _gemm(A, B, C):
    """C = conj(A)^T B through mspde_zgemm_tn on torch's current stream (A: k x m, B: k x n, row-major)."""
    from mspde import _ffi
    k, m = A.shape
    n = B.shape[1]
    _ffi.check(_ffi.lib().mspde_zgemm_tn(_ffi.current_stream(), 1, m, n, k, ctypes.c_double(1.0), _ffi.ptr(A), A.stride(0),
                                         _ffi.ptr(B), B.stride(0), ctypes.c_double(0.0), _ffi.ptr(C), C.stride(0), 0, _ffi.ptr(None),
                                         1), "zgemm_tn")


def test_packed_gemm_workspaces_are_returned():
    """The packed-operand kernel keeps one workspace per stream (zgemm_tn.cu: prews_get).  mspde_release_gemm_workspaces waits
    for the device and gives all of them back; the next large product allocates afresh and produces the same bits."""
    from mspde import _ffi
    mode0, pre0 = _ffi.get_gemm_mode(), _ffi.get_gemm_pre_min()
    try:
        _ffi.set_gemm_mode(_ffi.GEMM_3M)
        _ffi.set_gemm_pre_min(1)
        A, B = crandn(600, 300, seed=4), crandn(600, 200, seed=5)
        C1, C2 = crandn(300, 200, seed=6), crandn(300, 200, seed=6)
        _gemm(A, B, C1)
        freed = _ffi.release_gemm_workspaces()
        assert freed >= (600 // 16) * 16 * (300 + 200) * 3 * 8          # at least this product's packed operands
        assert _ffi.release_gemm_workspaces() == 0                         # nothing left to return
        _gemm(A, B, C2)
        torch.cuda.synchronize()
        assert torch.equal(C1, C2)
        want = A.conj().transpose(0, 1) @ B
        assert float((C1 - want).abs().max()) <= 1e-11 * float(want.abs().max())
    finally:
        _ffi.set_gemm_mode(mode0)
        _ffi.set_gemm_pre_min(pre0)
        _ffi.release_gemm_workspaces()
