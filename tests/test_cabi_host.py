"""CPU tests of the boundary: libmspde.so loads, exports every symbol include/mspde.h declares, and its host-side
entry points (index tables, BTTB index map) are bit-exact against the oracle / golden vectors.  No compute calls."""
import ctypes
import os
import re

import numpy as np
import pytest

from mspde import _ffi
from oracle import mspde_oracle as o

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "mspde.h")).read()
    names = sorted(set(re.findall(r"\b(mspde_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    lib = _ffi.lib()
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_last_error_is_a_string():
    assert isinstance(_ffi.lib().mspde_last_error(), bytes)


@pytest.mark.parametrize("n", [2, 3, 4, 16, 32])
def test_index_tables_bit_exact(n):
    il, N = 2 * n - 1, (2 * n - 1) ** 2
    ix = np.empty((il, il), np.int32); iy = np.empty((il, il), np.int32)
    km = np.empty((2, N), np.int32); D = np.empty(N, np.float64)
    _ffi.check(_ffi.lib().mspde_index_maps(n, _ffi.ptr(ix), _ffi.ptr(iy), _ffi.ptr(km), _ffi.ptr(D)))
    ft = o.FourierTables(n, max(n, 8))
    assert np.array_equal(ix, ft.ix) and np.array_equal(iy, ft.iy)
    assert np.array_equal(km, ft.Kmatrix)
    assert np.array_equal(D, ft.D_diag)            # float32 (2 pi)^2 promoted to float64: bit-exact


@pytest.mark.parametrize("n", [2, 3, 4])
def test_bttb_index_map_bit_exact_vs_reference_golden(golden_dir, n):
    g = np.load(os.path.join(golden_dir, f"ref_construct_U_n{n}.npz"))
    il, N = 2 * n - 1, (2 * n - 1) ** 2
    r = np.empty((N, N), np.int32); c = np.empty((N, N), np.int32); m = np.empty((N, N), np.uint8)
    _ffi.check(_ffi.lib().mspde_bttb_index_map(n, _ffi.ptr(r), _ffi.ptr(c), _ffi.ptr(m)))
    ro, co, mo = o.bttb_index_map(n)
    assert np.array_equal(r, ro) and np.array_equal(c, co) and np.array_equal(m.astype(bool), mo)
    gathered = np.where(m.astype(bool), g["labels"][np.clip(r, 0, il - 1), np.clip(c, 0, il - 1)], 0)
    assert np.array_equal(gathered, g["U"])


def test_error_mapping():
    _ffi.check(0)
    with pytest.raises(np.linalg.LinAlgError):
        _ffi.check(3, "potrf")
    with pytest.raises(_ffi.MspdeError):
        _ffi.check(-1, "cuda")


def test_ctx_create_rejects_bad_arguments_without_touching_the_gpu():
    h = ctypes.c_void_p()
    rc = _ffi.lib().mspde_ctx_create(ctypes.byref(h), 0, 1, 1, 10, 3, None)    # n < 2
    assert rc < 0 and b"bad arguments" in _ffi.lib().mspde_last_error()


def test_gemm_mode_switch_is_host_only():
    """mspde_set_gemm_mode / mspde_get_gemm_mode (include/mspde.h): 3M is the default, 4M selectable, anything else rejected."""
    default = _ffi.get_gemm_mode()
    if not os.environ.get("MSPDE_GEMM"):
        assert default == _ffi.GEMM_3M
    try:
        _ffi.set_gemm_mode(_ffi.GEMM_4M)
        assert _ffi.get_gemm_mode() == _ffi.GEMM_4M
        with pytest.raises(_ffi.MspdeError):
            _ffi.set_gemm_mode(7)
        assert _ffi.get_gemm_mode() == _ffi.GEMM_4M
    finally:
        _ffi.set_gemm_mode(default)


def test_process_wide_knobs_are_host_only():
    """mspde_set/get_gemm_pre_min, mspde_set/get_qr_block and mspde_release_gemm_workspaces (include/mspde.h) work without a
    device: the threshold round-trips (negative values clamp to 0 = never), the block width accepts 64 and 128 only, and
    releasing when nothing was ever allocated returns 0 bytes."""
    pre0, qr0 = _ffi.get_gemm_pre_min(), _ffi.get_qr_block()
    try:
        if not os.environ.get("MSPDE_GEMM_PRE_MIN"):
            assert pre0 == 2048
        _ffi.set_gemm_pre_min(0)
        assert _ffi.get_gemm_pre_min() == 0
        _ffi.set_gemm_pre_min(-5)
        assert _ffi.get_gemm_pre_min() == 0
        _ffi.set_gemm_pre_min(4096)
        assert _ffi.get_gemm_pre_min() == 4096
        for w in (64, 128):
            _ffi.set_qr_block(w)
            assert _ffi.get_qr_block() == w
        with pytest.raises(_ffi.MspdeError):
            _ffi.set_qr_block(96)
        assert _ffi.release_gemm_workspaces() == 0
    finally:
        _ffi.set_gemm_pre_min(pre0)
        _ffi.set_qr_block(qr0)
