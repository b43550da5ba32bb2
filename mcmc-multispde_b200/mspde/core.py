"""Thin object wrapper over the libmspde C-ABI: one ``MspdeContext`` per (device, n, n_ext, M, n_layers).

torch is used only for device memory and streams; every numerical operation is a C-ABI call."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _ffi

c_double = ctypes.c_double
CDTYPE = torch.complex128


def _round_up(x, a):
    return (x + a - 1) // a * a


class MspdeContext:
    def __init__(self, n: int, n_ext: int, M: int, n_layers: int, device: int = 0):
        self.lib = _ffi.lib()
        self.device = torch.device("cuda", device)
        self.n, self.n_ext, self.M, self.n_layers = n, n_ext, M, n_layers
        self.il = 2 * n - 1
        self.N = self.il * self.il
        self.n_ravel = 2 * n * n - 2 * n + 1
        self.m = 2 * n_ext - 1
        self.ldN = _round_up(self.N, 8)
        self.ldM = _round_up(M, 8)
        self._h = ctypes.c_void_p()
        torch.cuda.set_device(self.device)
        _ffi.check(self.lib.mspde_ctx_create(ctypes.byref(self._h), device, n, n_ext, M, n_layers, _ffi.current_stream()),
                   "mspde_ctx_create")
        self._keep = {}
        self.scal = torch.zeros(16, dtype=torch.float64, device=self.device)

    def close(self):
        if self._h:
            self.lib.mspde_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- inputs ---------------------------------------------------------------------------------------
    def to_dev(self, a, dtype=None):
        t = torch.as_tensor(np.ascontiguousarray(a))
        if dtype is not None:
            t = t.to(dtype)
        return t.to(self.device)

    def new_matrix(self, rows: int, cols: int):
        """rows x cols complex128 matrix with a padded leading dimension; returns the [rows, cols] view."""
        ld = _round_up(cols, 8)
        return torch.zeros((rows, ld), dtype=CDTYPE, device=self.device)[:, :cols]

    def set_inputs(self, H, HtH, y, D_diag, stdev_sym, delta: float):
        """H (M x N), HtH (N x N), y (M), D_diag (N), stdev_sym (N): numpy or torch; copied to the device once."""
        N, M = self.N, self.M
        Hd = self.new_matrix(M, N); Hd.copy_(self.to_dev(H, CDTYPE))
        Hct = self.new_matrix(N, M); Hct.copy_(Hd.conj().transpose(0, 1))      # H_conj_T, image_cupy.py:441
        HtHd = self.new_matrix(N, N); HtHd.copy_(self.to_dev(HtH, CDTYPE))
        yd = self.to_dev(y, torch.float64)
        Dd = self.to_dev(D_diag, torch.float64)
        sd = self.to_dev(stdev_sym, torch.float64)
        self._keep.update(H=Hd, Hct=Hct, HtH=HtHd, y=yd, D=Dd, stdev=sd)
        self.H, self.Hct, self.HtH, self.y, self.D_diag, self.stdev_sym, self.delta = Hd, Hct, HtHd, yd, Dd, sd, float(delta)
        _ffi.check(self.lib.mspde_ctx_set_inputs(self._h, _ffi.ptr(Hd), Hd.stride(0), _ffi.ptr(Hct), Hct.stride(0),
                                                 _ffi.ptr(HtHd), HtHd.stride(0), _ffi.ptr(yd), _ffi.ptr(Dd), _ffi.ptr(sd),
                                                 c_double(float(delta))), "mspde_ctx_set_inputs")

    def _sync_stream(self):
        _ffi.check(self.lib.mspde_ctx_set_stream(self._h, _ffi.current_stream()), "set_stream")

    # ---- operators ------------------------------------------------------------------------------------
    def field_coeffs(self, u_sym):
        self._sync_stream()
        tp = torch.empty((self.il, self.il), dtype=CDTYPE, device=self.device)
        tm = torch.empty_like(tp)
        _ffi.check(self.lib.mspde_field_coeffs(self._h, _ffi.ptr(u_sym), _ffi.ptr(tp), _ffi.ptr(tm)), "field_coeffs")
        return tp, tm

    def assemble_L(self, tp, tm, sqrt_beta: float, out=None):
        self._sync_stream()
        L = self.new_matrix(self.N, self.N) if out is None else out
        _ffi.check(self.lib.mspde_assemble_L(self._h, _ffi.ptr(tp), _ffi.ptr(tm), c_double(float(sqrt_beta)), _ffi.ptr(L),
                                             L.stride(0)), "assemble_L")
        return L

    def construct_L(self, u_sym, sqrt_beta: float, out=None):
        self._sync_stream()
        L = self.new_matrix(self.N, self.N) if out is None else out
        _ffi.check(self.lib.mspde_construct_L(self._h, _ffi.ptr(u_sym), c_double(float(sqrt_beta)), _ffi.ptr(L),
                                              L.stride(0)), "construct_L")
        return L

    def phi(self, L, parts: bool = False):
        """Phi(L) of pCN._log_ratio; synchronises to return Python floats."""
        self._sync_stream()
        _ffi.check(self.lib.mspde_phi(self._h, _ffi.ptr(L), L.stride(0), _ffi.ptr(self.scal)), "phi")
        self.check_info("phi (Cholesky)")
        v = self.scal[:3].cpu().numpy()
        return (float(v[0]), float(v[1]), float(v[2])) if parts else float(v[0])

    def check_info(self, what: str):
        info = self.lib.mspde_read_info(self._h, 1)
        _ffi.check(info, what)

    def buffer(self, name: str, rows: int, cols: int):
        """Strided torch view of an internal workspace (parity tests only)."""
        ld = ctypes.c_int(0)
        self.lib.mspde_ctx_buffer.restype = ctypes.c_void_p
        p = self.lib.mspde_ctx_buffer(self._h, name.encode(), ctypes.byref(ld))
        if not p:
            raise _ffi.MspdeError(self.lib.mspde_last_error().decode())
        return _view_from_ptr(p, rows, cols, ld.value, self.device)


def _view_from_ptr(p: int, rows: int, cols: int, ld: int, device) -> torch.Tensor:
    """Zero-copy complex128 view over raw device memory via __cuda_array_interface__."""

    class _Raw:
        pass

    raw = _Raw()
    raw.__cuda_array_interface__ = {"shape": (rows * ld * 2,), "typestr": "<f8", "data": (int(p), False), "version": 2}
    flat = torch.as_tensor(raw, device=device)
    return torch.view_as_complex(flat.view(rows, ld, 2))[:, :cols]
