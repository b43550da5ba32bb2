"""Thin object wrapper over the libmspde C-ABI: one ``MspdeContext`` per (device, n, n_ext, M, n_layers).

torch is used only for device memory and streams; every numerical operation is a C-ABI call."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _ffi

c_double = ctypes.c_double
CDTYPE = torch.complex128

STEP_CACHE_PHI_CUR = 1
STEP_SKIP_GIBBS = 2
STEP_SERIAL = 4
STEP_FAST_PHI = 8
STEP_NAIVE_PHI = 16


class _Chain(ctypes.Structure):
    """mspde_chain of include/mspde.h."""
    _fields_ = [("n_layers", ctypes.c_int), ("beta", c_double), ("betaZ", c_double),
                ("sqrt_betas", ctypes.POINTER(c_double)),
                ("noise_cur", ctypes.POINTER(ctypes.c_void_p)), ("noise_new", ctypes.POINTER(ctypes.c_void_p)),
                ("u_cur", ctypes.POINTER(ctypes.c_void_p)), ("u_new", ctypes.POINTER(ctypes.c_void_p)),
                ("L_cur", ctypes.POINTER(ctypes.c_void_p)), ("L_latest", ctypes.POINTER(ctypes.c_void_p)),
                ("ldl", ctypes.c_int), ("phi_cur_valid", ctypes.c_int), ("phi_cache", ctypes.c_void_p)]


class _StepNoise(ctypes.Structure):
    """mspde_step_noise of include/mspde.h."""
    _fields_ = [("w_half", ctypes.POINTER(ctypes.c_void_p)), ("w_last", ctypes.c_void_p),
                ("e", ctypes.c_void_p), ("log_u", c_double)]


def _ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


def _round_up(x, a):
    return (x + a - 1) // a * a


class MspdeContext:
    def __init__(self, n: int, n_ext: int, M: int, n_layers: int, device: int = 0, assembly_only: bool = False):
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
        _ffi.check(self.lib.mspde_ctx_create_ex(ctypes.byref(self._h), device, n, n_ext, M, n_layers, _ffi.current_stream(),
                                                1 if assembly_only else 0), "mspde_ctx_create")
        self._keep = {}
        self.scal = torch.zeros(16, dtype=torch.float64, device=self.device)
        self.generation = 0          # bumped whenever H / y / delta / stdev_sym change: invalidates every chain's cached Phi

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
        t = a if isinstance(a, torch.Tensor) else torch.as_tensor(np.ascontiguousarray(a))
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
        self.generation += 1
        _ffi.check(self.lib.mspde_ctx_set_inputs(self._h, _ffi.ptr(Hd), Hd.stride(0), _ffi.ptr(Hct), Hct.stride(0),
                                                 _ffi.ptr(HtHd), HtHd.stride(0), _ffi.ptr(yd), _ffi.ptr(Dd), _ffi.ptr(sd),
                                                 c_double(float(delta))), "mspde_ctx_set_inputs")

    def update_scalars(self, stdev_sym=None, delta=None):
        """Re-register the inputs after a change of the diagonal prior or the Cholesky stabiliser (no matrix copies)."""
        if stdev_sym is not None:
            self.stdev_sym = self.to_dev(stdev_sym, torch.float64)
            self._keep["stdev"] = self.stdev_sym
        if delta is not None:
            self.delta = float(delta)
        self.generation += 1
        _ffi.check(self.lib.mspde_ctx_set_inputs(self._h, _ffi.ptr(self.H), self.H.stride(0), _ffi.ptr(self.Hct),
                                                 self.Hct.stride(0), _ffi.ptr(self.HtH), self.HtH.stride(0), _ffi.ptr(self.y),
                                                 _ffi.ptr(self.D_diag), _ffi.ptr(self.stdev_sym), c_double(self.delta)),
                   "mspde_ctx_set_inputs")

    def gemm_tn(self, A, B, conjA=True, alpha=1.0, beta=0.0, C=None, upper=False):
        """C = alpha * op(A)^T B + beta * C with A (k x m), B (k x n) row-major (mspde_zgemm_tn)."""
        k, m = A.shape
        n = B.shape[1]
        if C is None:
            C = self.new_matrix(m, n)
        _ffi.check(self.lib.mspde_zgemm_tn(_ffi.current_stream(), int(conjA), m, n, k, c_double(alpha), _ffi.ptr(A),
                                           A.stride(0), _ffi.ptr(B), B.stride(0), c_double(beta), _ffi.ptr(C), C.stride(0),
                                           1 if upper else 0, None, 1), "zgemm_tn")
        return C

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

    def phi(self, L, parts: bool = False, fast: bool = False, naive: bool = False):
        """Phi(L) of pCN._log_ratio; synchronises to return Python floats.  fast=True: determinant-lemma form;
        naive=True: the reference's naive branch (644-657, no stabiliser)."""
        self._sync_stream()
        fn = self.lib.mspde_phi_naive if naive else (self.lib.mspde_phi_fast if fast else self.lib.mspde_phi)
        _ffi.check(fn(self._h, _ffi.ptr(L), L.stride(0), _ffi.ptr(self.scal)), "phi")
        self.check_info("phi (Cholesky)")
        v = self.scal[:3].cpu().numpy()
        return (float(v[0]), float(v[1]), float(v[2])) if parts else float(v[0])

    def solve(self, L, rhs, want_logdet: bool = False, method: str = "lu"):
        """x = L^-1 rhs (cp.linalg.solve call sites); optionally also log|det L| (Lmatrix_2D.logDet).
        method: "lu" (partial pivoting, default) or "qr" (Householder)."""
        self._sync_stream()
        x = torch.empty(self.N, dtype=CDTYPE, device=self.device)
        ld = self.scal[8:9]
        fn = self.lib.mspde_solve if method == "lu" else self.lib.mspde_solve_qr
        _ffi.check(fn(self._h, _ffi.ptr(L), L.stride(0), _ffi.ptr(rhs), _ffi.ptr(x), _ffi.ptr(ld)), "solve")
        if method == "lu" and want_logdet:
            self.check_info("solve (singular matrix)")
        return (x, float(ld.item())) if want_logdet else x

    def lstsq(self, a, b):
        """util.lstsq(a, b)[0] for a (rows x cols) with unit column stride."""
        self._sync_stream()
        rows, cols = a.shape
        if a.stride(1) != 1:
            a = a.contiguous()
        x = torch.empty(cols, dtype=CDTYPE, device=self.device)
        _ffi.check(self.lib.mspde_lstsq(self._h, _ffi.ptr(a), a.stride(0), rows, cols, _ffi.ptr(b), _ffi.ptr(x)), "lstsq")
        return x

    def gibbs_lstsq(self, L, rhs):
        self._sync_stream()
        v = torch.empty(self.N, dtype=CDTYPE, device=self.device)
        _ffi.check(self.lib.mspde_gibbs_lstsq(self._h, _ffi.ptr(L), L.stride(0), _ffi.ptr(rhs), _ffi.ptr(v)), "gibbs_lstsq")
        return v

    def pcn_step(self, chain: "ChainBuffers", w_half, log_u: float, w_last, e, flags: int = 0, record=None, out=None):
        """One pCN.one_step on the device.  Returns the 4-double device tensor {accepted, logRatio, Phi_cur, Phi_new}."""
        self._sync_stream()
        K = self.n_layers
        out = torch.empty(4, dtype=torch.float64, device=self.device) if out is None else out
        # the cached Phi(L_cur) is only reusable for the same inputs and the same Phi formulation
        tag = (self.generation, int(flags) & (STEP_FAST_PHI | STEP_NAIVE_PHI))
        if chain.phi_cache_tag != tag:
            chain.phi_cur_valid = 0
        cs = chain.as_struct()
        wh = _ptr_array(list(w_half))
        nz = _StepNoise(ctypes.cast(wh, ctypes.POINTER(ctypes.c_void_p)), w_last.data_ptr(), e.data_ptr(), float(log_u))
        _ffi.check(self.lib.mspde_pcn_step(self._h, ctypes.byref(cs), ctypes.byref(nz), int(flags), _ffi.ptr(out),
                                           _ffi.ptr(record)), "pcn_step")
        chain.phi_cur_valid, chain.phi_cache_tag = 1, tag
        return out

    def posterior_stats_update(self, u_sym, count: int, stats: dict):
        """Welford update of stats = dict(mean_u, m2_u, mean_ell, m2_ell) (m x m float64 device tensors)."""
        self._sync_stream()
        _ffi.check(self.lib.mspde_posterior_stats_update(self._h, _ffi.ptr(u_sym), ctypes.c_longlong(int(count)),
                                                         _ffi.ptr(stats["mean_u"]), _ffi.ptr(stats["m2_u"]),
                                                         _ffi.ptr(stats["mean_ell"]), _ffi.ptr(stats["m2_ell"])), "stats")

    def new_stats(self):
        z = lambda: torch.zeros((self.m, self.m), dtype=torch.float64, device=self.device)
        return dict(mean_u=z(), m2_u=z(), mean_ell=z(), m2_ell=z(), count=0)

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


class ChainBuffers:
    """Device-resident state of one chain: per layer noise_cur/new, u_cur/new (N complex)
    and, for layers k >= 1, L_cur / L_latest (N x ldN)."""

    def __init__(self, ctx: MspdeContext, sqrt_betas, beta: float = 1.0):
        K, N = ctx.n_layers, ctx.N
        self.ctx = ctx
        z = lambda: torch.zeros(N, dtype=CDTYPE, device=ctx.device)
        self.noise_cur = [z() for _ in range(K)]
        self.noise_new = [z() for _ in range(K)]
        self.u_cur = [z() for _ in range(K)]
        self.u_new = [z() for _ in range(K)]
        self.L_cur = [None] + [ctx.new_matrix(N, N) for _ in range(K - 1)]
        self.L_latest = [None] + [ctx.new_matrix(N, N) for _ in range(K - 1)]
        self.sqrt_betas = [float(s) for s in sqrt_betas]
        self._sb = (c_double * K)(*self.sqrt_betas)
        self.phi_cur_valid = 0
        self.phi_cache = torch.zeros(4, dtype=torch.float64, device=ctx.device)   # per-chain {Phi, quad, logdet} of L_cur
        self.phi_cache_tag = None
        self.set_step(beta)

    def invalidate_phi_cache(self):
        """Call after writing L_cur from outside pcn_step (e.g. the sqrt-beta move promotes a new current_L)."""
        self.phi_cur_valid = 0

    def set_step(self, beta: float):
        self.beta = float(beta)
        self.betaZ = float(np.sqrt(max(0.0, 1.0 - self.beta ** 2)))

    def as_struct(self) -> _Chain:
        K = self.ctx.n_layers
        self._arrs = [_ptr_array(x) for x in (self.noise_cur, self.noise_new, self.u_cur, self.u_new, self.L_cur,
                                              self.L_latest)]
        casted = [ctypes.cast(a, ctypes.POINTER(ctypes.c_void_p)) for a in self._arrs]
        return _Chain(K, self.beta, self.betaZ, ctypes.cast(self._sb, ctypes.POINTER(c_double)), *casted,
                      self.L_cur[1].stride(0), self.phi_cur_valid, self.phi_cache.data_ptr())
