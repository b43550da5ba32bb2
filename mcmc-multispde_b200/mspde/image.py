"""Host-side mirror of /root/reference/mcmc/image_cupy.py for the pCN hot path.

Same class names, constructor arguments, attributes and error behaviour as the reference's
``FourierAnalysis_2D``, ``Lmatrix_2D``, ``RandomGenerator_2D``, ``Sinogram``, ``pCN``, ``Layer`` and ``Simulation``
(reference lines cited per method), but every operation on the per-iteration path is a call into libmspde.so
(hand-written sm_100a CUDA) through ``mspde.core``.  State lives in FP64/complex128 device buffers; there is no
CPU fallback.  Setup-only pieces (tomography matrix, phantom, sinogram) are numpy on the host.
"""
from __future__ import annotations

import ctypes
import math
import os
import pathlib
import time

import numpy as np
import torch

from . import util
from .core import (CDTYPE, ChainBuffers, MspdeContext, STEP_CACHE_PHI_CUR, STEP_FAST_PHI, STEP_NAIVE_PHI,
                   STEP_SKIP_GIBBS)

ORDER = util.ORDER


# --------------------------------------------------------------------------------------------------------------
class FourierAnalysis_2D:
    """image_cupy.py:33-157.  Index tables live on the host (numpy int32/float64) and, where the kernels need
    them, on the device.  The dense int16 ``_Umask`` (83-85) is not materialised (only used by dead code)."""

    def __init__(self, basis_number, extended_basis_number, t_start=0, t_end=1, mempool=None):
        n = basis_number
        self.basis_number = n
        self.extended_basis_number = extended_basis_number
        self.basis_number_2D = (2 * n - 1) * n
        self.basis_number_2D_ravel = 2 * n * n - 2 * n + 1
        self.basis_number_2D_sym = (2 * n - 1) * (2 * n - 1)
        self.extended_basis_number_2D = (2 * extended_basis_number - 1) * extended_basis_number
        self.extended_basis_number_2D_sym = (2 * extended_basis_number - 1) ** 2
        self.t_end, self.t_start = t_end, t_start
        self.verbose = False
        il, N = 2 * n - 1, self.basis_number_2D_sym
        from . import _ffi
        ix = np.empty((il, il), np.int32); iy = np.empty((il, il), np.int32)
        km = np.empty((2, N), np.int32); D = np.empty(N, np.float64)
        _ffi.check(_ffi.lib().mspde_index_maps(n, _ffi.ptr(ix), _ffi.ptr(iy), _ffi.ptr(km), _ffi.ptr(D)), "index_maps")
        self.ix, self.iy, self.Kmatrix, self.D_diag = ix, iy, km, D     # 52-57, 63, 73-74

    # FFT conventions of 108-129; used off the hot path (initialisation / analysis) -- the per-iteration field
    # synthesis is mspde_field_coeffs inside the library.
    def rfft2(self, z):
        m = z.shape[0]
        t = torch.fft.fft2(torch.as_tensor(z).to(torch.complex128))
        t = torch.fft.fftshift(t, dim=0)
        n = self.basis_number
        return t[m // 2 - (n - 1):m // 2 + n, :n] / (2 * self.extended_basis_number - 1)

    def irfft2(self, uHalf2D):
        u2D = util.extend2D(util.symmetrize_2D(torch.as_tensor(uHalf2D)), self.extended_basis_number)
        u2D = torch.fft.ifftshift(u2D)
        return torch.fft.ifft2(u2D).real * (2 * self.extended_basis_number - 1)


# --------------------------------------------------------------------------------------------------------------
class Lmatrix_2D:
    """image_cupy.py:160-284.  ``construct_from`` writes into one of two persistent device buffers and returns it
    (the reference returns a fresh array and aliases ``current_L``/``latest_computed_L``; here promotion is a
    buffer swap, same observable semantics)."""

    def __init__(self, f, sqrt_beta, ctx: MspdeContext | None = None):
        self.fourier = f
        self.sqrt_beta = sqrt_beta
        self.ctx = ctx
        self.current_L = None
        self.latest_computed_L = self.current_L

    def _target(self):
        if self.latest_computed_L is None or self.latest_computed_L is self.current_L:
            return None          # allocate a fresh buffer (keeps current_L intact, like the reference)
        return self.latest_computed_L

    def construct_from(self, u_sym, hybrid=False):                       # 231-233
        L = self.ctx.construct_L(u_sym, float(self.sqrt_beta), out=self._target())
        self.latest_computed_L = L
        return L

    def construct_from_with_sqrt_beta(self, u_sym, sqrt_beta, hybrid=False):   # 235-261
        L = self.ctx.construct_L(u_sym, float(sqrt_beta), out=self._target())
        self.latest_computed_L = L
        return L

    def logDet(self, new):                                               # 268-276
        L = self.latest_computed_L if new else self.current_L
        rhs = torch.zeros(L.shape[0], dtype=CDTYPE, device=L.device)
        return self.ctx.solve(L, rhs, want_logdet=True)[1]

    def set_current_L_to_latest(self):                                   # 279-280
        self.current_L = self.latest_computed_L

    def is_current_L_equals_to_the_latest(self):                         # 283-284
        return bool(torch.all(self.current_L == self.latest_computed_L))


# --------------------------------------------------------------------------------------------------------------
class RandomGenerator_2D:
    """image_cupy.py:287-314.  Draws FP64 normals on the device (torch Philox) unless a noise source is injected
    at the pCN level."""

    def __init__(self, basis_number, device=None, seed=None):
        self.basis_number = basis_number
        self.basis_number_2D_ravel = 2 * basis_number * basis_number - 2 * basis_number + 1
        self.sqrt2 = util.SQRT2
        self.device = device if device is not None else torch.device("cuda")
        self.gen = torch.Generator(device=self.device)
        if seed is not None:
            self.gen.manual_seed(int(seed))

    def construct_w_Half(self):                                          # 297-298 -> util_cupy.py:34-39
        nr = self.basis_number_2D_ravel
        z = torch.randn(2, nr, dtype=torch.float64, device=self.device, generator=self.gen)
        w = torch.complex(z[0], z[1])
        w[0] = 2.0 * w[0].real
        return w / float(self.sqrt2)

    def construct_w(self):                                               # 300-303
        return self.symmetrize(self.construct_w_Half())

    def symmetrize(self, w_half):                                        # 305-308
        return util.symmetrize(w_half)

    def symmetrize_2D(self, uHalf2D):                                    # 313-314
        return util.symmetrize_2D(uHalf2D)


# --------------------------------------------------------------------------------------------------------------
def shepp_logan(m: int) -> np.ndarray:
    """Analytic modified Shepp-Logan phantom on an m x m grid (used when the PNG phantom is unavailable)."""
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
    return img[::-1].copy()


class TwoDMeasurement:
    """image_cupy.py:316-347 (phantom loading).  skimage is not available: the PNG (if present) is read with PIL
    and resized bilinearly; otherwise the analytic Shepp-Logan phantom is used."""

    def __init__(self, file_name, target_size=128, stdev=0.1, relative_location="", rng=None):
        if target_size > 512:
            raise Exception("Target image dimension are too large (" + str(target_size) + " x " + str(target_size) + ")")
        self.dim = target_size
        self.stdev = stdev
        self.rng = rng if rng is not None else np.random.Generator(np.random.PCG64(1))
        path = pathlib.Path.cwd() / relative_location / file_name
        img = None
        if path.exists():
            try:
                from PIL import Image
                im = Image.open(path).convert("L")
                if im.size[0] != im.size[1]:
                    raise Exception("Image is not square")
                img = np.asarray(im.resize((self.dim, self.dim), Image.BILINEAR), dtype=np.float64) / 255.0
            except ImportError:
                img = None
        if img is None:
            img = shepp_logan(self.dim)
        self.target_image = img
        self.corrupted_image = img + self.stdev * self.rng.standard_normal(img.shape)
        self.v = self.target_image.ravel(ORDER) / self.stdev
        self.y = self.corrupted_image.ravel(ORDER) / self.stdev
        self.num_sample = self.y.size


class Sinogram(TwoDMeasurement):
    """image_cupy.py:349-404.  ``get_measurement_matrix`` is the closed form of the ``_calculate_H_Tomography``
    kernel (mcmc/util_cupy.py:445-470) plus the reference's post-fixes (377-385), in FP64 numpy (setup only).
    The sinogram is synthesised through H itself (skimage.radon is unavailable): ``Re(H0 v_true) + noise``."""

    def __init__(self, file_name, target_size=128, n_theta=50, stdev=0.1, relative_location="", rng=None):
        super().__init__(file_name, target_size, stdev, relative_location, rng)
        self.use_skimage = False
        self.n_theta = n_theta
        self.theta = np.linspace(0.0, 180.0, self.n_theta, endpoint=False)          # 354
        self.n_r = self.dim
        self.r = np.linspace(-0.5, 0.5, self.n_r, endpoint=False)[::-1]            # 358-361
        self.pure_sinogram = None
        self.sinogram = None
        self.num_sample = self.n_r * self.n_theta

    def get_measurement_matrix(self, ix, iy):
        """ix, iy: 'F'-ravelled index vectors (length N).  Returns the un-normalised M x N complex128 H."""
        n_r, n_theta = self.n_r, self.n_theta
        shifted = self.theta + self.theta[1]                                        # 365
        tg, rg = np.meshgrid(shifted * np.float64(util.PI) / 180.0, self.r)         # 366
        rv, tv = rg.ravel(ORDER), tg.ravel(ORDER)
        kx, ky = ix.astype(np.float64), iy.astype(np.float64)
        th = tv[::-1] + 0.5 * math.pi                                               # util_cupy.py:457
        H = np.empty((rv.size, kx.size), dtype=np.complex128)
        blk = 4096
        for a in range(0, rv.size, blk):                                            # row blocks keep temporaries small
            s, c = np.sin(th[a:a + blk])[:, None], np.cos(th[a:a + blk])[:, None]
            r = rv[a:a + blk][:, None]
            ku = kx[None, :] * c + ky[None, :] * s
            kv = -kx[None, :] * s + ky[None, :] * c
            ell = np.sqrt(0.25 - r * r)
            phase = np.exp(1j * math.pi * ((kx + ky)[None, :] - 2.0 * kv * r))
            with np.errstate(divide="ignore", invalid="ignore"):
                val = np.sin(2.0 * math.pi * ku * ell) / (math.pi * ku)
            H[a:a + blk] = np.where(kv * kv > 0.0, phase * val, phase * (2.0 * ell))
        H *= n_r / (2 * (n_r // 2))                                                 # 377-378
        for b in range(n_theta // 4):                                               # 381-385 ("temporary fixing")
            H[b * n_r:(b + 1) * n_r, :] = np.roll(H[b * n_r:(b + 1) * n_r, :], 1, axis=0)
        return H

    def get_measurement_matrix_device(self, ix, iy, device, apply_fixes=True):
        """Same matrix built on the GPU by mspde_tomography_H: returns an M x N complex128 torch tensor."""
        import ctypes
        from . import _ffi
        shifted = self.theta + self.theta[1]
        tg, rg = np.meshgrid(shifted * np.float64(util.PI) / 180.0, self.r)
        rv = torch.as_tensor(np.ascontiguousarray(rg.ravel(ORDER))).to(device)
        tv = torch.as_tensor(np.ascontiguousarray(tg.ravel(ORDER))).to(device)
        ixd = torch.as_tensor(np.ascontiguousarray(ix, dtype=np.int32)).to(device)
        iyd = torch.as_tensor(np.ascontiguousarray(iy, dtype=np.int32)).to(device)
        M, N = rv.numel(), ixd.numel()
        ld = (N + 7) // 8 * 8
        H = torch.zeros((M, ld), dtype=CDTYPE, device=device)[:, :N]
        with torch.cuda.device(device):
            _ffi.check(_ffi.lib().mspde_tomography_H(_ffi.current_stream(), M, N, _ffi.ptr(rv), _ffi.ptr(tv), _ffi.ptr(ixd),
                                                     _ffi.ptr(iyd), self.n_r, self.n_theta, int(bool(apply_fixes)), _ffi.ptr(H),
                                                     H.stride(0)), "tomography_H")
        return H

    def synthesize(self, H0, fourier: FourierAnalysis_2D):
        """Synthetic stand-in for set_theta (398-404): sinogram = Re(H0 v_true) + stdev * N(0,1)."""
        v_half = fourier.rfft2(torch.as_tensor(self.target_image)).numpy()
        v_true = util.symmetrize_2D(v_half).ravel(ORDER)
        if isinstance(H0, torch.Tensor):
            pure = (H0 @ torch.as_tensor(v_true).to(H0.device)).real.cpu().numpy()
        else:
            pure = (H0 @ v_true).real
        self.pure_sinogram = pure.reshape(self.n_r, self.n_theta, order=ORDER)
        self.sinogram = self.pure_sinogram + self.stdev * self.rng.standard_normal(self.pure_sinogram.shape)
        self.y = self.sinogram.ravel(ORDER) / self.stdev                            # 402
        self.v = self.pure_sinogram.ravel(ORDER) / self.stdev
        self.num_sample = self.y.size
        self.v_true = v_true


# --------------------------------------------------------------------------------------------------------------
class pCN:
    """image_cupy.py:411-710.  ``one_step`` runs the whole iteration on the device through mspde_pcn_step."""

    def __init__(self, n_layers, rg, measurement, f, beta=1, variant="dunlop", verbose=True, hybrid_mode=False,
                 mempool=None, device=0, H0=None):
        if hybrid_mode:
            raise NotImplementedError("hybrid GPU/CPU mode does not exist in the B200 build (180 GB HBM; no CPU path)")
        self.n_layers = n_layers
        self.beta = float(beta)
        self.betaZ = math.sqrt(max(0.0, 1 - self.beta ** 2))
        self.random_gen = rg
        self.measurement = measurement
        self.fourier = f
        self.variant = variant
        self.meas_var = self.measurement.stdev ** 2
        self.verbose = verbose
        self.hybrid_mode = False
        self.epsilon = 0
        self.cholesky_stabilizer = 0.0
        N = f.basis_number_2D_sym
        dev = torch.device("cuda", device)
        if H0 is None:                                                              # create_H_matrix, 496-508 (on the device)
            H0 = measurement.get_measurement_matrix_device(f.ix.ravel(ORDER), f.iy.ravel(ORDER), dev)
        if getattr(measurement, "sinogram", 1) is None:
            measurement.synthesize(H0, f)
        M = H0.shape[0]
        self.ctx = MspdeContext(f.basis_number, f.extended_basis_number, M, n_layers, device=device)
        ctx = self.ctx
        if isinstance(H0, torch.Tensor) and H0.is_cuda and H0.stride(0) % 8 == 0:
            H0d = H0
        else:
            H0d = ctx.new_matrix(M, N); H0d.copy_(ctx.to_dev(H0, CDTYPE))
        t2 = ctx.gemm_tn(H0d, H0d, conjA=True)                                      # H^H H on the DMMA pipe (505)
        HtH0 = 0.5 * (t2 + t2.conj().transpose(0, 1))                               # 506
        self.H = H0d / self.measurement.stdev                                       # 437
        self.H_t_H = HtH0 / self.meas_var                                           # 442
        del H0d, t2, HtH0
        self.y = ctx.to_dev(self.measurement.y, torch.float64)                      # 448
        self.yBar = torch.cat((self.y, torch.zeros(2 * f.basis_number_2D_ravel - 1, dtype=torch.float64,
                                                    device=ctx.device)))            # 449
        self.stdev_sym = None
        self.aggresiveness = 0.2
        self.target_acceptance_rate = 0.234
        self.beta_feedback_gain = 2.1
        self.record_skip = 1
        self.record_count = 0
        self.max_record_history = 1000000
        self.use_beta_adaptation = False
        self.pcn_step_sqrtBetas = 5e-1                                              # 467
        self.pcn_step_sqrtBetas_Z = math.sqrt(1 - self.pcn_step_sqrtBetas)          # 468 (sic: no square, as in the reference)
        self.stdev_sqrtBetas = np.ones(n_layers)                                    # 469
        self.use_naive_logRatio_implementation = False
        self.cache_phi_current = False      # results-preserving shortcut (SURVEY 8d note i); off = reference-faithful
        self.skip_gibbs = False
        self.accumulate_stats = False       # on-device Welford mean/variance of every layer's field
        self.fast_phi = False               # determinant-lemma Phi (SURVEY 8d note ii); off = reference-faithful Woodbury
        self.noise_source = None            # object with draw_iteration() -> dict (injected noise), else device RNG
        self._inputs_set = False
        self._uploaded = False
        self._chain = None
        self._chain_layers = None
        self._out = torch.zeros(4, dtype=torch.float64, device=ctx.device)
        self._out_host = torch.zeros(4, dtype=torch.float64).pin_memory()
        self.last_step = None

    @property
    def H_conj_T(self):                                                             # 441
        return self.ctx.Hct

    # -- stabiliser (473-478) ------------------------------------------------------------------------------
    def reset_chol_stabilizer(self):
        self.cholesky_stabilizer = float(self.epsilon * torch.linalg.norm(self.H_t_H).item())
        self._inputs_set = False

    def set_chol_epsilon(self, cholEpsilon):
        self.epsilon = cholEpsilon
        self.reset_chol_stabilizer()

    def _ensure_inputs(self, stdev_sym=None):
        if stdev_sym is not None:
            self.stdev_sym = stdev_sym
        if not self._inputs_set:
            sd = self.stdev_sym if self.stdev_sym is not None else np.ones(self.fourier.basis_number_2D_sym)
            if not self._uploaded:
                self.ctx.set_inputs(self.H, self.H_t_H, self.y, self.fourier.D_diag, sd, self.cholesky_stabilizer)
                self.H, self.H_t_H, self.y = self.ctx.H, self.ctx.HtH, self.ctx.y
                self._uploaded = True
            else:
                self.ctx.update_scalars(stdev_sym=sd, delta=self.cholesky_stabilizer)
            self._inputs_set = True

    # -- step adaptation (528-540) -------------------------------------------------------------------------
    def adapt_step(self, current_acceptance_rate):
        self.set_step(self.beta * math.exp(self.beta_feedback_gain * (current_acceptance_rate - self.target_acceptance_rate)))

    def more_aggresive(self):
        self.set_step(min((1 + self.aggresiveness) * self.beta, 1))

    def less_aggresive(self):
        self.set_step(min((1 - self.aggresiveness) * self.beta, 1e-10))

    def set_step(self, newBeta):
        if 1e-7 < newBeta < 1:
            self.beta = float(newBeta)
            self.betaZ = math.sqrt(1 - self.beta ** 2)

    def set_step_sqrtBeta(self, newStep):                                           # 542-544
        self.pcn_step_sqrtBetas = min(1., newStep)
        self.pcn_step_sqrtBetas_Z = math.sqrt(1 - self.pcn_step_sqrtBetas ** 2)

    def adapt_step_sqrtBeta(self, current_acceptance_rate):                         # 546-548
        self.set_step_sqrtBeta(self.pcn_step_sqrtBetas * math.exp(self.beta_feedback_gain * (current_acceptance_rate
                                                                                             - self.target_acceptance_rate)))

    # -- Phi (618-662) -------------------------------------------------------------------------------------
    def _log_ratio(self, L):
        self._ensure_inputs()
        return self.ctx.phi(L, fast=self.fast_phi, naive=self.use_naive_logRatio_implementation)    # 633 / 644

    # -- one iteration (553-617) ---------------------------------------------------------------------------
    def _bind(self, Layers):
        if self._chain is None or self._chain_layers is not Layers:
            cb = ChainBuffers.__new__(ChainBuffers)
            cb.ctx = self.ctx
            cb.noise_cur = [l._noise_cur for l in Layers]
            cb.noise_new = [l._noise_new for l in Layers]
            cb.u_cur = [l._u_cur for l in Layers]
            cb.u_new = [l._u_new for l in Layers]
            cb.phi_cur_valid, cb.phi_cache_tag = 0, None
            cb.phi_cache = torch.zeros(4, dtype=torch.float64, device=self.ctx.device)
            self._chain, self._chain_layers = cb, Layers
        cb = self._chain
        # The assembly divides by Lmatrix_2D.sqrt_beta (image_cupy.py:179), which is fixed at construction (161-163): an
        # accepted sqrt-beta move changes Layer.sqrt_beta (704) but NOT what construct_from uses.  Reproduced as is.
        cb.sqrt_betas = [float(l.LMat.sqrt_beta) for l in Layers]
        cb._sb = (ctypes.c_double * len(Layers))(*cb.sqrt_betas)
        # L buffers can be re-pointed by set_current_L_to_latest / construct_from: refresh every step
        for l in Layers[1:]:
            l._ensure_two_L_buffers()
        cb.L_cur = [None] + [l.LMat.current_L for l in Layers[1:]]
        cb.L_latest = [None] + [l.LMat.latest_computed_L for l in Layers[1:]]
        cb.beta, cb.betaZ = self.beta, self.betaZ
        return cb

    def _draw_noise(self):
        """RNG call order of the reference (558, 573, 583, 585); a noise source may add a "move" entry (dict z, e, log_u:
        the draws of one_step_for_sqrtBetas, 668, 684, 699)."""
        if self.noise_source is not None:
            nz = self.noise_source.draw_iteration()
            dev = self.ctx.to_dev
            return [dev(w) for w in nz["w_half"]], float(nz["log_u"]), dev(nz["w_last"]), dev(nz["e"]), nz.get("move")
        rg = self.random_gen
        w = [rg.construct_w_Half() for _ in range(self.n_layers - 1)]
        log_u = math.log(float(torch.rand(1, dtype=torch.float64, device=self.ctx.device, generator=rg.gen).item()))
        w_last = rg.construct_w_Half()
        e = torch.randn(self.measurement.num_sample, dtype=torch.float64, device=self.ctx.device, generator=rg.gen)
        return w, log_u, w_last, e, None

    def one_step(self, Layers, noise=None):
        """Returns (accepted, accepted_SqrtBeta) like the reference; raises numpy.linalg.LinAlgError when a
        Cholesky factorisation fails (the only error Simulation.run catches, 969).
        noise: (w_half list, log_u, w_last, e[, move]) device tensors to inject, else the noise source / device RNG."""
        self._ensure_inputs(Layers[0].stdev_sym_host)
        cb = self._bind(Layers)
        nz = tuple(noise) if noise is not None else self._draw_noise()
        w_half, log_u, w_last, e = nz[:4]
        move_noise = nz[4] if len(nz) > 4 else None
        flags = ((STEP_CACHE_PHI_CUR if self.cache_phi_current else 0) | (STEP_SKIP_GIBBS if self.skip_gibbs else 0)
                 | (STEP_FAST_PHI if self.fast_phi else 0) | (STEP_NAIVE_PHI if self.use_naive_logRatio_implementation else 0))
        do_record = (self.record_count % self.record_skip) == 0
        rec = self._record_buffer(Layers) if do_record else None
        self.ctx.pcn_step(cb, w_half, log_u, w_last, e, flags=flags, record=rec, out=self._out)
        self._out_host.copy_(self._out, non_blocking=True)
        if do_record:
            self._rec_host.copy_(rec, non_blocking=True)
        self.ctx.check_info("pCN.one_step (Cholesky)")          # synchronises the stream
        accepted = int(self._out_host[0].item())
        self.last_step = dict(accepted=accepted, log_ratio=float(self._out_host[1]), phi_cur=float(self._out_host[2]),
                              phi_new=float(self._out_host[3]), log_u=log_u)
        # 604-607: the hyper-move runs BEFORE the record (609-613); it changes sqrt_beta / current_L / stdev_sym but not the
        # current samples, so the half-samples copied out by the fused step are still what record_sample must store.
        accepted_SqrtBeta = self.one_step_for_sqrtBetas(Layers, noise=move_noise) if self.use_beta_adaptation else 0
        if do_record:
            for k, l in enumerate(Layers):
                l.record_sample(self._rec_host[k].numpy())
                l.record_sqrt_beta()
                if self.accumulate_stats:
                    if l.stats is None:
                        l.stats = self.ctx.new_stats()
                    l.stats["count"] += 1
                    self.ctx.posterior_stats_update(l._u_cur, l.stats["count"], l.stats)
        self.record_count += 1
        return accepted, accepted_SqrtBeta

    # -- optional hyper-parameter move (666-710) ------------------------------------------------------------
    def one_step_for_sqrtBetas(self, Layers, noise=None):
        """pCN.one_step_for_sqrtBetas with the reference's semantics (see the oracle restatement for the quirks).
        Composed from the C-ABI pieces (construct_L, solve, gibbs_lstsq, phi); off by default like the reference
        (use_beta_adaptation = False, image_cupy.py:470).  noise: dict(z, e, log_u) to inject, else device RNG."""
        ctx, K = self.ctx, self.n_layers
        self._ensure_inputs(Layers[0].stdev_sym_host)
        if noise is None:
            g = self.random_gen.gen
            noise = dict(z=torch.randn(K, dtype=torch.float64, device=ctx.device, generator=g).cpu().numpy(),
                         e=torch.randn(ctx.M, dtype=torch.float64, device=ctx.device, generator=g),
                         log_u=math.log(float(torch.rand(1, dtype=torch.float64, device=ctx.device, generator=g).item())))
        s = self.pcn_step_sqrtBetas
        z = np.asarray(self.stdev_sqrtBetas) * np.asarray(noise["z"])                                  # 668
        e = ctx.to_dev(noise["e"], torch.float64)
        prop = np.zeros(K)
        stdev_sym_temp = None
        for i in range(K):
            prop[i] = math.exp(math.sqrt(1 - s ** 2) * math.log(Layers[i].sqrt_beta) + s * z[i])       # 673-674
            if i == 0:
                stdev_sym_temp = (prop[0] / Layers[0].sqrt_beta) * Layers[0].stdev_sym_host            # 676
                Layers[0]._u_new.copy_(ctx.to_dev(stdev_sym_temp, torch.float64) * Layers[0]._noise_cur)   # 677
            else:
                Layers[i]._ensure_two_L_buffers()
                Layers[i].LMat.construct_from_with_sqrt_beta(Layers[i - 1]._u_new, prop[i])            # 679
                if i < K - 1:
                    Layers[i]._u_new.copy_(ctx.solve(Layers[i].LMat.latest_computed_L, Layers[i]._noise_cur))   # 681
                else:
                    wBar = torch.cat((e.to(CDTYPE), Layers[-1]._noise_cur))                            # 683-685
                    rhs = self.yBar.to(CDTYPE) - wBar
                    Layers[-1]._u_new.copy_(ctx.gibbs_lstsq(Layers[-1].LMat.latest_computed_L, rhs))   # 686-688
        phi_cur = self._log_ratio(Layers[-1].LMat.current_L)                                           # 692-693
        L = Layers[-1].LMat.construct_from(Layers[-2]._u_new)                                          # 695 (LMat.sqrt_beta)
        phi_new = self._log_ratio(L)                                                                   # 696
        log_ratio = phi_cur - phi_new
        accepted = int(log_ratio > float(noise["log_u"]))                                              # 699
        if accepted:
            for i in range(K):                                                                         # 703-708
                Layers[i].sqrt_beta = float(prop[i])
                if not Layers[i].is_stationary:
                    Layers[i].LMat.set_current_L_to_latest()
                    Layers[i].LMat.latest_computed_L = Layers[i]._clone_padded(Layers[i].LMat.current_L)
                else:
                    Layers[i].stdev_sym = stdev_sym_temp
            if self._chain is not None:
                self._chain.invalidate_phi_cache()      # current_L of the last layer was replaced
        self.last_sqrt_beta_step = dict(accepted=accepted, log_ratio=log_ratio, phi_cur=phi_cur, phi_new=phi_new, prop=prop)
        return accepted

    def _record_buffer(self, Layers):
        if getattr(self, "_rec", None) is None:
            nr = self.fourier.basis_number_2D_ravel
            self._rec = torch.zeros((len(Layers), nr), dtype=CDTYPE, device=self.ctx.device)
            self._rec_host = torch.zeros((len(Layers), nr), dtype=CDTYPE).pin_memory()
        return self._rec


# --------------------------------------------------------------------------------------------------------------
class Layer:
    """image_cupy.py:717-807.  Attributes are persistent device buffers; assigning to them copies into place."""

    def __init__(self, is_stationary, sqrt_beta, order_number, n_samples, pcn, init_sample_sym):
        self.is_stationary = is_stationary
        self.sqrt_beta = float(sqrt_beta)
        self.order_number = order_number
        self.n_samples = n_samples
        self.pcn = pcn
        f, ctx = pcn.fourier, pcn.ctx
        nr, N = f.basis_number_2D_ravel, f.basis_number_2D_sym
        self._nr = nr
        z = lambda: torch.zeros(N, dtype=CDTYPE, device=ctx.device)
        self._noise_cur, self._noise_new, self._u_cur, self._u_new = z(), z(), z(), z()
        self.stdev_sym_host = np.ones(N)
        self.samples_history = np.empty((self.n_samples, nr), dtype=np.complex128)          # 729 (complex64 there)
        self.sqrt_beta_history = np.empty(self.n_samples, dtype=np.float64)
        self.sqrt_beta_history[0] = self.sqrt_beta
        self.LMat = Lmatrix_2D(f, self.sqrt_beta, ctx)
        self._noise_cur.copy_(pcn.random_gen.construct_w())                                 # 735
        self._noise_new.copy_(self._noise_cur)                                              # 736
        init_sample_sym = torch.as_tensor(init_sample_sym).to(ctx.device).to(CDTYPE)
        if self.is_stationary:                                                              # 739-749
            self._u_new.copy_(init_sample_sym)
            self._u_cur.copy_(init_sample_sym)
            self.new_log_L_det = 0.0
        else:                                                                               # 751-763
            pcn._ensure_inputs()
            self.LMat.construct_from(init_sample_sym)
            self.LMat.set_current_L_to_latest()
            x, logdet = ctx.solve(self.LMat.current_L, pcn.random_gen.construct_w(), want_logdet=True)   # 755, 758
            self._u_new.copy_(x)
            self._u_cur.copy_(x)
            self.new_log_L_det = logdet
        self.current_log_L_det = self.new_log_L_det
        self.i_record = 0
        self.stats = None

    def field_statistics(self):
        """finalizeWelford (mcmc/util_cupy.py:181-187): dict of mean / variance / sample variance of u(x) and exp(u(x))."""
        s = self.stats
        c = s["count"]
        return dict(count=c, u_mean=s["mean_u"], u_var=s["m2_u"] / c, u_s_var=s["m2_u"] / max(c - 1, 1),
                    ell_mean=s["mean_ell"], ell_var=s["m2_ell"] / c, ell_s_var=s["m2_ell"] / max(c - 1, 1))

    def _ensure_two_L_buffers(self):
        lm = self.LMat
        if lm.latest_computed_L is lm.current_L:
            lm.latest_computed_L = lm.current_L.clone() if lm.current_L.stride(0) == lm.current_L.shape[1] else \
                self._clone_padded(lm.current_L)

    def _clone_padded(self, L):
        out = self.pcn.ctx.new_matrix(L.shape[0], L.shape[1])
        out.copy_(L)
        return out

    # buffers exposed under the reference's attribute names
    def _mk(name):
        def get(self):
            return getattr(self, name)

        def set_(self, v):
            getattr(self, name).copy_(torch.as_tensor(v).to(getattr(self, name).device))
        return property(get, set_)

    current_noise_sample = _mk("_noise_cur")
    new_noise_sample = _mk("_noise_new")
    current_sample_sym = _mk("_u_cur")
    new_sample_sym = _mk("_u_new")
    del _mk

    @property
    def current_sample(self):
        return self._u_cur[self._nr - 1:]

    @property
    def new_sample(self):
        return self._u_new[self._nr - 1:]

    @property
    def stdev_sym(self):
        return self.stdev_sym_host

    @stdev_sym.setter
    def stdev_sym(self, v):
        self.stdev_sym_host = np.asarray(v, dtype=np.float64)
        self.pcn._inputs_set = False

    def sample_non_centered(self):                                                          # 789-790
        w = self.pcn.random_gen.construct_w()
        self._noise_new.copy_(self.pcn.betaZ * self._noise_cur + self.pcn.beta * w)

    def record_sample(self, half_host=None):                                                # 792-795
        if self.i_record < self.samples_history.shape[0]:
            if half_host is None:
                half_host = self.current_sample.cpu().numpy()
            self.samples_history[self.i_record, :] = half_host
            self.i_record += 1

    def record_sqrt_beta(self):                                                             # 797-799
        if self.i_record < self.sqrt_beta_history.shape[0]:
            self.sqrt_beta_history[self.i_record] = self.sqrt_beta

    def update_current_sample(self):                                                        # 802-807
        self._u_cur.copy_(self._u_new)
        self.current_log_L_det = self.new_log_L_det
        self._noise_cur.copy_(self._noise_new)


# --------------------------------------------------------------------------------------------------------------
class Simulation:
    """image_cupy.py:810-1014, same constructor signature."""

    def __init__(self, n_layers, n_samples, n, n_extended, beta, kappa, sigma_0, sigma_v, sigma_scaling, meas_std,
                 evaluation_interval, printProgress, seed, burn_percentage, enable_step_adaptation, pcn_variant,
                 phantom_name, meas_type='tomo', n_theta=50, verbose=False, hybrid_GPU_CPU=False, device=0):
        if meas_type != 'tomo':
            raise NotImplementedError("only the tomography measurement is on the accelerated path")
        self.n_samples = n_samples
        self.evaluation_interval = evaluation_interval
        self.burn_percentage = burn_percentage
        self.random_seed = seed
        self.printProgress = printProgress
        self.n_layers = n_layers
        self.kappa, self.sigma_0, self.sigma_v, self.sigma_scaling = kappa, sigma_0, sigma_v, sigma_scaling
        self.enable_step_adaptation = enable_step_adaptation
        self.verbose = verbose
        self.hybrid = hybrid_GPU_CPU
        self.acceptancePercentage = 0.
        self.accepted_count = 0
        self.accepted_count_SqrtBeta = 0
        self.total_time = 0.0
        dev = torch.device("cuda", device)
        # 839-847
        self.d = 2
        self.nu = 2 - self.d / 2
        self.alpha = self.nu + self.d / 2
        self.t_start, self.t_end = -0.5, 0.5
        pi = np.float64(util.PI)
        self.beta_u = (sigma_0 ** 2) * (2 ** self.d * pi ** (self.d / 2) * math.gamma(self.alpha)) / math.gamma(self.nu)
        self.beta_v = self.beta_u * (sigma_v / sigma_0) ** 2
        self.sqrtBeta_v = np.float32(np.sqrt(self.beta_v))
        self.sqrtBeta_0 = np.float32(np.sqrt(self.beta_u))
        f = FourierAnalysis_2D(n, n_extended, self.t_start, self.t_end)
        self.fourier = f
        rg = RandomGenerator_2D(f.basis_number, device=dev, seed=seed)
        self.random_gen = rg
        # 869, 878-880 (float32 rounding of the diagonal prior and the DC aliasing quirk are reproduced)
        Lu_diag = ((f.D_diag * self.kappa ** (-self.nu) - self.kappa ** (2 - self.nu) * np.ones(f.basis_number_2D_sym, np.float32))
                   / np.float64(self.sqrtBeta_0)).astype(np.float32)
        uStdev_sym = (-1.0 / Lu_diag.astype(np.float64)).astype(np.float32).astype(np.float64)
        uStdev_sym[f.basis_number_2D_ravel - 1] /= 2
        self.measurement = Sinogram(phantom_name, target_size=2 * f.extended_basis_number - 1, n_theta=n_theta,
                                    stdev=meas_std, relative_location='phantom_images',
                                    rng=np.random.Generator(np.random.PCG64(seed)))
        self.pcn_variant = pcn_variant
        self.pcn = pCN(n_layers, rg, self.measurement, f, beta, self.pcn_variant, verbose=self.verbose, device=device)
        self.pcn.stdev_sym = uStdev_sym
        self.pcn.record_skip = max(1, self.n_samples // self.pcn.max_record_history)           # 899
        history_length = min(self.n_samples, self.pcn.max_record_history)                      # 900
        Layers = []
        for i in range(self.n_layers):                                                          # 903-942
            if i == 0:
                init_sample_sym = torch.as_tensor(uStdev_sym, device=dev) * rg.construct_w()
                lay = Layer(True, self.sqrtBeta_0, i, n_samples, self.pcn, init_sample_sym)
                lay.stdev_sym = uStdev_sym
            elif i == n_layers - 1:
                lay = Layer(False, self.sqrtBeta_v, i, self.n_samples, self.pcn, Layers[i - 1].current_sample_sym)
            else:
                lay = Layer(False, float(np.float64(self.sqrtBeta_v) * np.sqrt(sigma_scaling)), i, self.n_samples,
                            self.pcn, Layers[i - 1].current_sample_sym)
            lay.update_current_sample()                                                         # 937
            lay.samples_history = np.empty((history_length, f.basis_number_2D_ravel), dtype=np.complex128)
            lay.sqrt_beta_history = np.empty(history_length, dtype=np.float64)
            lay.record_sqrt_beta()
            Layers.append(lay)
        self.Layers = Layers

    def run(self):                                                                              # 949-1009
        self.accepted_count = 0
        self.accepted_count_SqrtBeta = 0
        if self.printProgress:
            util.printProgressBar(0, self.n_samples, prefix='Preparation . . . . ', suffix='Complete', length=50)
        start_time = time.time()
        start_time_intv = start_time
        average_time_intv = 0.0
        accepted_count_partial = 0
        accepted_count_partial_SqrtBeta = 0
        linalg_error_occured = False
        i = 0
        for i in range(self.n_samples):
            try:
                accepted, accepted_SqrtBeta = self.pcn.one_step(self.Layers)
                accepted_count_partial += accepted
                accepted_count_partial_SqrtBeta += accepted_SqrtBeta
            except np.linalg.LinAlgError as err:
                linalg_error_occured = True
                print("Linear Algebra Error :", err)
                break
            else:
                if (i + 1) % self.evaluation_interval == 0:
                    self.accepted_count += accepted_count_partial
                    self.accepted_count_SqrtBeta += accepted_count_partial_SqrtBeta
                    self.acceptancePercentage = self.accepted_count / (i + 1)
                    if self.enable_step_adaptation:                                               # 977-979
                        self.pcn.adapt_step(self.acceptancePercentage)
                        self.pcn.adapt_step_sqrtBeta(self.accepted_count_SqrtBeta / (i + 1))
                    accepted_count_partial = 0
                    accepted_count_partial_SqrtBeta = 0
                    mTime = (i + 1) / self.evaluation_interval
                    end_time_intv = time.time()
                    average_time_intv += (end_time_intv - start_time_intv - average_time_intv) / mTime
                    start_time_intv = end_time_intv
                    remaining = average_time_intv * ((self.n_samples - i) / self.evaluation_interval)
                    if self.printProgress:
                        util.printProgressBar(i + 1, self.n_samples,
                                              prefix='Time Remaining {0:.0f}s- Acceptance Rate {1:.2%} - Progress:'.format(
                                                  remaining, self.acceptancePercentage), suffix='Complete', length=50)
        if linalg_error_occured:
            if self.printProgress:                                                                # 998-1004 (only when printing)
                end_index = max(i % self.pcn.record_skip, 1)
                for l in range(self.n_layers):
                    self.Layers[l].samples_history = self.Layers[l].samples_history[:end_index, :]
                print('Linear algebra errors occured during some simulation step(s). The simulation result may not be valid')
        else:
            self.total_time = time.time() - start_time

    def save(self, file_name, include_history=False):
        """Simulation.save (1012-1014): same key layout; HDF5 when h5py is available, else .npz (mspde/io.py)."""
        from . import io
        return io.save_simulation(self, file_name)
