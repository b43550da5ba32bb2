"""result files, chain continuation and FBP initialisation around the hot path (host side, setup only).

* save_simulation / load_last_samples keep the key layout of util._save_object (/root/reference/mcmc/util_cupy.py:553-583,
  SURVEY Appendix A.9): top-level scalars, groups 'fourier', 'measurement', 'pcn', 'Layers {i}' with 'samples_history'.
  HDF5 is written when h5py is importable (it is not in this image), otherwise an .npz with the same slash-separated keys.
* initialize_from_file mirrors twoDsim.initialize_from_folder (/root/reference/twoDsim.py:36-54).
* initialize_using_FBP mirrors twoDsim.initialize_using_FBP (25-34) with a numpy filtered back-projection standing in for
  skimage.transform.iradon(circle=True) (ramp filter, linear interpolation) -- unpinned against skimage.
"""
from __future__ import annotations

import numpy as np
import torch

from . import util

ORDER = util.ORDER


def _np(x):
    return x.detach().cpu().numpy() if isinstance(x, torch.Tensor) else np.asarray(x)


def _umask(n):
    """FourierAnalysis_2D._Umask (image_cupy.py:83-85): kron(toeplitz(x), toeplitz(x)) as int16, x = [1..n, 0...]."""
    x = np.concatenate((np.arange(n) + 1, np.zeros(n - 1)))
    idx = np.abs(np.arange(2 * n - 1)[:, None] - np.arange(2 * n - 1)[None, :])
    toep = x[idx]
    return np.kron(toep, toep).astype(np.int16)


def _payload(sim, reference_quirks=True):
    """Every key util._save_object writes for a Simulation (util_cupy.py:553-583; the key set is pinned by
    tests/golden/ref_save_layout.json, recorded from the reference's own save() through a recording h5py stand-in):
    scalar attributes of the Simulation at the top level, the groups 'fourier', 'measurement', 'pcn' and 'Layers {i}' one
    level deep, arrays named in excluded_matrix (554) left out.  The reference also writes three large arrays it never reads
    back (pcn/H_conj_T, fourier/_Umask, measurement/tx|ty); reference_quirks=False drops those."""
    f, m, p = sim.fourier, sim.measurement, sim.pcn
    n_ext = f.extended_basis_number
    out = {}
    for k in ("n_samples", "evaluation_interval", "burn_percentage", "random_seed", "printProgress", "n_layers", "kappa", "sigma_0",
              "sigma_v", "sigma_scaling", "enable_step_adaptation", "verbose", "hybrid", "acceptancePercentage", "accepted_count",
              "accepted_count_SqrtBeta", "d", "nu", "alpha", "t_start", "t_end", "beta_u", "beta_v", "pcn_variant", "total_time"):
        out[k] = getattr(sim, k)
    out["sqrtBeta_v"], out["sqrtBeta_0"] = np.float32(sim.sqrtBeta_v), np.float32(sim.sqrtBeta_0)      # 0-d CuPy arrays there (559-566)
    for k in ("basis_number", "extended_basis_number", "basis_number_2D", "basis_number_2D_ravel", "basis_number_2D_sym",
              "extended_basis_number_2D", "extended_basis_number_2D_sym", "t_end", "t_start", "verbose"):
        out["fourier/" + k] = getattr(f, k)
    out["fourier/D_diag"], out["fourier/Kmatrix"] = f.D_diag, f.Kmatrix
    dim = m.dim
    t = np.linspace(0., 1., num=dim, endpoint=True)
    tx, ty = np.meshgrid(t, t)
    out.update({"measurement/dim": dim, "measurement/stdev": m.stdev, "measurement/num_sample": int(m.num_sample),
                "measurement/n_theta": m.n_theta, "measurement/n_r": m.n_r, "measurement/use_skimage": False,
                "measurement/target_image": np.asarray(m.target_image, dtype=np.float32),
                "measurement/corrupted_image": np.asarray(m.corrupted_image, dtype=np.float32),
                "measurement/v": np.asarray(m.v), "measurement/theta": np.asarray(m.theta), "measurement/r": np.asarray(m.r),
                "measurement/pure_sinogram": np.asarray(m.pure_sinogram), "measurement/sinogram": np.asarray(m.sinogram)})
    yBar = _np(p.yBar)
    out.update({"pcn/n_layers": p.n_layers, "pcn/beta": np.float32(p.beta), "pcn/betaZ": np.float32(p.betaZ), "pcn/variant": p.variant,
                "pcn/meas_var": p.meas_var, "pcn/verbose": bool(p.verbose), "pcn/hybrid_mode": False, "pcn/epsilon": float(p.epsilon),
                "pcn/cholesky_stabilizer": float(p.cholesky_stabilizer), "pcn/yBar": yBar, "pcn/y_": _np(p.y), "pcn/yBar_": yBar,
                "pcn/aggresiveness": p.aggresiveness, "pcn/target_acceptance_rate": p.target_acceptance_rate,
                "pcn/beta_feedback_gain": p.beta_feedback_gain, "pcn/record_skip": int(p.record_skip),
                "pcn/record_count": int(p.record_count), "pcn/max_record_history": p.max_record_history,
                "pcn/pcn_step_sqrtBetas": p.pcn_step_sqrtBetas, "pcn/pcn_step_sqrtBetas_Z": float(p.pcn_step_sqrtBetas_Z),
                "pcn/stdev_sqrtBetas": np.asarray(p.stdev_sqrtBetas, dtype=np.float32),
                "pcn/use_beta_adaptation": bool(p.use_beta_adaptation),
                "pcn/use_naive_logRatio_implementation": bool(p.use_naive_logRatio_implementation)})
    if reference_quirks:
        out["fourier/_Umask"] = _umask(f.basis_number)
        out["measurement/tx"], out["measurement/ty"] = tx.ravel(ORDER), ty.ravel(ORDER)
        out["pcn/H_conj_T"] = _np(p.H_conj_T).astype(np.complex64)
    nr = f.basis_number_2D_ravel
    for k, lay in enumerate(sim.Layers):
        g = f"Layers {k}/"
        sd = np.asarray(lay.stdev_sym_host)
        out.update({g + "is_stationary": bool(lay.is_stationary), g + "sqrt_beta": np.float32(lay.sqrt_beta), g + "order_number": lay.order_number,
                    g + "n_samples": lay.n_samples, g + "i_record": lay.i_record, g + "stdev": sd[nr - 1:], g + "stdev_sym": sd,
                    g + "samples_history": lay.samples_history.astype(np.complex64),               # complex64 there (729, 939)
                    g + "sqrt_beta_history": lay.sqrt_beta_history.astype(np.float32),
                    g + "current_noise_sample": _np(lay.current_noise_sample), g + "new_noise_sample": _np(lay.new_noise_sample),
                    g + "current_sample": _np(lay.current_sample), g + "current_sample_sym": _np(lay.current_sample_sym),
                    g + "new_sample": _np(lay.new_sample), g + "new_sample_sym": _np(lay.new_sample_sym),
                    g + "new_log_L_det": float(lay.new_log_L_det), g + "current_log_L_det": float(lay.current_log_L_det),
                    g + "new_sample_scaled_norm": float(getattr(lay, "new_sample_scaled_norm", 0.0)),
                    g + "current_sample_scaled_norm": float(getattr(lay, "current_sample_scaled_norm", 0.0))})
        if lay.stats is not None:                   # extension (f4): on-device posterior statistics, not in the reference's files
            st = lay.field_statistics()
            for name in ("u_mean", "u_var", "ell_mean", "ell_var"):
                out[g + name] = st[name].cpu().numpy()
    return out


def save_simulation(sim, file_name: str, reference_quirks: bool = True):
    data = _payload(sim, reference_quirks)
    if file_name.endswith((".hdf5", ".h5")):
        try:
            import h5py
        except ImportError:
            file_name = file_name.rsplit(".", 1)[0] + ".npz"
        else:
            with h5py.File(file_name, "w") as f:
                for key, val in data.items():
                    if isinstance(val, np.ndarray) and val.ndim > 0:
                        f.create_dataset(key, data=val, compression="gzip")
                    else:
                        f.create_dataset(key, data=val)
            return file_name
    np.savez_compressed(file_name, **{k: np.asarray(v) for k, v in data.items()})
    return file_name


def load_last_samples(file_name: str):
    """Returns the list of last recorded half-samples per layer (what initialize_from_folder reads, twoDsim.py:46-49)."""
    if file_name.endswith((".hdf5", ".h5")):
        import h5py
        with h5py.File(file_name, "r") as f:
            n_layers = int(f["n_layers"][()])
            return [np.asarray(f[f"Layers {i}/samples_history"][()])[-1] for i in range(n_layers)]
    with np.load(file_name, allow_pickle=False) as d:
        n_layers = int(d["n_layers"])
        return [d[f"Layers {i}/samples_history"][-1] for i in range(n_layers)]


def initialize_from_file(sim, file_name: str):
    """twoDsim.initialize_from_folder (36-54): current_sample_sym <- symmetrize(last recorded half sample), then record."""
    for i, half in enumerate(load_last_samples(file_name)):
        sym = util.symmetrize(np.asarray(half, dtype=np.complex128))
        sim.Layers[i].current_sample_sym = torch.as_tensor(sym)
        sim.Layers[i].record_sample()


def fbp(sinogram: np.ndarray, theta_deg: np.ndarray) -> np.ndarray:
    """Filtered back-projection standing in for ``skimage.transform.iradon(sinogram, theta, circle=True)`` (twoDsim.py:27).
    scikit-image is a third-party dependency of the reference that is absent from this image (and unpinned in the reference:
    no requirements file), so this restates its published algorithm (iradon of scikit-image >= 0.16): zero-pad the detector
    axis to max(64, next power of two of 2 n_r); ramp filter built from the band-limited spatial impulse response
    (f[0] = 1/4, f[odd k] = -1/(pi k)^2, filter = 2 Re fft(f)); linear interpolation of every filtered projection at
    t = y cos(a) - x sin(a); zero outside the reconstruction circle; scale pi / (2 n_theta).  Unpinned against a real skimage
    output; tests/test_host_logic.py checks it reconstructs a disc phantom from its analytic sinogram."""
    n_r, n_theta = sinogram.shape
    size = max(64, int(2 ** np.ceil(np.log2(2 * n_r))))
    pad = np.zeros((size, n_theta))
    pad[:n_r] = sinogram
    k = np.concatenate((np.arange(1, size / 2 + 1, 2, dtype=int), np.arange(size / 2 - 1, 0, -2, dtype=int)))
    f = np.zeros(size)
    f[0] = 0.25
    f[1::2] = -1.0 / (np.pi * k) ** 2
    filt = 2 * np.real(np.fft.fft(f))
    proj = np.real(np.fft.ifft(np.fft.fft(pad, axis=0) * filt[:, None], axis=0))[:n_r]
    radius = n_r // 2
    xpr, ypr = np.mgrid[:n_r, :n_r] - radius
    x = np.arange(n_r) - n_r // 2
    out = np.zeros((n_r, n_r))
    for col, ang in zip(proj.T, np.deg2rad(theta_deg)):
        t = ypr * np.cos(ang) - xpr * np.sin(ang)
        out += np.interp(t.ravel(), x, col, left=0.0, right=0.0).reshape(n_r, n_r)
    out[(xpr ** 2 + ypr ** 2) > radius ** 2] = 0.0
    return out * np.pi / (2 * n_theta)


def initialize_using_FBP(sim, image=None):
    """twoDsim.initialize_using_FBP (25-34): last layer's current sample <- symmetrised Fourier coefficients of the FBP image
    (rfft2 -> symmetrize_2D -> ravel('F')), then one record_sample().  image: optional precomputed FBP image."""
    if image is None:
        image = fbp(np.asarray(sim.measurement.sinogram), np.asarray(sim.measurement.theta))
    half = sim.fourier.rfft2(torch.as_tensor(np.asarray(image, dtype=np.float64)))
    sym2d = util.symmetrize_2D(half)
    sym = util.ravel_F(sym2d)
    sim.Layers[-1].current_sample_sym = sym
    sim.Layers[-1].record_sample()
    return image
