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


def _payload(sim):
    out = {k: v for k, v in sim.__dict__.items() if isinstance(v, (int, float, str, bool, np.floating, np.integer))}
    out["fourier/basis_number"] = sim.fourier.basis_number
    out["fourier/extended_basis_number"] = sim.fourier.extended_basis_number
    out["measurement/sinogram"] = sim.measurement.sinogram
    out["measurement/theta"] = sim.measurement.theta
    out["measurement/target_image"] = sim.measurement.target_image
    out["measurement/corrupted_image"] = sim.measurement.corrupted_image
    out["measurement/stdev"] = sim.measurement.stdev
    out["pcn/beta"] = sim.pcn.beta
    out["pcn/record_skip"] = sim.pcn.record_skip
    for k, lay in enumerate(sim.Layers):
        out[f"Layers {k}/samples_history"] = lay.samples_history[:max(lay.i_record, 1)]
        out[f"Layers {k}/sqrt_beta_history"] = lay.sqrt_beta_history[:max(lay.i_record, 1)]
        out[f"Layers {k}/sqrt_beta"] = lay.sqrt_beta
        if lay.stats is not None:
            st = lay.field_statistics()
            for name in ("u_mean", "u_var", "ell_mean", "ell_var"):
                out[f"Layers {k}/{name}"] = st[name].cpu().numpy()
    return out


def save_simulation(sim, file_name: str):
    data = _payload(sim)
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
    """Filtered back-projection of an (n_r x n_theta) sinogram onto an n_r x n_r grid (ramp filter, linear interpolation,
    reconstruction circle), the numpy stand-in for skimage.transform.iradon(sinogram, theta, circle=True)."""
    n_r, n_theta = sinogram.shape
    size = max(64, int(2 ** np.ceil(np.log2(2 * n_r))))
    pad = np.zeros((size, n_theta))
    pad[:n_r] = sinogram
    freqs = np.fft.fftfreq(size)
    filt = 2 * np.abs(freqs)
    proj = np.real(np.fft.ifft(np.fft.fft(pad, axis=0) * filt[:, None], axis=0))[:n_r]
    mid = n_r // 2
    xpr, ypr = np.mgrid[:n_r, :n_r] - mid
    out = np.zeros((n_r, n_r))
    for k, ang in enumerate(np.deg2rad(theta_deg)):
        t = ypr * np.cos(ang) - xpr * np.sin(ang) + mid
        out += np.interp(t.ravel(), np.arange(n_r), proj[:, k], left=0.0, right=0.0).reshape(n_r, n_r)
    out[(xpr ** 2 + ypr ** 2) > mid ** 2] = 0.0
    return out * np.pi / (2 * n_theta)


def initialize_using_FBP(sim):
    """twoDsim.initialize_using_FBP (25-34): last layer's current sample <- Fourier coefficients of the FBP image."""
    image = fbp(np.asarray(sim.measurement.sinogram), np.asarray(sim.measurement.theta))
    half = sim.fourier.rfft2(torch.as_tensor(image))
    sym2d = util.symmetrize_2D(half)
    sym = util.ravel_F(sym2d)
    sim.Layers[-1].current_sample_sym = sym
    sim.Layers[-1].record_sample()
    return image
