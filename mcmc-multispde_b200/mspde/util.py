"""Host-side helpers mirroring /root/reference/mcmc/util_cupy.py for the pCN hot path (torch tensors on the
device, numpy on the host).  Only index/reshape plumbing lives here; arithmetic on the path is in libmspde."""
from __future__ import annotations

import numpy as np
import torch

ORDER = "F"                          # mcmc/image_cupy.py:24
PI = np.float32(np.pi)               # mcmc/util_cupy.py:22
SQRT2 = np.float32(1.41421356)       # mcmc/util_cupy.py:21


def _xp(a):
    return torch if isinstance(a, torch.Tensor) else np


def symmetrize(w_half):
    """util.symmetrize (mcmc/util_cupy.py:201-204): concat(conj(half[:0:-1]), half)."""
    if isinstance(w_half, torch.Tensor):
        return torch.cat((torch.flip(w_half[1:], dims=(0,)).conj().resolve_conj(), w_half))
    return np.concatenate((w_half[:0:-1].conj(), w_half))


def symmetrize_2D(uHalf2D):
    """util.symmetrize_2D (mcmc/util_cupy.py:234-238)."""
    if isinstance(uHalf2D, torch.Tensor):
        w = torch.flip(uHalf2D[:, 1:], dims=(0, 1)).conj().resolve_conj()
        return torch.hstack((w, uHalf2D))
    w = uHalf2D[:, 1:]
    return np.hstack((w[::-1, :][:, ::-1].conj(), uHalf2D))


def from_u_2D_ravel_to_uHalf_2D(u, n):
    """util.from_u_2D_ravel_to_uHalf_2D (mcmc/util_cupy.py:245-246): reshape(il, il, 'F')[:, n-1:]."""
    il = 2 * n - 1
    if isinstance(u, torch.Tensor):
        return u.reshape(il, il).transpose(0, 1)[:, n - 1:]
    return u.reshape(il, il, order=ORDER)[:, n - 1:]


def ravel_F(a2d):
    """ndarray.ravel('F') for a 2-D torch tensor / numpy array."""
    if isinstance(a2d, torch.Tensor):
        return a2d.transpose(0, 1).reshape(-1)
    return a2d.ravel(ORDER)


def extend2D(uIn, num):
    """util.extend2D (mcmc/util_cupy.py:251-267) for the square (symmetrised) case."""
    xp = _xp(uIn)
    n = (uIn.shape[0] + 1) // 2
    if num <= n:
        return uIn
    if xp is torch:
        z = torch.zeros((2 * num - 1, 2 * num - 1), dtype=uIn.dtype, device=uIn.device)
    else:
        z = np.zeros((2 * num - 1, 2 * num - 1), dtype=uIn.dtype)
    z[(num - 1) - (n - 1):(num - 1) + n, (num - 1) - (n - 1):(num - 1) + n] = uIn
    return z


def w_half_from_normals(re, im):
    """util.construct_w_Half (mcmc/util_cupy.py:34-39) applied to injected standard normals."""
    w = re.astype(np.float64) + 1j * im.astype(np.float64)
    w[0] = 2.0 * w[0].real
    return w / np.float64(SQRT2)


def printProgressBar(iteration, total, prefix="", suffix="", decimals=1, length=100, fill="#"):
    """Same call signature as the reference's progress printer (mcmc/util_cupy.py:139-157); one carriage-return-refreshed line."""
    import sys
    frac = 0.0 if total <= 0 else min(1.0, max(0.0, iteration / float(total)))
    done = int(round(length * frac))
    sys.stdout.write("\r{} [{}{}] {:.{d}f}% {}".format(prefix, fill * done, "." * (length - done), 100.0 * frac, suffix, d=decimals))
    if iteration >= total:
        sys.stdout.write("\n")
    sys.stdout.flush()
