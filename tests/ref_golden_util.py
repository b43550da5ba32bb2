"""Shared loader for tests/golden/ref_step_*.npz -- vectors computed by the REFERENCE'S OWN classes
(``oracle/make_ref_goldens.py`` imports /root/reference/mcmc/image_cupy.py unmodified through ``oracle/ref_shim.py``).

Rebuilds, from one golden file, what the oracle (CPU test) and the CUDA path (GPU test) need to replay the same steps:
the fixed inputs, the chain state the reference had before its first ``pCN.one_step``, and the injected noise of every
step in the reference's own draw order (image_cupy.py:558, 573, 583, 585; 668, 684, 699)."""
from __future__ import annotations

import glob
import os

import numpy as np

from oracle import mspde_oracle as o

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# Cases whose CUDA-path tests ran on B200s (profiles/r02_gputest_c.log).  Goldens added later pin the oracle on CPU only until
# a GPU run has covered them; tests/test_gpu_ref_goldens.py enumerates gpu_cases().
GPU_VERIFIED = ("n16", "n2", "n3", "n4", "n4k2", "n4k4", "n8")


def gpu_cases(mode):
    return [c for c in cases(mode) if c in GPU_VERIFIED]


def cases(mode):
    return sorted(os.path.basename(p)[len("ref_step_"):-len(f"_{mode}.npz")] for p in glob.glob(os.path.join(GOLDEN, f"ref_step_*_{mode}.npz")))


def checksum(a):
    a = np.asarray(a)
    w = np.cos(np.arange(a.size, dtype=np.float64) * 0.37).reshape(a.shape)
    return np.array([a.real.sum(), a.imag.sum() if np.iscomplexobj(a) else 0.0, np.abs(a).sum(),
                     (a.real * w).sum(), ((a.imag if np.iscomplexobj(a) else 0 * a.real) * w).sum()])


def checksum_close(a, ref, rtol):
    """Position-weighted sums agree relative to sum |a| (entry 2)."""
    c = checksum(a)
    return np.abs(c - ref).max() <= rtol * abs(ref[2])


class RefGolden:
    def __init__(self, case, mode):
        self.g = g = np.load(os.path.join(GOLDEN, f"ref_step_{case}_{mode}.npz"))
        self.case, self.mode = case, mode
        self.n, self.n_ext, self.n_theta, self.K = int(g["n"]), int(g["n_ext"]), int(g["n_theta"]), int(g["n_layers"])
        self.ft = o.FourierTables(self.n, self.n_ext)
        self.M = self.ft.m * self.n_theta
        self.y, self.delta = g["y"], float(g["delta"])
        self.n_steps = len(g["accepted"])
        self.n_sb_steps = len(g["b_accepted"])
        self._H = None

    # ---- fixed inputs.  H is stored for the small cases; otherwise it is regenerated from the closed form and checked
    # against the stored rows + checksums of the reference's matrix before use.
    def inputs(self):
        if self._H is None:
            g = self.g
            if "H" in g.files:
                H, HtH = g["H"], g["HtH"]
            else:
                H, HtH = o.normalised_inputs(o.tomography_H(self.ft, self.n_theta), float(g["meas_std"]))
                if self.mode == "native":       # the reference stores H and HtH as complex64 (image_cupy.py:396, 507)
                    H, HtH = H.astype(np.complex64).astype(np.complex128), HtH.astype(np.complex64).astype(np.complex128)
                tol = 1e-13 if self.mode == "alias" else 1e-6
                assert np.abs(H[g["H_rows_idx"]] - g["H_rows"]).max() <= tol * np.abs(g["H_rows"]).max()
                assert checksum_close(H, g["H_checksum"], tol) and checksum_close(HtH, g["HtH_checksum"], 10 * tol)
            self._H = (H, HtH)
        return self._H

    def beta_for_step(self, s, prefix="s"):
        """The step size the reference used in plain step ``s`` ('s') or in the sqrt-beta steps ('b')."""
        g = self.g
        if prefix == "s":
            return float(g["beta_used"][s])
        if "adapt_after_step" in g.files and int(g["adapt_after_step"]) == self.n_steps - 1:
            return float(g["adapt_beta"])
        return float(g["beta_used"][-1])

    def state(self, prefix="init", beta=None):
        """ChainState as the reference held it after ``prefix`` ('init', 's0', 's1', ...), tables taken from the reference;
        ``beta`` = the step size of the step that will be run from it."""
        g, K, ft = self.g, self.K, self.ft
        beta = float(g["beta0"]) if beta is None else float(beta)
        st = o.ChainState(beta=beta, betaZ=float(np.sqrt(1 - beta * beta)), sqrt_betas=[float(x) for x in g["sqrt_betas"]],
                          stdev_sym=g["stdev_sym"].copy(), noise_cur=[None] * K, noise_new=[None] * K, u_cur_sym=[None] * K,
                          u_new_sym=[None] * K, L_cur=[None] * K, L_latest=[None] * K)
        for k in range(K):
            st.noise_cur[k] = g[f"{prefix}_noise_cur_{k}"].copy()
            st.noise_new[k] = (g[f"{prefix}_noise_new_{k}"] if prefix != "init" else g[f"init_noise_cur_{k}"]).copy()
            st.u_cur_sym[k] = g[f"{prefix}_u_cur_{k}"].copy()
            st.u_new_sym[k] = g[f"{prefix}_u_new_{k}"].copy()
            if k >= 1:
                # invariant of plain steps: current_L[k] was assembled from the current sample of layer k-1 (753, 577-579)
                st.L_cur[k] = o.assemble_L(st.u_cur_sym[k - 1], st.sqrt_betas[k], ft)
                st.L_latest[k] = st.L_cur[k]
        return st

    def _split(self, prefix):
        d, sizes = self.g[f"{prefix}_draws"], self.g[f"{prefix}_draw_sizes"]
        out, p = [], 0
        for s in sizes:
            out.append(d[p:p + s])
            p += s
        return out

    def step_noise(self, s, prefix="s"):
        """Noise of plain step ``s`` in oracle form.  For 'b' steps also the hyper-move draws (z, e2, log_u2)."""
        parts = self._split(f"{prefix}{s}")
        K = self.K
        w = [o.w_half_from_normals(parts[2 * k], parts[2 * k + 1]) for k in range(K - 1)]
        p = 2 * (K - 1)
        nz = dict(w_half=w, log_u=float(np.log(parts[p][0])), w_last=o.w_half_from_normals(parts[p + 1], parts[p + 2]), e=parts[p + 3])
        if prefix == "b":
            nz["move"] = dict(z=parts[p + 4], e=parts[p + 5], log_u=float(np.log(parts[p + 6][0])))
        return nz

    def init_draws(self):
        """(init_w_half, init_solve_w_half) per layer, from Simulation.__init__'s draw order (905, 735, 755)."""
        g, K = self.g, self.K
        d, sizes = g["init_draws"], g["init_draw_sizes"]
        parts, p = [], 0
        for s in sizes:
            parts.append(d[p:p + s])
            p += s
        parts = parts[2:]                      # image noise (333) and sinogram noise (401)
        noise, solve = [None] * K, [None] * K
        q = 0
        for k in range(K):
            if k == 0:
                solve[0] = o.w_half_from_normals(parts[q], parts[q + 1])
                noise[0] = o.w_half_from_normals(parts[q + 2], parts[q + 3])
            else:
                noise[k] = o.w_half_from_normals(parts[q], parts[q + 1])
                solve[k] = o.w_half_from_normals(parts[q + 2], parts[q + 3])
            q += 4
        return noise, solve
