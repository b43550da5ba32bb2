"""Chain-parallel sharding: independent chains, chain c -> rank c mod G, no data-path collective; one
final all_gather of the recorded samples (precedent: Legacy/runNumbaParallel.py:24-43, 121-136, mpi4py).

Works with any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_chains(n_chains: int, world: int, rank: int) -> list[int]:
    """Chain ids owned by `rank` (round-robin)."""
    return [c for c in range(n_chains) if c % world == rank]


def chain_seed(base_seed: int, chain: int) -> int:
    """Seeds differ per chain like `seed-1+rank` in Legacy/runNumbaParallel.py:28."""
    return int(base_seed) + int(chain)


def gather_samples(local: torch.Tensor, n_chains: int, world: int | None = None) -> torch.Tensor:
    """local: [chains_on_this_rank, ...] histories of this rank's chains (same trailing shape on every rank).
    Returns [n_chains, ...] ordered by chain id on every rank.  Ranks may own different numbers of chains."""
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank()
    per = (n_chains + world - 1) // world
    pad = torch.zeros((per,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    if local.is_complex():
        buf = torch.view_as_real(pad).contiguous()
    else:
        buf = pad.contiguous()
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    res = torch.zeros((n_chains,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r in range(world):
        ids = shard_chains(n_chains, world, r)
        blk = torch.view_as_complex(out[r]) if local.is_complex() else out[r]
        for slot, c in enumerate(ids):
            res[c] = blk[slot]
    return res


def pooled_acceptance(accepted_local: int, steps_local: int) -> float:
    """Sum-reduced acceptance rate over all chains (the MPI.Reduce(SUM) of the Legacy runner)."""
    t = torch.tensor([float(accepted_local), float(steps_local)], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda()
        dist.all_reduce(t)
    return float(t[0] / max(t[1], 1.0))


# ----------------------------------------------------------------------------------------------------------------------
# Pooled posterior statistics (Legacy/runNumbaParallel.py:121-136, 165-180: per-chain means / variances reduced to rank 0
# with MPI.Reduce(SUM) and broadcast back).  Here: Welford (count, mean, M2) triples merged with the parallel-variance formula
# of Chan et al., first across the chains of one rank, then across ranks with two all_reduce(SUM) calls.
# ----------------------------------------------------------------------------------------------------------------------
def merge_welford(parts):
    """parts: iterable of (count, mean, m2) with tensors of equal shape.  Returns the merged (count, mean, m2):
    n = sum n_i; mean = sum n_i mean_i / n; M2 = sum (M2_i + n_i (mean_i - mean)^2)."""
    parts = [p for p in parts if p[0] > 0]
    if not parts:
        return 0, None, None
    n = sum(p[0] for p in parts)
    mean = sum(p[1] * (p[0] / n) for p in parts)
    m2 = sum(p[2] + p[0] * (p[1] - mean) ** 2 for p in parts)
    return n, mean, m2


def pool_welford(count: int, mean: torch.Tensor, m2: torch.Tensor, group=None):
    """All ranks contribute one (count, mean, M2) triple (count may be 0, then mean / m2 are ignored but must have the common
    shape); every rank receives the pooled triple.  Two all_reduce(SUM): [n, n*mean] then [M2 + n (mean - pooled mean)^2]."""
    if not (dist.is_initialized() and dist.get_world_size(group) > 1):
        return count, mean, m2
    c = float(count)
    if count == 0:
        mean, m2 = torch.zeros_like(mean), torch.zeros_like(m2)
    head = torch.cat((torch.tensor([c], dtype=mean.dtype, device=mean.device), (mean * c).reshape(-1)))
    dist.all_reduce(head, group=group)
    n = float(head[0].item())
    pooled_mean = (head[1:] / max(n, 1.0)).reshape(mean.shape)
    part = (m2 + c * (mean - pooled_mean) ** 2).contiguous()
    dist.all_reduce(part, group=group)
    return int(round(n)), pooled_mean, part


class MultiChain:
    """Several independent chains on ONE GPU sharing one context (H, HtH, y and every workspace are per problem, not per chain;
    a chain owns only its state: 4 vectors per layer + two L buffers per non-stationary layer, ~1 GB at n=32 / 3 layers).
    Each chain keeps its own RNG stream, step size beta and acceptance counters, exactly like one rank of the Legacy MPI runner
    (Legacy/runNumbaParallel.py:24-43); steps are issued round-robin, every step already fills the GPU through its three
    internal streams."""

    def __init__(self, sim, chain_ids, base_seed: int, accumulate_stats: bool = False):
        from . import image as im
        self.sim, self.pcn = sim, sim.pcn
        self.ids = list(chain_ids)
        pcn, f = sim.pcn, sim.fourier
        dev = pcn.ctx.device
        self.chains = []
        sd = sim.Layers[0].stdev_sym_host
        for slot, c in enumerate(self.ids):
            rg = im.RandomGenerator_2D(f.basis_number, device=dev, seed=chain_seed(base_seed, c))
            if slot == 0 and getattr(sim, "_multichain_reuse_first", True):
                Layers = sim.Layers                          # the Simulation's own chain is chain ids[0]
            else:
                pcn.random_gen = rg
                Layers = []
                for i in range(sim.n_layers):                # Simulation.__init__ 903-942 for one more chain
                    if i == 0:
                        init = torch.as_tensor(sd, device=dev) * rg.construct_w()
                        lay = im.Layer(True, sim.Layers[0].sqrt_beta, 0, sim.n_samples, pcn, init)
                        lay.stdev_sym_host = sd
                    else:
                        lay = im.Layer(False, sim.Layers[i].LMat.sqrt_beta, i, sim.n_samples, pcn, Layers[i - 1].current_sample_sym)
                    lay.update_current_sample()
                    lay.record_sqrt_beta()
                    Layers.append(lay)
            self.chains.append(dict(id=c, Layers=Layers, rg=rg if Layers is not sim.Layers else sim.random_gen, beta=pcn.beta,
                                    betaZ=pcn.betaZ, accepted=0, accepted_partial=0, steps=0, bind=None))
        pcn.random_gen = sim.random_gen
        pcn.accumulate_stats = accumulate_stats
        self._binds = {}

    def step_all(self, noise=None):
        """One pCN iteration of every local chain (noise: optional list of injected per-chain noise tuples)."""
        pcn = self.pcn
        for slot, ch in enumerate(self.chains):
            pcn.beta, pcn.betaZ, pcn.random_gen = ch["beta"], ch["betaZ"], ch["rg"]
            pcn._chain = ch["bind"]
            pcn._chain_layers = ch["Layers"] if ch["bind"] is not None else None
            acc, _ = pcn.one_step(ch["Layers"], noise=None if noise is None else noise[slot])
            ch["bind"] = pcn._chain
            ch["accepted_partial"] += acc
            ch["steps"] += 1
            if ch["steps"] % self.sim.evaluation_interval == 0:          # Simulation.run 975-982, per chain
                ch["accepted"] += ch["accepted_partial"]
                ch["accepted_partial"] = 0
                if self.sim.enable_step_adaptation:
                    pcn.adapt_step(ch["accepted"] / ch["steps"])
                    ch["beta"], ch["betaZ"] = pcn.beta, pcn.betaZ

    def histories(self):
        """[local chains, layers, recorded samples, n_ravel] complex128 (host)."""
        import numpy as np
        n = min(l.i_record for ch in self.chains for l in ch["Layers"])
        return np.stack([np.stack([l.samples_history[:n] for l in ch["Layers"]]) for ch in self.chains])

    def accepted_and_steps(self):
        return (sum(ch["accepted"] + ch["accepted_partial"] for ch in self.chains), sum(ch["steps"] for ch in self.chains))

    def pooled_field_statistics(self, group=None):
        """Per layer: Welford statistics of u(x) and exp(u(x)) pooled over every chain of every rank (needs
        accumulate_stats=True).  Returns a list (one dict per layer) with count, u_mean, u_var, ell_mean, ell_var."""
        out = []
        K = self.sim.n_layers
        for k in range(K):
            res = {}
            for key_mean, key_m2, name in (("mean_u", "m2_u", "u"), ("mean_ell", "m2_ell", "ell")):
                parts = [(ch["Layers"][k].stats["count"], ch["Layers"][k].stats[key_mean], ch["Layers"][k].stats[key_m2])
                         for ch in self.chains if ch["Layers"][k].stats is not None]
                n, mean, m2 = merge_welford(parts)
                if mean is None:
                    m = self.pcn.ctx.m
                    mean = torch.zeros((m, m), dtype=torch.float64, device=self.pcn.ctx.device)
                    m2 = torch.zeros_like(mean)
                n, mean, m2 = pool_welford(n, mean, m2, group)
                res.update({"count": n, name + "_mean": mean, name + "_var": m2 / max(n, 1)})
            out.append(res)
        return out
