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
