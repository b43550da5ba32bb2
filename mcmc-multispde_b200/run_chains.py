#!/usr/bin/env python
"""BASELINE configs[3]: many independent pCN chains sharded over the GPUs of one node, one final NCCL gather of the samples.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 run_chains.py --chains 64 --basis-n 32 \
         --n-layers 3 --n-theta 90 --n-samples 1000 --out chains.npz

Chain c runs on rank c mod G with seed `seed + c` (the scheme of /root/reference/Legacy/runNumbaParallel.py:24-43, there
with mpi4py); a rank's chains share one context and are stepped round-robin (mspde.parallel.MultiChain).  There is no
communication until the final all_gather of every layer's recorded half-samples, the sum-reduction of the acceptance
counters and the pooling of the per-layer Welford field statistics (121-136, 165-180 there).  Works on a single GPU without
torchrun as well."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

import numpy as np
import torch
import torch.distributed as dist

from mspde import image as im, parallel


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--chains", type=int, default=8)
    ap.add_argument("--n-layers", type=int, default=3)
    ap.add_argument("--n", "--basis-n", dest="n", type=int, default=32, help="Fourier modes per axis (use --basis-n under torchrun: its parser treats --n as ambiguous)")
    ap.add_argument("--n-extended", type=int, default=64)
    ap.add_argument("--n-theta", type=int, default=90)
    ap.add_argument("--n-samples", type=int, default=100)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--beta", type=float, default=1.0)
    ap.add_argument("--evaluation-interval", type=int, default=10)
    ap.add_argument("--out", type=str, default="")
    args = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mine = parallel.shard_chains(args.chains, world, rank)
    nr = 2 * args.n * args.n - 2 * args.n + 1
    t0 = time.perf_counter()
    accepted = steps = 0
    stats = None
    if mine:
        # one Simulation (context: H, HtH, y, workspaces) per GPU, shared by all of this rank's chains (parallel.MultiChain)
        sim = im.Simulation(n_layers=args.n_layers, n_samples=args.n_samples, n=args.n, n_extended=args.n_extended, beta=args.beta,
                            kappa=5e9, sigma_0=4e7, sigma_v=3.2e4, sigma_scaling=0.4, meas_std=0.2,
                            evaluation_interval=args.evaluation_interval, printProgress=False,
                            seed=parallel.chain_seed(args.seed, mine[0]), burn_percentage=25.0, enable_step_adaptation=True,
                            pcn_variant="dunlop", phantom_name="shepp.png", meas_type="tomo", n_theta=args.n_theta, verbose=False,
                            device=local)
        sim.pcn.set_chol_epsilon(1e-6)
        mc = parallel.MultiChain(sim, mine, base_seed=args.seed, accumulate_stats=True)
        for _ in range(args.n_samples):
            mc.step_all()
        local_hist = torch.as_tensor(mc.histories())                               # [local chains, layers, samples, n_ravel]
        accepted, steps = mc.accepted_and_steps()
    else:
        mc = None
        local_hist = torch.zeros((0, args.n_layers, args.n_samples, nr), dtype=torch.complex128)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    full = parallel.gather_samples(local_hist.to(dev), args.chains, world)     # [chains, layers, samples, n_ravel]
    acc = parallel.pooled_acceptance(accepted, steps)
    if mc is not None and world == 1:
        stats = mc.pooled_field_statistics()
    elif world > 1:
        # every rank takes part in the pooled-statistics collectives (a rank without chains contributes count 0)
        if mc is None:
            m = 2 * args.n_extended - 1
            z = torch.zeros((m, m), dtype=torch.float64, device=dev)
            stats = []
            for _ in range(args.n_layers):
                res = {}
                for name in ("u", "ell"):
                    n_, mean_, m2_ = parallel.pool_welford(0, z, z)
                    res.update({"count": n_, name + "_mean": mean_, name + "_var": m2_ / max(n_, 1)})
                stats.append(res)
        else:
            stats = mc.pooled_field_statistics()
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        rate = args.chains * args.n_samples / float(t.item())
        print(f"{args.chains} chains x {args.n_samples} samples on {world} GPU(s): {rate:.3f} iterations/s aggregate, "
              f"pooled acceptance {acc:.3f}, gathered {tuple(full.shape)}")
        if args.out:
            extra = {}
            if stats is not None:
                for k, st in enumerate(stats):
                    extra.update({f"layer{k}_u_mean": st["u_mean"].cpu().numpy(), f"layer{k}_u_var": st["u_var"].cpu().numpy(),
                                  f"layer{k}_ell_mean": st["ell_mean"].cpu().numpy(), f"layer{k}_ell_var": st["ell_var"].cpu().numpy()})
                extra["stats_count"] = stats[0]["count"]
            np.savez_compressed(args.out, samples=full.cpu().numpy(), acceptance=acc, chains=args.chains, **extra)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
