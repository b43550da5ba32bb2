"""torchrun tool: the block-distributed sampler (mspde.dist.DistPCN) against (a) the reference-generated golden chain and
(b) the single-GPU mspde_pcn_step on the same inputs.  Prints one JSON line on rank 0.

  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/dist_chain_check.py [golden case] [n_big]

(a) replays tests/golden/ref_step_<case>_alias.npz (vectors computed by the reference's own classes) on all ranks: accept
    decisions must be identical, Phi within 1e-10, samples within 1e-10 of the golden.
(b) builds a synthetic chain at n = n_big (default 32), runs one step on the distributed sampler and the same step through the
    single-GPU kernels on rank 0: Phi relative difference, bit-identity of the LU-solved and QR-solved samples.
"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch, torch.distributed as dist


def rel(a, b):
    a = a.cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def golden_replay(case, dev):
    from ref_golden_util import RefGolden
    from mspde.dist import DistPCN
    R = RefGolden(case, "alias")
    g, K = R.g, R.K
    H, HtH = R.inputs()
    to = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    Hd = torch.zeros((R.M, (R.ft.N + 7) // 8 * 8), dtype=torch.complex128, device=dev)[:, :R.ft.N]
    Hd.copy_(to(H))
    ch = DistPCN(R.n, R.n_ext, K, Hd, to(R.y), g["D_diag"], g["stdev_sym"], g["lmat_sqrt_betas"], dev, delta=R.delta,
                 beta=R.beta_for_step(0), nb=64)
    ch.init_state([g[f"init_noise_cur_{k}"] for k in range(K)], [g[f"init_u_cur_{k}"] for k in range(K)],
                  [g[f"init_u_new_{k}"] for k in range(K)])
    out = dict(case=case, steps=R.n_steps, accept_equal=True, phi_rel=0.0, u_rel=0.0, noise_rel=0.0)
    for s in range(R.n_steps):
        nz = R.step_noise(s)
        ch.set_step(R.beta_for_step(s))
        d = ch.one_step([to(w) for w in nz["w_half"]], nz["log_u"], to(nz["w_last"]), to(nz["e"]))
        ph = g[f"s{s}_phi"]
        out["accept_equal"] &= (int(d["accepted"]) == int(g["accepted"][s]))
        out["phi_rel"] = max(out["phi_rel"], abs(d["phi_cur"] / ph[0] - 1), abs(d["phi_new"] / ph[1] - 1))
        for k in range(K):
            out["u_rel"] = max(out["u_rel"], rel(ch.u_new[k], g[f"s{s}_u_new_{k}"]), rel(ch.u_cur[k], g[f"s{s}_u_cur_{k}"]))
            out["noise_rel"] = max(out["noise_rel"], rel(ch.noise_new[k], g[f"s{s}_noise_new_{k}"]))
    return out


def single_vs_dist(n, dev, rank):
    """One step of a synthetic 3-layer chain at basis size n: distributed sampler on all ranks vs mspde_pcn_step on rank 0."""
    from mspde.dist import build_dist_chain
    from mspde.core import ChainBuffers, MspdeContext
    n_ext, n_theta, K = max(n, 64), 90 if n >= 32 else 16, 3
    ch, noise_fn = build_dist_chain(n, n_ext, n_theta, K, dev, seed=5, beta=0.3)
    state = dict(noise_cur=[x.clone() for x in ch.noise_cur], u_cur=[x.clone() for x in ch.u_cur], u_new=[x.clone() for x in ch.u_new],
                 noise_new=[x.clone() for x in ch.noise_new])
    nz = noise_fn()
    d = ch.one_step(*nz)
    out = dict(n=n, N=ch.N, M=ch.M, dist=dict(accepted=bool(d["accepted"]), phi_cur=d["phi_cur"], phi_new=d["phi_new"]))
    if rank == 0:
        ctx = MspdeContext(n, n_ext, ch.M, K, device=dev.index)
        t2 = ctx.gemm_tn(ch.H, ch.H, conjA=True)
        HtH = 0.5 * (t2 + t2.conj().transpose(0, 1))
        ctx.set_inputs(ch.H, HtH, ch.y, ch._D, ch.stdev_sym, ch.delta)
        del t2, HtH
        cb = ChainBuffers(ctx, ch.sqrt_betas, ch.beta)
        for k in range(K):
            cb.noise_cur[k].copy_(state["noise_cur"][k]); cb.noise_new[k].copy_(state["noise_new"][k])
            cb.u_cur[k].copy_(state["u_cur"][k]); cb.u_new[k].copy_(state["u_new"][k])
            if k >= 1:
                ctx.construct_L(cb.u_cur[k - 1], ch.sqrt_betas[k], out=cb.L_cur[k]); cb.L_latest[k].copy_(cb.L_cur[k])
        o4 = ctx.pcn_step(cb, nz[0], nz[1], nz[2], nz[3]).cpu().numpy()
        ctx.check_info("single-GPU step")
        out["single"] = dict(accepted=bool(o4[0]), phi_cur=float(o4[2]), phi_new=float(o4[3]))
        out["phi_rel_diff"] = max(abs(d["phi_cur"] / o4[2] - 1), abs(d["phi_new"] / o4[3] - 1))
        out["accept_equal"] = bool(o4[0]) == bool(d["accepted"])
        out["lu_sample_bit_identical"] = bool(torch.equal(cb.u_new[1], ch.u_new[1]))
        out["qr_sample_bit_identical"] = bool(torch.equal(cb.u_new[2], ch.u_new[2]))
        out["lu_sample_rel"] = rel(ch.u_new[1], cb.u_new[1])
        out["qr_sample_rel"] = rel(ch.u_new[2], cb.u_new[2])
        ctx.close()
    return out


def main():
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
        os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
        dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank()
    cases = [c for c in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["n4", "n8"]) if c]
    res = dict(world=dist.get_world_size(), golden=[golden_replay(c, dev) for c in cases])
    nbig = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["32"]) if x]
    res["single_vs_dist"] = [single_vs_dist(n, dev, rank) for n in nbig]
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
