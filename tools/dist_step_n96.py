"""torchrun tool: steps of the block-distributed sampler (mspde.dist.DistPCN, default n=96) with per-step timings.

  python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/dist_step_n96.py [n] [steps]"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mspde.dist import build_dist_chain
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
chain, noise_fn = build_dist_chain(n, max(n, 64), 90, 3, dev, seed=7, beta=0.3)
for rep in range(steps):
    nz = noise_fn()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    d = chain.one_step(*nz)
    torch.cuda.synchronize(); dist.barrier(); dt = time.perf_counter() - t0
    if dist.get_rank() == 0:
        print(json.dumps(dict(n=n, N=chain.N, M=chain.M, world=dist.get_world_size(), step=rep, seconds=dt, **d)), flush=True)
dist.destroy_process_group()
