"""torchrun tool: one block-distributed pCN iteration's worth of dense work (default n=96) -- see mspde.dist.DistStep."""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200")); sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mspde.dist import DistStep
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 96
stp = DistStep(n, max(n, 64), 90, dev)
for rep in range(2):
    r = stp.run_overlapped() if (len(sys.argv) > 2 and sys.argv[2] == "overlap") else stp.run_once()
    if dist.get_rank() == 0:
        print(json.dumps(dict(n=n, N=stp.N, M=stp.M, world=dist.get_world_size(), rep=rep, **r)), flush=True)
dist.destroy_process_group()
