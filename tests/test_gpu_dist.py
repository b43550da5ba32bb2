"""GPU tests (-m gpu) of the block-distributed sampler ``mspde.dist.DistPCN`` (BASELINE config 5).

Single rank: always.  Two ranks (torchrun, NCCL): when the box has >= 2 GPUs (the tool the test launches is also what
``gpurun --gpus 2`` runs for profiles/r02_dist_chain_check_*.log)."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TOOL = os.path.join(ROOT, "tools", "dist_chain_check.py")


def _run(nproc, cases, nbig, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), TOOL, cases, nbig]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=1500)
    assert p.returncode == 0, p.stderr[-3000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def _check(res, world):
    assert res["world"] == world
    for g in res["golden"]:                       # vs vectors computed by the reference's own classes
        assert g["accept_equal"] and g["phi_rel"] <= 1e-10 and g["u_rel"] <= 1e-10 and g["noise_rel"] <= 1e-13, g
    for s in res["single_vs_dist"]:               # vs the single-GPU kernels on the same inputs
        assert s["accept_equal"] and s["phi_rel_diff"] <= 1e-11, s
        assert s["lu_sample_rel"] <= 1e-13, s     # same panel kernels (measured 6e-18: rounding-level, not bit-identical)
        assert s["qr_sample_rel"] <= 1e-10, s


def test_dist_sampler_single_rank_vs_reference_golden_and_single_gpu():
    _check(_run(1, "n4,n4k2,n8", "16", 29541), 1)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_dist_sampler_two_ranks_vs_reference_golden_and_single_gpu():
    _check(_run(2, "n4,n4k4,n8", "16,32", 29543), 2)
