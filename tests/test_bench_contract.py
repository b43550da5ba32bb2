"""bench.py contract checks that need no GPU: the CPU reference arm prints exactly one JSON line with the keys the driver
reads, and the helper formulas (flop count, workload labels) are self-consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_flop_count_matches_survey_figures():
    import bench
    f, N, M = bench.flops_iter(**bench.CONFIGS[2])
    assert (N, M) == (3969, 11430)
    assert abs(f - 1.218e13) / 1.218e13 < 2e-3                      # F_iter at n=32, 3 layers
    f5, N5, M5 = bench.flops_iter(**bench.CONFIGS[5])
    assert (N5, M5) == (36481, 17190) and abs(f5 - 1.372e15) / 1.372e15 < 2e-3
    for key, idx in ((1, 0), (2, 1), (3, 2), (5, 4)):               # labels point at the right BASELINE.json config
        assert f"configs[{idx}]" in bench.workload_config(bench.CONFIGS[key], 1)["workload"]


def test_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the oracle port on the host cores) on the smallest config; OMP_NUM_THREADS=1 mimics what
    torchrun exports -- the arm must still use every core it may run on."""
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "1", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f64"
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else os.cpu_count()
    assert d["cpu_baseline"]["cores"] == avail or avail == 1


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_config5_memory_guard_is_calibrated_on_the_observed_peak():
    """bench.config5_bytes_per_rank: the collective free-HBM check before the n=96 block.  At the measured point (8 ranks) it
    asks for the observed allocator peak plus fragmentation headroom and stays below what a B200 has free; fewer ranks need
    more per rank, a smaller basis less."""
    import bench
    at8 = bench.config5_bytes_per_rank(36481, 17190, 8)
    assert 126.5 * 2**30 < at8 < 165 * 2**30
    assert bench.config5_bytes_per_rank(36481, 17190, 4) > at8 > bench.config5_bytes_per_rank(36481, 17190, 16)
    assert bench.config5_bytes_per_rank(16129, 11430, 8) < 0.3 * at8


def _run_guard_script(body):
    code = ("import os, sys, json\nsys.path.insert(0, %r)\nimport bench\nbench.reserve_stdout()\n"
            "g = bench.LineGuard({'metric': 'm', 'value': 1.5})\n" % ROOT) + body
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)


def test_line_guard_prints_the_fallback_when_the_process_dies():
    """bench.LineGuard: if rank 0 is killed inside a collective block (NCCL watchdog abort), the fallback line still reaches the
    result descriptor; on the normal path (disarm, then emit) stdout carries exactly the one real line."""
    died = _run_guard_script("os.abort()\n")
    assert died.returncode != 0
    lines = [l for l in died.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0]) == {"metric": "m", "value": 1.5}
    ok = _run_guard_script("g.disarm()\nbench.emit({'metric': 'm', 'value': 2.5})\n")
    assert ok.returncode == 0, ok.stderr
    lines = [l for l in ok.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1 and json.loads(lines[0])["value"] == 2.5
