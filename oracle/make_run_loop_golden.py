"""Golden record of the REFERENCE'S OWN ``Simulation.run`` bookkeeping (run in this container only; needs /root/reference).

    python oracle/make_run_loop_golden.py            # writes tests/golden/ref_run_loop.json

``Simulation.run`` (/root/reference/mcmc/image_cupy.py:949-1009) is host bookkeeping around ``pCN.one_step``: acceptance
counters per evaluation interval, ``pCN.adapt_step`` (528-540) and ``pCN.adapt_step_sqrtBeta`` (546-548) driven by the running
rates, and -- after a ``LinAlgError`` -- the truncation of the sample histories that only happens with ``printProgress``.
None of that depends on what a step computes, so the reference's ``run`` is executed here on a real (n=2) reference
``Simulation`` whose ``pcn.one_step`` is replaced by a scripted sequence of accept flags.  The trajectory of ``beta``,
``betaZ``, ``pcn_step_sqrtBetas`` / ``_Z`` seen at every call, the final counters and the history shapes are stored;
``tests/test_run_loop.py`` drives ``mspde.image.Simulation.run`` with the same script and compares.

Test infrastructure only (like everything under oracle/).
"""
from __future__ import annotations

import io
import json
import os
import sys
import tempfile
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden", "ref_run_loop.json")

# name -> n_samples, evaluation_interval, enable_step_adaptation, printProgress, beta0, fail_at (step index raising LinAlgError)
SCENARIOS = {
    "adapt": dict(n_samples=57, evaluation_interval=10, enable_step_adaptation=True, printProgress=False, beta=0.5, fail_at=None),
    "adapt_short_interval": dict(n_samples=40, evaluation_interval=3, enable_step_adaptation=True, printProgress=True, beta=0.9,
                                 fail_at=None),
    "no_adapt": dict(n_samples=30, evaluation_interval=10, enable_step_adaptation=False, printProgress=False, beta=0.5, fail_at=None),
    "error_quiet": dict(n_samples=40, evaluation_interval=10, enable_step_adaptation=True, printProgress=False, beta=0.5, fail_at=23),
    "error_printing": dict(n_samples=40, evaluation_interval=10, enable_step_adaptation=True, printProgress=True, beta=0.5, fail_at=23),
    "error_first_step": dict(n_samples=12, evaluation_interval=5, enable_step_adaptation=True, printProgress=True, beta=0.5, fail_at=0),
}


def script(i):
    """Scripted accept flags of step i: (layer chain accepted, sqrt-beta move accepted).  Deterministic, irregular."""
    return int((i * 7 + 3) % 11 < 4), int((i * 5 + 1) % 13 < 3)


def run_scenario(im, np, name, sc):
    sim = im.Simulation(n_layers=3, n_samples=sc["n_samples"], n=2, n_extended=64, beta=sc["beta"], kappa=5e9, sigma_0=4e7,
                        sigma_v=3.2e4, sigma_scaling=0.4, meas_std=0.2, evaluation_interval=sc["evaluation_interval"],
                        printProgress=sc["printProgress"], seed=3, burn_percentage=25.0,
                        enable_step_adaptation=sc["enable_step_adaptation"], pcn_variant="dunlop", phantom_name="shepp.png",
                        meas_type="tomo", n_theta=8, verbose=False, hybrid_GPU_CPU=False)
    sim.pcn.target_acceptance_rate = 0.234                 # twoDsim.py:107
    pcn = sim.pcn
    start = dict(beta=float(pcn.beta), betaZ=float(pcn.betaZ), pcn_step_sqrtBetas=float(pcn.pcn_step_sqrtBetas),
                 pcn_step_sqrtBetas_Z=float(pcn.pcn_step_sqrtBetas_Z), beta_feedback_gain=float(pcn.beta_feedback_gain),
                 target_acceptance_rate=float(pcn.target_acceptance_rate), aggresiveness=float(pcn.aggresiveness),
                 record_skip=int(pcn.record_skip))
    seen = []
    calls = [0]

    def one_step(Layers):
        i = calls[0]
        calls[0] += 1
        seen.append([float(pcn.beta), float(pcn.betaZ), float(pcn.pcn_step_sqrtBetas), float(pcn.pcn_step_sqrtBetas_Z)])
        if sc["fail_at"] is not None and i == sc["fail_at"]:
            raise np.linalg.LinAlgError("scripted failure")
        return script(i)

    pcn.one_step = one_step
    rows_before = [int(l.samples_history.shape[0]) for l in sim.Layers]
    with redirect_stdout(io.StringIO()):
        sim.run()
    return dict(scenario=sc, start=start, calls=calls[0], seen=seen,
                final=dict(beta=float(pcn.beta), betaZ=float(pcn.betaZ), pcn_step_sqrtBetas=float(pcn.pcn_step_sqrtBetas),
                           pcn_step_sqrtBetas_Z=float(pcn.pcn_step_sqrtBetas_Z)),
                accepted_count=int(sim.accepted_count), accepted_count_SqrtBeta=int(sim.accepted_count_SqrtBeta),
                acceptancePercentage=(float(sim.acceptancePercentage) if hasattr(sim, "acceptancePercentage") else None),
                has_total_time=hasattr(sim, "total_time"),
                history_rows_before=rows_before, history_rows_after=[int(l.samples_history.shape[0]) for l in sim.Layers])


def main():
    import numpy as np
    sys.path.insert(0, ROOT)
    from oracle import ref_shim
    os.chdir(tempfile.mkdtemp(prefix="refrun_"))
    ns = ref_shim.install(alias64=True, seed=3)
    out = {"source": "/root/reference/mcmc/image_cupy.py:949-1009 (Simulation.run), 528-548 (adapt_step, set_step, "
                     "set_step_sqrtBeta, adapt_step_sqrtBeta), executed unmodified through oracle/ref_shim.py with a scripted one_step",
           "script": "accepted = (7 i + 3) mod 11 < 4 ; accepted_sqrtBeta = (5 i + 1) mod 13 < 3",
           "scenarios": {name: run_scenario(ns.im, np, name, sc) for name, sc in SCENARIOS.items()}}
    with open(OUT, "w") as fh:
        json.dump(out, fh, indent=1)
    for name, r in out["scenarios"].items():
        print(name, "calls", r["calls"], "final", r["final"], "acc", r["accepted_count"], r["accepted_count_SqrtBeta"],
              r["acceptancePercentage"], "rows", r["history_rows_before"], "->", r["history_rows_after"])


if __name__ == "__main__":
    main()
