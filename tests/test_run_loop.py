"""mspde.image.Simulation.run / pCN step adaptation against the REFERENCE'S OWN run loop (CPU, no device).

tests/golden/ref_run_loop.json was recorded by oracle/make_run_loop_golden.py from the unmodified reference
(/root/reference/mcmc/image_cupy.py:949-1009, 528-548) with pcn.one_step replaced by a scripted sequence of accept flags.
Here the product's run() -- the real method, on objects built without a device -- is driven by the same script: the step
sizes it presents to every step, the counters, the acceptance rate and the history truncation must equal the reference's."""
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "mcmc-multispde_b200"))

with open(os.path.join(ROOT, "tests", "golden", "ref_run_loop.json")) as _fh:
    GOLD = json.load(_fh)["scenarios"]


def script(i):
    return int((i * 7 + 3) % 11 < 4), int((i * 5 + 1) % 13 < 3)


class _Lay:
    def __init__(self, rows):
        self.samples_history = np.zeros((rows, 3), dtype=np.complex64)


def _bare_simulation(g):
    """Simulation and pCN objects carrying only what run() and the adaptation methods read (no context, no device)."""
    from mspde import image as im
    sc, st = g["scenario"], g["start"]
    pcn = object.__new__(im.pCN)
    pcn.beta, pcn.betaZ = st["beta"], st["betaZ"]
    pcn.pcn_step_sqrtBetas, pcn.pcn_step_sqrtBetas_Z = st["pcn_step_sqrtBetas"], st["pcn_step_sqrtBetas_Z"]
    pcn.beta_feedback_gain, pcn.target_acceptance_rate = st["beta_feedback_gain"], st["target_acceptance_rate"]
    pcn.aggresiveness, pcn.record_skip = st["aggresiveness"], st["record_skip"]
    sim = object.__new__(im.Simulation)
    sim.pcn = pcn
    sim.n_samples, sim.evaluation_interval = sc["n_samples"], sc["evaluation_interval"]
    sim.enable_step_adaptation, sim.printProgress = sc["enable_step_adaptation"], sc["printProgress"]
    sim.n_layers = 3
    sim.Layers = [_Lay(r) for r in g["history_rows_before"]]
    sim.acceptancePercentage = 0.0            # Simulation.__init__ (image_cupy.py:862)
    return sim


@pytest.mark.parametrize("name", sorted(GOLD))
def test_run_bookkeeping_matches_the_reference(name, capsys):
    g = GOLD[name]
    sim = _bare_simulation(g)
    pcn = sim.pcn
    seen, calls = [], [0]

    def one_step(Layers):
        i = calls[0]
        calls[0] += 1
        seen.append([pcn.beta, pcn.betaZ, pcn.pcn_step_sqrtBetas, pcn.pcn_step_sqrtBetas_Z])
        if g["scenario"]["fail_at"] is not None and i == g["scenario"]["fail_at"]:
            raise np.linalg.LinAlgError("scripted failure")
        return script(i)

    pcn.one_step = one_step
    sim.run()
    capsys.readouterr()
    assert calls[0] == g["calls"]
    np.testing.assert_allclose(np.array(seen), np.array(g["seen"]), rtol=1e-14, atol=0)
    for k, v in g["final"].items():
        assert getattr(pcn, k) == pytest.approx(v, rel=1e-14), k
    assert sim.accepted_count == g["accepted_count"]
    assert sim.accepted_count_SqrtBeta == g["accepted_count_SqrtBeta"]
    assert sim.acceptancePercentage == pytest.approx(g["acceptancePercentage"], rel=1e-15)
    assert hasattr(sim, "total_time") == g["has_total_time"]
    assert [l.samples_history.shape[0] for l in sim.Layers] == g["history_rows_after"]


def test_step_setters_follow_the_reference_guards():
    """set_step only accepts 1e-7 < beta < 1 (image_cupy.py:536-540); set_step_sqrtBeta clips at 1 (542-544); the
    aggressiveness helpers (531-534) keep their min() quirks."""
    from mspde import image as im
    pcn = object.__new__(im.pCN)
    pcn.beta, pcn.betaZ, pcn.aggresiveness = 0.5, np.sqrt(0.75), 0.2
    pcn.set_step(1.0)
    assert pcn.beta == 0.5
    pcn.set_step(1e-8)
    assert pcn.beta == 0.5
    pcn.set_step(0.25)
    assert pcn.beta == 0.25 and pcn.betaZ == pytest.approx(np.sqrt(1 - 0.0625), rel=1e-15)
    pcn.more_aggresive()
    assert pcn.beta == pytest.approx(0.3, rel=1e-15)
    pcn.less_aggresive()                       # min((1 - a) beta, 1e-10) = 1e-10: rejected by set_step, beta unchanged (reference quirk)
    assert pcn.beta == pytest.approx(0.3, rel=1e-15)
    pcn.set_step_sqrtBeta(1.7)
    assert pcn.pcn_step_sqrtBetas == 1.0 and pcn.pcn_step_sqrtBetas_Z == 0.0
    pcn.set_step_sqrtBeta(0.6)
    assert pcn.pcn_step_sqrtBetas_Z == pytest.approx(0.8, rel=1e-15)


class _FakeChain:
    """Stands in for mspde.dist.DistPCN on CPU: scripted accept flags, a state that changes when a step is accepted."""

    def __init__(self, g, K=3, nr=5):
        import torch
        self.K, self.nr = K, nr
        self.sqrt_betas = [3.0, 2.0, 1.0]
        self.beta, self.betaZ = g["start"]["beta"], g["start"]["betaZ"]
        self.fail_at = g["scenario"]["fail_at"]
        self.calls, self.seen = 0, []
        self.state = torch.zeros((K, nr), dtype=torch.complex128)

    def one_step(self, w_half, log_u, w_last, e):
        i = self.calls
        self.calls += 1
        self.seen.append([self.beta, self.betaZ])
        if self.fail_at is not None and i == self.fail_at:
            raise np.linalg.LinAlgError("scripted failure")
        acc = script(i)[0]
        if acc:
            self.state = self.state + (i + 1)
        return dict(accepted=bool(acc), log_ratio=0.0, phi_cur=0.0, phi_new=0.0)

    def current_halves(self):
        return self.state.clone()


@pytest.mark.parametrize("name", sorted(GOLD))
def test_distributed_simulation_surface_runs_the_same_loop(name, capsys):
    """mspde.dist.DistSimulation (Simulation(..., distributed=True)): run() is Simulation.run, the step rules are pCN's, and
    they act on the chain's own beta / betaZ.  Driven by a stand-in chain with the reference golden's accept script, the step
    sizes the CHAIN sees at every step equal the reference's; every step's current half-samples land in the histories."""
    from mspde import dist as md
    g = GOLD[name]
    sc = g["scenario"]
    chain = _FakeChain(g)
    sim = md.DistSimulation.from_chain(chain, lambda: (None, 0.0, None, None), sc["n_samples"], sc["evaluation_interval"],
                                       sc["printProgress"], 25.0, sc["enable_step_adaptation"])
    sim.pcn.target_acceptance_rate = g["start"]["target_acceptance_rate"]
    sim.run()
    capsys.readouterr()
    assert chain.calls == g["calls"]
    np.testing.assert_allclose(np.array(chain.seen), np.array(g["seen"])[:, :2], rtol=1e-14, atol=0)
    assert chain.beta == pytest.approx(g["final"]["beta"], rel=1e-14) and chain.betaZ == pytest.approx(g["final"]["betaZ"], rel=1e-14)
    assert sim.accepted_count == g["accepted_count"] and sim.accepted_count_SqrtBeta == 0
    assert sim.acceptancePercentage == pytest.approx(g["acceptancePercentage"], rel=1e-15)
    assert [l.samples_history.shape[0] for l in sim.Layers] == g["history_rows_after"]
    if sc["fail_at"] is None:
        # row r holds the state after step r: the sum of (i + 1) over the accepted steps i <= r
        want = np.cumsum([(i + 1) * script(i)[0] for i in range(sc["n_samples"])])
        for l in sim.Layers:
            assert l.i_record == sc["n_samples"]
            np.testing.assert_array_equal(l.samples_history[:, 0].real, want)
            assert np.all(l.sqrt_beta_history == l.sqrt_beta)


def test_distributed_simulation_result_file(tmp_path):
    from mspde import dist as md
    g = GOLD["no_adapt"]
    sim = md.DistSimulation.from_chain(_FakeChain(g), lambda: (None, 0.0, None, None), 12, 5, False, 25.0, False)
    sim.run()
    path = str(tmp_path / "result_0.npz")
    sim.save(path)
    z = np.load(path)
    assert int(z["n_layers"]) == 3 and z["Layers 2/samples_history"].shape == (12, 5)
    assert float(z["pcn/beta"]) == 0.5 and int(z["accepted_count"]) == sim.accepted_count
