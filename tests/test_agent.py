"""Agent state machine -- CPU only.

  * oracle/agent_oracle.py vs tests/golden/agent_traces.json: the traces were recorded from the reference's own
    InternVLAN1Agent (oracle/gen_golden_agent.py), so this pins the restatement;
  * internnav_b200.agent.InternVLAN1Agent (batched, no worker thread) vs the same traces with one environment, and with
    all traces running side by side as one batch -- actions, per-step policy calls, dual_forward_step, look_down.
"""
import json
import os

import numpy as np
import pytest

from oracle import agent_script
from oracle.agent_oracle import AgentOracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "agent_traces.json")) as fh:
    TRACES = json.load(fh)["traces"]


def _norm(calls):
    return json.loads(json.dumps(calls))


@pytest.mark.parametrize("ti", range(len(TRACES)))
def test_oracle_matches_reference_trace(ti):
    tr = TRACES[ti]
    case = tr["case"]
    pol = agent_script.ScriptedPolicy(tr["script"])
    ag = AgentOracle(pol, mode=case["mode"], sys2_max_forward_step=case["max_fwd"])
    ag.reset()
    assert _norm(pol.drain()) == tr["steps"][0]["reset_calls"]
    for st in tr["steps"][1:]:
        if st.get("reset"):
            ag.reset(reset_index=[0])
            assert _norm(pol.drain()) == st["calls"]
            continue
        out = ag.step([agent_script.make_obs(st["k"])])
        assert out[0]["action"] == st["action"], (ti, st["k"])
        assert _norm(pol.drain()) == st["calls"], (ti, st["k"])
        assert ag.dual_forward_step == st["dual_forward_step"] and ag.look_down == st["look_down"]


@pytest.mark.parametrize("ti", range(len(TRACES)))
def test_batched_agent_single_env_matches_reference_trace(ti):
    from internnav_b200.agent import InternVLAN1Agent, PerEnvPolicies
    tr = TRACES[ti]
    case = tr["case"]
    pol = agent_script.ScriptedPolicy(tr["script"])
    ag = InternVLAN1Agent(PerEnvPolicies([pol]), num_envs=1, infer_mode=case["mode"],
                          sys2_max_forward_step=case["max_fwd"])
    ag.reset()
    assert _norm(pol.drain()) == tr["steps"][0]["reset_calls"]
    for st in tr["steps"][1:]:
        if st.get("reset"):
            ag.reset([0])
            assert _norm(pol.drain()) == st["calls"]
            continue
        out = ag.step([agent_script.make_obs(st["k"])])
        assert out[0]["action"] == st["action"] and out[0]["ideal_flag"] is True, (ti, st["k"])
        assert _norm(pol.drain()) == st["calls"], (ti, st["k"])
        assert int(ag.dual_forward_step[0]) == st["dual_forward_step"] and bool(ag.look_down[0]) == st["look_down"]


@pytest.mark.parametrize("mode", ["partial_async", "sync"])
def test_batched_agent_runs_traces_side_by_side(mode):
    """All traces of one mode as ONE batch: every environment must behave exactly as it did alone in the reference,
    while System 2 / System 1 are each called at most once per frame for the whole batch."""
    from internnav_b200.agent import InternVLAN1Agent, PerEnvPolicies
    trs = [t for t in TRACES if t["case"]["mode"] == mode and t["case"]["max_fwd"] == 8]
    assert len(trs) >= 2
    pols = [agent_script.ScriptedPolicy(t["script"]) for t in trs]
    ag = InternVLAN1Agent(PerEnvPolicies(pols), num_envs=len(trs), infer_mode=mode, sys2_max_forward_step=8)
    ag.reset()
    cursors = [1] * len(trs)
    for p in pols:
        p.drain()
    n = min(t["case"]["steps"] for t in trs)
    for k in range(n):
        for e, t in enumerate(trs):                      # per-env episode resets at their own frames
            st = t["steps"][cursors[e]]
            if st.get("reset"):
                ag.reset([e])
                assert _norm(pols[e].drain()) == st["calls"]
                cursors[e] += 1
        before = dict(ag.calls)
        outs = ag.step([agent_script.make_obs(k) for _ in trs])
        assert ag.calls["s2"] - before["s2"] <= 1 and ag.calls["s1"] - before["s1"] <= 1
        for e, t in enumerate(trs):
            st = t["steps"][cursors[e]]
            cursors[e] += 1
            assert st["k"] == k
            assert outs[e]["action"] == st["action"], (e, k)
            assert _norm(pols[e].drain()) == st["calls"], (e, k)
            assert int(ag.dual_forward_step[e]) == st["dual_forward_step"]
            assert bool(ag.look_down[e]) == st["look_down"]
    assert ag.calls["s2_envs"] > ag.calls["s2"] or len(trs) == 1


def test_invalid_mode_and_missing_plan():
    from internnav_b200.agent import InternVLAN1Agent, PerEnvPolicies
    with pytest.raises(ValueError):
        InternVLAN1Agent(PerEnvPolicies([]), num_envs=0, infer_mode="async")
    pol = agent_script.ScriptedPolicy({"s2": [{"actions": []}], "s1": [[1]]})
    ag = InternVLAN1Agent(PerEnvPolicies([pol]), num_envs=1, infer_mode="sync")
    ag.reset()
    ag.plans[0].actions = None
    with pytest.raises((AssertionError, IndexError)):
        ag.step([agent_script.make_obs(0)])


def test_batched_agent_equals_oracle_on_random_scripts():
    """Beyond the recorded traces: 60 random scripts (both modes, three System-2 budgets, random resets and failures),
    batched agent with 5 environments per run vs five independent single-environment oracles -- the oracle being pinned to
    the reference class by the traces above."""
    from internnav_b200.agent import InternVLAN1Agent, PerEnvPolicies
    rng = np.random.Generator(np.random.PCG64(2024))
    for trial in range(12):
        mode = "partial_async" if trial % 3 else "sync"
        max_fwd = int(rng.choice([4, 8, 12]))
        B, steps = 5, 45
        scripts = [agent_script.random_script(rng, p_latent=float(rng.uniform(0.2, 0.9)), p_raise=float(rng.uniform(0, 0.12)))
                   for _ in range(B)]
        resets = {int(rng.integers(5, steps)): int(rng.integers(0, B)) for _ in range(2)}
        pols_a = [agent_script.ScriptedPolicy(s) for s in scripts]
        pols_b = [agent_script.ScriptedPolicy(s) for s in scripts]
        ag = InternVLAN1Agent(PerEnvPolicies(pols_a), num_envs=B, infer_mode=mode, sys2_max_forward_step=max_fwd)
        oracles = [AgentOracle(p, mode=mode, sys2_max_forward_step=max_fwd) for p in pols_b]
        ag.reset()
        for o in oracles:
            o.reset()
        for k in range(steps):
            if k in resets:
                ag.reset([resets[k]])
                oracles[resets[k]].reset(reset_index=[0])
            obs = [agent_script.make_obs(k + 3 * e) for e in range(B)]
            out = ag.step(obs)
            for e in range(B):
                ref = oracles[e].step([obs[e]])
                assert out[e]["action"] == ref[0]["action"], (trial, k, e)
                assert _norm(pols_a[e].drain()) == _norm(pols_b[e].drain()), (trial, k, e)
                assert int(ag.dual_forward_step[e]) == oracles[e].dual_forward_step
                assert bool(ag.look_down[e]) == oracles[e].look_down
