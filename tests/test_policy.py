"""Host logic of the batched policy (prompt building, history selection, answer parsing, System-1 action lists) against
tests/golden/policy_traces.json, which oracle/gen_golden_policy.py recorded from the reference's own InternVLAN1Net
methods driven by the scripted processor / language model of oracle/policy_script.py.  CPU only."""
import json
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

from oracle import agent_script, policy_script

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "policy_traces.json"), encoding="utf-8") as fh:
    TRACES = json.load(fh)["traces"]


class BatchedScriptedModel:
    """generate_with_latents / generate_traj of the n1b200 model mirror, answering from per-environment scripts."""

    def __init__(self, traces):
        self.answers = [t["answers"] for t in traces]
        self.trajs = [t["trajs"] for t in traces]
        self.n_ans = [0] * len(traces)
        self.n_traj = [0] * len(traces)
        self.calls = []
        self.order = None  # environments of the call in flight (set by the test: the model itself is env-agnostic)

    def generate_with_latents(self, prompts, pixel_values, image_grid_thw, max_new_tokens=128):
        gen = []
        for e in self.order:
            ans = self.answers[e][self.n_ans[e] % len(self.answers[e])]
            self.n_ans[e] += 1
            gen.append(policy_script.encode(ans) + [151645])
        self.calls.append({"prompts": [len(p) for p in prompts], "pixels": int(pixel_values.shape[0]),
                           "grids": image_grid_thw.tolist(), "max_new_tokens": max_new_tokens})
        lat = torch.arange(len(prompts), dtype=torch.float32).reshape(-1, 1, 1)
        return SimpleNamespace(generated=gen, latents=lat, sequences=None, decode_passes=0)

    def generate_traj(self, traj_latents=None, images_dp=None, depths_dp=None):
        out = []
        for e in self.order:
            out.append(torch.tensor(self.trajs[e][self.n_traj[e] % len(self.trajs[e])], dtype=torch.float32))
            self.n_traj[e] += 1
        assert traj_latents.shape[0] == len(self.order) == images_dp.shape[0]
        return torch.cat(out, dim=0)   # env-major [B * Ns, T, 3]


def _run(trace_ids):
    from internnav_b200.policy import InternVLAN1Policy
    traces = [TRACES[i] for i in trace_ids]
    model = BatchedScriptedModel(traces)
    proc = policy_script.FakeProcessor()
    assert len({t["case"]["num_history"] for t in traces}) == 1
    pol = InternVLAN1Policy(model, proc, num_envs=len(traces), num_history=traces[0]["case"]["num_history"])
    pol.reset()
    last_latent = [None] * len(traces)
    rounds = max(len(t["ops"]) for t in traces)
    n_calls = 0
    for r in range(rounds):
        live = [(e, t["ops"][r], t["records"][r]) for e, t in enumerate(traces) if r < len(t["ops"])]
        for e, op, _ in live:
            if op[0] == "reset":
                pol.reset([e])
        ni = [(e, op) for e, op, _ in live if op[0] == "noinfer"]
        if ni:
            pol.step_no_infer([e for e, _ in ni], [agent_script.make_obs(op[1], size=(24, 32))["rgb"] for _, op in ni])
        s2 = [(e, op, rec) for e, op, rec in live if op[0] == "s2"]
        if s2:
            envs = [e for e, _, _ in s2]
            obs = [agent_script.make_obs(op[1], size=(24, 32)) for _, op, _ in s2]
            model.order = envs
            before = len(proc.log)
            res = pol.s2_step(envs, [o["rgb"] for o in obs], [o["depth"] for o in obs], [None] * len(envs),
                              [o["instruction"] for o in obs], None, [op[2] for _, op, _ in s2])
            n_calls += 1
            call = model.calls[-1]
            plog = proc.log[before:]
            assert len(plog) == len(envs)
            grid_cursor = 0
            for j, (e, op, rec) in enumerate(s2):
                out, ep = res[j], pol.episodes[e]
                assert not isinstance(out, Exception), out
                assert plog[j] == rec["processor"], (e, r)                        # chat text, attached frames, sizes
                assert call["prompts"][j] == rec["model"][0][1]                   # prompt length given to generate
                assert ep.llm_output == rec["llm_output"]
                assert (None if out.output_pixel is None else out.output_pixel.tolist()) == rec["pixel"]
                assert out.output_action == rec["actions"]
                assert (out.output_latent is not None) == rec["has_latent"]
                assert (ep.episode_idx, len(ep.rgb_list), len(ep.conversation_history)) == \
                    (rec["episode_idx"], rec["n_rgb"], rec["n_turns"])
                n_img = len(rec["processor"]["images"])
                mine = [v for g in call["grids"][grid_cursor:grid_cursor + n_img] for v in g]
                grid_cursor += n_img
                if rec["has_latent"]:                                              # same images reach the latent plan
                    assert mine == rec["model"][1][3]
                    last_latent[e] = out.output_latent
            assert grid_cursor == len(call["grids"])
        s1 = [(e, op, rec) for e, op, rec in live if op[0] == "s1"]
        if s1:
            envs = [e for e, _, _ in s1]
            model.order = envs
            outs = pol.s1_step_latent(envs, [torch.zeros(1, 2, 4, 4, 3)] * len(envs), [torch.zeros(1, 2, 4, 4, 1)] * len(envs),
                                      [last_latent[e] for e in envs])
            for (e, _, rec), o in zip(s1, outs):
                assert o.idx == rec["idx"], (e, r)
        for e, op, rec in live:
            if op[0] == "noinfer":
                ep = pol.episodes[e]
                assert (ep.episode_idx, len(ep.rgb_list)) == (rec["episode_idx"], rec["n_rgb"])
    return n_calls


@pytest.mark.parametrize("ti", range(len(TRACES)))
def test_policy_matches_reference_trace(ti):
    _run([ti])


def test_policy_batched_environments():
    ids = [i for i, t in enumerate(TRACES) if t["case"]["num_history"] == 8]
    assert len(ids) >= 2
    n_calls = _run(ids)
    assert n_calls < sum(sum(1 for op in TRACES[i]["ops"] if op[0] == "s2") for i in ids)  # calls were shared


def test_parse_helpers():
    from internnav_b200.policy import parse_actions, split_and_clean
    assert parse_actions("↑↑←STOP→↓") == [1, 1, 2, 0, 3, 5]
    assert parse_actions("no action here") == []
    assert split_and_clean(" a\n<image>\n b <image>") == ["a", "<image>", "b", "<image>"]


def test_look_down_needs_a_previous_answer():
    from internnav_b200.policy import InternVLAN1Policy
    pol = InternVLAN1Policy(BatchedScriptedModel([TRACES[0]]), policy_script.FakeProcessor(), num_envs=1)
    o = agent_script.make_obs(0, size=(24, 32))
    res = pol.s2_step([0], [o["rgb"]], [o["depth"]], [None], [o["instruction"]], None, [True])
    assert isinstance(res[0], AssertionError)


def test_malformed_pixel_answer_is_reported_per_environment():
    """An answer with digits but fewer than two numbers raises IndexError inside the reference's s2_step (L181); here
    the environment that produced it gets the exception, its batch neighbours are unaffected, and the agent's retry
    rule (reset, retry without look_down, then STOP) applies to it alone."""
    from internnav_b200.agent import InternVLAN1Agent
    from internnav_b200.policy import InternVLAN1Policy
    traces = [dict(answers=["7"], trajs=TRACES[0]["trajs"]), dict(answers=["↑→"], trajs=TRACES[0]["trajs"])]
    model = BatchedScriptedModel(traces)
    pol = InternVLAN1Policy(model, policy_script.FakeProcessor(), num_envs=2)
    obs = [agent_script.make_obs(0, size=(24, 32)) for _ in range(2)]
    model.order = [0, 1]
    res = pol.s2_step([0, 1], [o["rgb"] for o in obs], [o["depth"] for o in obs], [None, None],
                      [o["instruction"] for o in obs], None, [False, False])
    assert isinstance(res[0], IndexError) and res[1].output_action == [1, 3]

    class Ordered:  # lets the scripted model know which environments a call serves
        def __init__(self, pol):
            self.pol = pol

        def __getattr__(self, k):
            return getattr(self.pol, k)

        def s2_step(self, env_ids, *a):
            model.order = list(env_ids)
            return self.pol.s2_step(env_ids, *a)
    pol2 = InternVLAN1Policy(model, policy_script.FakeProcessor(), num_envs=2)
    ag = InternVLAN1Agent(Ordered(pol2), num_envs=2, infer_mode="sync")
    ag.reset()
    out = ag.step(obs)
    assert [o["action"] for o in out] == [[0], [1]]      # env 0: two failures -> STOP; env 1: first arrow
