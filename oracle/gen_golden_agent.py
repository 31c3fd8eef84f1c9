"""Generate tests/golden/agent_traces.json from the REFERENCE's own InternVLAN1Agent -- build container only.

    python -m oracle.gen_golden_agent

The reference class (internnav/agent/internvla_n1_agent.py, imported untouched through oracle/ref_loader.py) is driven
by oracle/agent_script.ScriptedPolicy for a number of seeded scripts in both `infer_mode`s; every step's returned action
and every policy call it made (with which frame, look_down flag, goal/current frame pair and latent tag) is recorded.
The S2 worker thread of the reference runs for real; only its polling sleeps are shortened.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import agent_script, ref_loader  # noqa: E402

CASES = [
    dict(seed=1, mode="partial_async", max_fwd=8, steps=70, resets=[]),
    dict(seed=2, mode="partial_async", max_fwd=8, steps=70, resets=[33]),
    dict(seed=3, mode="partial_async", max_fwd=4, steps=60, resets=[]),
    dict(seed=4, mode="sync", max_fwd=8, steps=60, resets=[25]),
    dict(seed=5, mode="sync", max_fwd=8, steps=50, resets=[]),
    dict(seed=6, mode="partial_async", max_fwd=8, steps=60, resets=[10, 11], p_latent=0.9, p_raise=0.0),
    dict(seed=7, mode="partial_async", max_fwd=8, steps=60, resets=[], p_latent=0.3, p_raise=0.15),
]


def run_case(case):
    rng = np.random.Generator(np.random.PCG64(case["seed"]))
    script = agent_script.random_script(rng, p_latent=case.get("p_latent", 0.6), p_raise=case.get("p_raise", 0.06))
    holder = {}

    def factory(config=None):  # the reference's own S2Output / S1Output dataclasses carry the scripted results
        holder["policy"] = agent_script.ScriptedPolicy(script, s2_output_cls=holder["mod"].S2Output,
                                                       s1_output_cls=holder["mod"].S1Output)
        return holder["policy"]

    mod = holder["mod"] = ref_loader.load_reference_agent(factory)
    AgentCfg = mod.AgentCfg
    settings = dict(policy_name="InternVLAN1_Policy", state_encoder=None, device="cpu", infer_mode=case["mode"],
                    sys2_max_forward_step=case["max_fwd"], width=640, height=480, hfov=79, vis_debug=False)
    agent = mod.InternVLAN1Agent(AgentCfg(model_name="internvla_n1", model_settings=settings))
    policy = holder["policy"]
    agent.reset()
    steps = [{"reset_calls": policy.drain()}]
    for k in range(case["steps"]):
        if k in case["resets"]:
            agent.reset(reset_index=[0])
            steps.append({"reset": True, "calls": policy.drain()})
        out = agent.step([agent_script.make_obs(k)])
        steps.append({"k": k, "action": [int(a) for a in out[0]["action"]], "ideal_flag": bool(out[0]["ideal_flag"]),
                      "calls": policy.drain(), "dual_forward_step": int(agent.dual_forward_step),
                      "look_down": bool(agent.look_down)})
    return {"case": case, "script": script, "steps": steps}


def main():
    import contextlib
    import io
    traces = []
    for case in CASES:
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):  # the reference prints every step
            traces.append(run_case(case))
        acts = [s["action"][0] for s in traces[-1]["steps"] if "action" in s]
        print("case", case, "actions", acts)
    out = os.path.join(ROOT, "tests", "golden", "agent_traces.json")
    with open(out, "w") as fh:
        json.dump({"traces": traces}, fh)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
