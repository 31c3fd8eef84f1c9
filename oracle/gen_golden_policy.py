"""Generate tests/golden/policy_traces.json from the REFERENCE's own InternVLAN1Net host logic -- build container only.

    python -m oracle.gen_golden_policy

The reference class (internvla_n1_policy.py, through oracle/ref_loader.load_reference_policy) runs its real
`s2_step` / `step_no_infer` / `s1_step_latent` / `reset` with the scripted processor and language model of
oracle/policy_script.py; recorded per call: the chat text handed to the processor, which frames were attached, the
generate / generate_latents arguments, the parsed result (pixel goal or action list), `llm_output`, and the System-1
action list.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import agent_script, policy_script, ref_loader  # noqa: E402

CASES = [dict(seed=11, steps=40, num_history=8, p_s2=0.45), dict(seed=12, steps=40, num_history=4, p_s2=0.8),
         dict(seed=13, steps=30, num_history=8, p_s2=0.3, reset_at=14)]


def plan_ops(case, answers):
    """Call sequence an agent would make: a look-down answer forces an S2 call with look_down on the next frame."""
    rng = np.random.Generator(np.random.PCG64(case["seed"] + 100))
    ops, n_s2, pending_look = [], 0, False
    for k in range(case["steps"]):
        if k == case.get("reset_at"):
            ops.append(["reset"])
            pending_look = False
        if k == 0 or pending_look or rng.random() < case["p_s2"] or ops[-1] == ["reset"]:
            ops.append(["s2", k, pending_look])
            ans = answers[n_s2 % len(answers)]
            n_s2 += 1
            pending_look = (not any(ch.isdigit() for ch in ans)) and ans.startswith("↓")
            if any(ch.isdigit() for ch in ans) and rng.random() < 0.7:
                ops.append(["s1", k])
        else:
            ops.append(["noinfer", k])
    return ops


def run_case(case):
    rng = np.random.Generator(np.random.PCG64(case["seed"]))
    answers = policy_script.random_answers(rng)
    trajs = policy_script.random_trajs(rng)
    ops = plan_ops(case, answers)
    mod, Net = ref_loader.load_reference_policy()
    proc, llm = policy_script.FakeProcessor(), policy_script.ScriptedLLM(answers, trajs)
    net = Net(llm, proc, num_history=case["num_history"])
    net.reset()
    rec = []
    last_latent = None
    for op in ops:
        if op[0] == "reset":
            net.reset()
            rec.append({"op": op})
        elif op[0] == "noinfer":
            o = agent_script.make_obs(op[1], size=(24, 32))
            net.step_no_infer(o["rgb"], o["depth"], None)
            rec.append({"op": op, "episode_idx": net.episode_idx, "n_rgb": len(net.rgb_list)})
        elif op[0] == "s2":
            o = agent_script.make_obs(op[1], size=(24, 32))
            out = net.s2_step(o["rgb"], o["depth"], None, o["instruction"], None, look_down=op[2])
            last_latent = out.output_latent
            rec.append({"op": op, "processor": proc.log.pop(), "model": llm.log, "llm_output": net.llm_output,
                        "pixel": None if out.output_pixel is None else [int(v) for v in out.output_pixel],
                        "actions": None if out.output_action is None else [int(a) for a in out.output_action],
                        "has_latent": out.output_latent is not None, "episode_idx": net.episode_idx,
                        "n_rgb": len(net.rgb_list), "n_turns": len(net.conversation_history)})
            llm.log = []
        else:
            s1 = net.s1_step_latent(None, None, last_latent)
            rec.append({"op": op, "idx": [int(a) for a in s1.idx], "model": llm.log})
            llm.log = []
    return {"case": case, "answers": answers, "trajs": trajs, "ops": ops, "records": rec}


def main():
    import contextlib
    import io
    traces = []
    for case in CASES:
        with contextlib.redirect_stdout(io.StringIO()):
            traces.append(run_case(case))
        r = traces[-1]["records"]
        print("case", case, "s2 calls", sum(1 for x in r if x["op"][0] == "s2"), "pixel goals",
              sum(1 for x in r if x.get("pixel")), "s1", [x["idx"] for x in r if x["op"][0] == "s1"][:4])
    out = os.path.join(ROOT, "tests", "golden", "policy_traces.json")
    with open(out, "w") as fh:
        json.dump({"traces": traces}, fh, ensure_ascii=False)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
