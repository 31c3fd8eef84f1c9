"""Scripted stand-in for the dual-system policy (InternVLAN1Net) -- TEST INFRASTRUCTURE.

Drives an agent state machine (the reference's InternVLAN1Agent in oracle/gen_golden_agent.py, the restatement in
oracle/agent_oracle.py, the batched scheduler in internnav_b200/agent.py) with pre-scripted System-2 / System-1 results
and logs every policy call, so that action sequences AND call sequences can be compared.

Script format (JSON-friendly):
  s2: list of items consumed one per `s2_step` call (cyclically):
        {"actions": [..]}            discrete answer (0 STOP, 1 fwd, 2 left, 3 right, 5 look-down)
        {"latent": n, "pixel": [y, x]}  pixel-goal answer; the latent is the opaque tag n
        {"raise": true}              the call raises
  s1: list of action-id lists consumed one per `s1_step_latent` call (cyclically); [] means "no action".
Observations are synthetic: frame k has rgb == k % 256 everywhere and depth == (k % 64) / 100.
"""
from types import SimpleNamespace

import numpy as np


def make_obs(k, size=8):
    h, w = (size, size) if isinstance(size, int) else size
    rgb = np.full((h, w, 3), k % 256, dtype=np.uint8)
    depth = np.full((h, w, 1), (k % 64) / 100.0, dtype=np.float32)
    return {"rgb": rgb, "depth": depth, "instruction": "go to frame %d" % k}


def frame_of(rgb):
    """Frame tag of an rgb array / tensor in either the raw uint8 form or the /255 float form."""
    v = float(np.asarray(rgb, dtype=np.float64).reshape(-1)[0])
    return int(round(v * 255)) if v <= 1.0 and np.asarray(rgb).dtype != np.uint8 else int(round(v))


class ScriptedPolicy:
    """Duck-types the four methods the agent calls on the policy (internvla_n1_agent.py L160, L178, L240, L335)."""

    def __init__(self, script, s2_output_cls=None, s1_output_cls=None):
        self.script = script
        self.n_s2 = self.n_s1 = 0
        self.log = []
        self._s2_cls = s2_output_cls or (lambda: SimpleNamespace(output_action=None, output_pixel=None, output_latent=None))
        self._s1_cls = s1_output_cls or (lambda **kw: SimpleNamespace(**kw))

    def eval(self):
        return self

    def reset(self):
        self.log.append(["reset"])

    def step_no_infer(self, rgb, depth, pose):
        self.log.append(["noinfer", frame_of(rgb)])

    def s2_step(self, rgb, depth, pose, instruction, intrinsic, look_down=False):
        item = self.script["s2"][self.n_s2 % len(self.script["s2"])]
        self.n_s2 += 1
        self.log.append(["s2", frame_of(rgb), bool(look_down), instruction])
        if item.get("raise"):
            raise RuntimeError("scripted System-2 failure")
        out = self._s2_cls()
        if "actions" in item:
            out.output_action = list(item["actions"])
        else:
            out.output_pixel = np.array(item["pixel"])
            out.output_latent = ("latent", item["latent"])
        return out

    def s1_step_latent(self, rgb, depth, latent):
        idx = list(self.script["s1"][self.n_s1 % len(self.script["s1"])])
        self.n_s1 += 1
        r = np.asarray(rgb, dtype=np.float64)
        d = np.asarray(depth, dtype=np.float64)
        if r.ndim == 5:      # partial_async: [1, 2, 224, 224, 3] floats, [goal frame, current frame]
            frames = [int(round(r[0, 0, 0, 0, 0] * 255)), int(round(r[0, 1, 0, 0, 0] * 255))]
            depths = [round(float(d[0, 0, 0, 0, 0]), 4), round(float(d[0, 1, 0, 0, 0]), 4)]
        else:                # sync: the raw frame, depth * 10000
            frames = [frame_of(rgb)]
            depths = [round(float(d.reshape(-1)[0]), 4)]
        self.log.append(["s1", frames, depths, latent[1]])
        return self._s1_cls(idx=idx)

    def drain(self):
        out, self.log = self.log, []
        return out


def random_script(rng, n_s2=24, n_s1=40, p_latent=0.6, p_raise=0.06):
    s2 = []
    for i in range(n_s2):
        u = rng.random()
        if i > 0 and u < p_raise:
            s2.append({"raise": True})
        elif u < p_raise + p_latent:
            s2.append({"latent": int(rng.integers(0, 1000)), "pixel": [int(rng.integers(0, 480)), int(rng.integers(0, 640))]})
        else:
            kind = rng.random()
            if kind < 0.2:
                s2.append({"actions": [5]})
            elif kind < 0.3:
                s2.append({"actions": [0]})
            else:
                s2.append({"actions": [int(a) for a in rng.choice([1, 2, 3], size=int(rng.integers(1, 5)))]})
    s1 = []
    for _ in range(n_s1):
        n = int(rng.choice([0, 1, 2, 3, 4, 4, 4, 4]))
        s1.append([int(a) for a in rng.choice([1, 2, 3], size=n)])
    return {"s2": s2, "s1": s1}
