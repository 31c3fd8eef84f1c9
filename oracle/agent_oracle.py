"""Single-environment restatement of the reference agent's decision logic -- TEST INFRASTRUCTURE, not product code.

Follows internnav/agent/internvla_n1_agent.py: `reset` L87-117, the S2 worker `s2_thread_func` L134-204 (collapsed into
a synchronous call: the reference's main thread blocks on `is_infering` / `validate()` L270-274 until the worker has
published its result, so the observable behaviour is that of a call), `should_infer_s2` L210-241 and `step` L243-407.
Pinned by tests/test_agent.py against tests/golden/agent_traces.json, which oracle/gen_golden_agent.py recorded from
the reference class itself.
"""
import copy

import numpy as np
import torch
from PIL import Image


class _S2Out:
    def __init__(self):
        self.idx = -1
        self.output_action = None
        self.output_pixel = None
        self.output_latent = None
        self.rgb_memory = None
        self.depth_memory = None


class AgentOracle:
    def __init__(self, policy, mode="sync", sys2_max_forward_step=8, width=640, height=480, hfov=79):
        self.policy = policy
        self.mode = mode
        self.sys2_max_forward_step = sys2_max_forward_step
        fx = (width / 2.0) / np.tan(np.deg2rad(hfov / 2.0))                      # L119-131
        self.camera_intrinsic = np.array([[fx, 0.0, (width - 1.0) / 2.0, 0.0], [0.0, fx, (height - 1.0) / 2.0, 0.0],
                                          [0.0, 0.0, 1.0, 0.0], [0.0, 0.0, 0.0, 1.0]])
        self.episode_step = 0
        self.episode_idx = 0
        self.look_down = False
        self.dual_forward_step = 0
        self.sys1_infer_times = 0
        self.sys1_depth_threshold = 5.0
        self.sys1_forward_step = 4
        self.s2_output = _S2Out()

    def reset(self, reset_index=None):                                           # L87-117
        self.episode_idx = self.episode_idx + 1 if reset_index is not None else -1
        self.episode_step = 0
        self.s2_output = _S2Out()
        self.dual_forward_step = 0
        self.sys1_infer_times = 0
        self.policy.reset()

    def _run_s2(self, rgb, depth, pose, instruction, look_down, idx):            # L134-204
        try:
            cur = self.policy.s2_step(rgb, depth, pose, instruction, self.camera_intrinsic, look_down)
        except Exception:
            self.policy.reset()
            try:
                cur = self.policy.s2_step(rgb, depth, pose, instruction, self.camera_intrinsic, False)
            except Exception:
                self.policy.reset()
                self.s2_output.output_pixel = None
                self.s2_output.output_action = [0]
                self.s2_output.output_latent = None
                return
        o = self.s2_output
        o.output_pixel, o.output_action, o.output_latent = cur.output_pixel, cur.output_action, cur.output_latent
        o.idx, o.rgb_memory, o.depth_memory = idx, rgb, depth

    def should_infer_s2(self, mode):                                             # L210-241
        if self.episode_step == 0:
            return True
        if mode == "sync":
            return self.s2_output.output_action is None
        if mode == "partial_async":
            if self.dual_forward_step >= self.sys2_max_forward_step:
                return True
            o = self.s2_output
            return o.output_action is None and o.output_pixel is None and o.output_latent is None
        raise ValueError("Invalid mode: {}".format(mode))

    def step(self, obs):                                                         # L243-407
        mode = self.mode
        obs = obs[0]
        rgb, depth, instruction = obs["rgb"], obs["depth"], obs["instruction"]
        pose = np.eye(4, dtype=np.int64)
        if self.should_infer_s2(mode) or self.look_down:
            self.dual_forward_step = 0
            self._run_s2(rgb, depth, pose, instruction, self.look_down, self.episode_step)
        else:
            self.policy.step_no_infer(rgb, depth, pose)
        o = self.s2_output
        output = {}
        if o.output_action is not None:
            output["action"] = [o.output_action[0]]
            o.output_action = o.output_action[1:]
            if o.output_action == []:
                o.output_action = None
            if output["action"][0] == 5:
                self.look_down = True
                o.output_action = o.output_pixel = o.output_latent = None
                output["action"] = [-1]
                self.sys1_infer_times = 0
            else:
                self.look_down = False
                if self.sys1_infer_times > 0:
                    self.dual_forward_step += 1
        else:
            self.look_down = False
            assert o.output_latent is not None, "S2 output should be either action or latent, but got neither!"
            if mode != "sync":
                def prep_rgb(x):
                    return np.array(Image.fromarray(x).resize((224, 224))) / 255.0

                def prep_depth(x):
                    d = np.array(Image.fromarray(x[:, :, 0]).resize((224, 224))) * 10.0
                    d[d > self.sys1_depth_threshold] = self.sys1_depth_threshold
                    return d
                rgbs = torch.stack([torch.from_numpy(prep_rgb(o.rgb_memory)), torch.from_numpy(prep_rgb(rgb))]).unsqueeze(0)
                depths = torch.stack([torch.from_numpy(prep_depth(o.depth_memory)),
                                      torch.from_numpy(prep_depth(depth))]).unsqueeze(0).unsqueeze(-1)
                s1 = self.policy.s1_step_latent(rgbs, depths, o.output_latent)
            else:
                s1 = self.policy.s1_step_latent(rgb, depth * 10000.0, o.output_latent)
            output["action"] = [-1] if s1.idx == [] else [s1.idx[0]]
            o.output_action = s1.idx[1:] if len(s1.idx) > 1 else None
            if o.output_action == []:
                o.output_action = None
            o.output_pixel = None
            if mode == "sync":
                o.output_latent = None
            else:
                if len(s1.idx) < self.sys1_forward_step:
                    if len(s1.idx) + self.dual_forward_step < self.sys2_max_forward_step:
                        self.dual_forward_step = self.sys2_max_forward_step - len(s1.idx)
                self.sys1_infer_times += 1
                self.dual_forward_step += 1
        self.episode_step += 1
        return [{"action": output["action"], "ideal_flag": True}]
