"""Batched dual-system policy: the host side of System 2 (prompt building, image history, answer parsing) and of
System 1 (trajectory -> action ids) for B environments, on top of the n1b200 model mirror.

Mirrors `InternVLAN1Net` (internnav/model/basemodel/internvla_n1/internvla_n1_policy.py): `init_prompts` L60-86,
`reset` L88-94, `parse_actions` L96-102, `step_no_infer` L104-108, `s2_step` L110-198, `s1_step_latent` L200-215 --
per-environment state kept in a list, and the three model calls of one `s2_step` (vision tower + `generate`, then vision
tower + prefill again inside `generate_latents`) replaced by ONE `generate_with_latents` call for the whole batch.

The HF processor (chat template + tokenizer + Qwen2-VL image processor) is injected: it is the reference's own
collaborator (`AutoProcessor.from_pretrained(model_path)`, L44-47) and needs the checkpoint directory, which is not
available offline.  Anything with `apply_chat_template`, `__call__(text=[...], images=[...], return_tensors="pt")` and
`.tokenizer.decode` works (tests use oracle/policy_script.FakeProcessor).
"""
import copy
import itertools
import re
from collections import OrderedDict
from types import SimpleNamespace

import numpy as np
import torch
from PIL import Image

from .postprocess import batched_traj_to_actions, chunk_token, s1_action_list

DEFAULT_IMAGE_TOKEN = "<image>"
PROMPT = ("You are an autonomous navigation assistant. Your task is to <instruction>. Where should you go next to stay "
          "on track? Please output the next waypoint\'s coordinates in the image. Please output STOP when you have "
          "successfully completed the task.")
CONJUNCTION = "you can see "
ACTIONS2IDX = OrderedDict({"STOP": [0], "↑": [1], "←": [2], "→": [3], "↓": [5]})
_ACTION_RE = re.compile("|".join(re.escape(a) for a in ACTIONS2IDX))


def parse_actions(output):
    """L96-102: every STOP / arrow occurrence, in order, as action ids."""
    return list(itertools.chain.from_iterable(ACTIONS2IDX[m] for m in _ACTION_RE.findall(output)))


def split_and_clean(text):
    """vln_utils.py L19-33: split on <image>, drop newlines and surrounding blanks, skip empty pieces."""
    out = []
    for part in re.split(r"(<image>)", text):
        if part == DEFAULT_IMAGE_TOKEN:
            out.append(part)
        else:
            part = part.replace("\n", "").strip()
            if part:
                out.append(part)
    return out


class S2Output(SimpleNamespace):
    def __init__(self):
        super().__init__(output_action=None, output_pixel=None, output_latent=None)


class _Episode:
    """Per-environment conversation state (the instance attributes of the reference class, L52-57 / L88-94)."""

    def __init__(self):
        self.rgb_list = []
        self.episode_idx = 0
        self.conversation_history = []
        self.llm_output = ""
        self.input_images = []


class InternVLAN1Policy:
    def __init__(self, model, processor, num_envs=1, num_history=8, resize_w=384, resize_h=384, continuous_traj=True,
                 max_new_tokens=128, device=None):
        self.model, self.processor = model, processor
        self.num_history, self.resize_w, self.resize_h = num_history, resize_w, resize_h
        self.continuous_traj = continuous_traj
        self.max_new_tokens = max_new_tokens
        self.device = device if device is not None else getattr(model, "device", "cpu")
        self.episodes = [_Episode() for _ in range(num_envs)]

    def eval(self):
        return self

    def reset(self, env_ids=None):
        for e in (range(len(self.episodes)) if env_ids is None else env_ids):
            self.episodes[e] = _Episode()

    def step_no_infer(self, env_ids, rgbs, depths=None, poses=None):
        for e, rgb in zip(env_ids, rgbs):
            ep = self.episodes[e]
            ep.rgb_list.append(Image.fromarray(rgb).convert("RGB").resize((self.resize_w, self.resize_h)))
            ep.episode_idx += 1

    # ------------------------------------------------------------------ System 2
    def _build_inputs(self, ep, rgb, instruction, look_down):
        """L113-164 for one environment -> processor output (input_ids [1, S], pixel_values, image_grid_thw)."""
        image = Image.fromarray(rgb).convert("RGB")
        if not look_down:
            image = image.resize((self.resize_w, self.resize_h))
            ep.rgb_list.append(image)
            ep.conversation_history = []
            text = PROMPT.replace("<instruction>.", instruction)
            if ep.episode_idx == 0:
                history_id = []
            else:
                history_id = np.unique(np.linspace(0, ep.episode_idx - 1, self.num_history, dtype=np.int32)).tolist()
                text += " These are your historical observations: %s." % ((DEFAULT_IMAGE_TOKEN + "\n") * len(history_id))
            ep.input_images = [ep.rgb_list[i] for i in sorted(history_id)] + ep.rgb_list[-1:]
            img_id = 0
            ep.episode_idx += 1
        else:  # the look-down frame continues the conversation and never enters the history
            ep.input_images.append(image)
            img_id = -1
            assert ep.llm_output != "", "Last llm_output should not be empty when look down"
            text = ""
            ep.conversation_history.append({"role": "assistant", "content": [{"type": "text", "text": ep.llm_output}]})
        text += " %s." % (CONJUNCTION + DEFAULT_IMAGE_TOKEN)
        content = []
        for part in split_and_clean(copy.deepcopy(text)):
            if part == DEFAULT_IMAGE_TOKEN:
                content.append({"type": "image", "image": ep.input_images[img_id]})
                img_id += 1
            else:
                content.append({"type": "text", "text": part})
        ep.conversation_history.append({"role": "user", "content": content})
        chat = self.processor.apply_chat_template(ep.conversation_history, tokenize=False, add_generation_prompt=True)
        return self.processor(text=[chat], images=ep.input_images, return_tensors="pt")

    def s2_step(self, env_ids, rgbs, depths, poses, instructions, intrinsic, look_downs):
        """One System-2 consultation for the listed environments (one model call).  Returns a list with, per
        environment, an S2Output (discrete `output_action` list, or `output_pixel` + `output_latent` [1, n_query, H]) or
        the Exception that environment's host-side preparation raised."""
        results = [None] * len(env_ids)
        prepared = []
        for j, (e, rgb, ins, ld) in enumerate(zip(env_ids, rgbs, instructions, look_downs)):
            try:
                prepared.append((j, self._build_inputs(self.episodes[e], rgb, ins, ld)))
            except Exception as exc:  # noqa: BLE001 -- reported per environment; the agent applies the retry rule
                results[j] = exc
        if not prepared:
            return results
        prompts = [inp["input_ids"][0].tolist() for _, inp in prepared]
        pixels = torch.cat([inp["pixel_values"] for _, inp in prepared], dim=0)
        grids = torch.cat([torch.stack(list(inp["image_grid_thw"])).reshape(-1, 3) for _, inp in prepared], dim=0)
        with torch.no_grad():
            out = self.model.generate_with_latents(prompts, pixels, grids, max_new_tokens=self.max_new_tokens)
        for n, (j, _) in enumerate(prepared):
            ep = self.episodes[env_ids[j]]
            ep.llm_output = self.processor.tokenizer.decode(out.generated[n], skip_special_tokens=True)
            res = S2Output()
            try:
                if re.search(r"\d", ep.llm_output):   # pixel goal "y x" -> [x, y] plus the latent plan (L179-190)
                    coord = [int(c) for c in re.findall(r"\d+", ep.llm_output)]
                    res.output_pixel = np.array([int(coord[1]), int(coord[0])])
                    res.output_latent = out.latents[n:n + 1]
                else:
                    res.output_action = parse_actions(ep.llm_output)
                results[j] = res
            except Exception as exc:  # noqa: BLE001 -- e.g. an answer with a single number (IndexError at L181 as well)
                results[j] = exc
        return results

    # ------------------------------------------------------------------ System 1
    def s1_step_latent(self, env_ids, rgbs, depths, latents):
        """L200-215 for the listed environments in one generate_traj call: rgbs / depths are the per-environment
        [1, 2, 224, 224, 3] / [1, 2, 224, 224, 1] stacks of the agent, latents the [1, n_query, H] plans."""
        lat = torch.cat([l.reshape(1, *l.shape[-2:]) for l in latents], dim=0)
        rgb = torch.cat([torch.as_tensor(r).reshape(1, *torch.as_tensor(r).shape[-4:]) for r in rgbs], dim=0)
        dep = torch.cat([torch.as_tensor(d).reshape(1, *torch.as_tensor(d).shape[-4:]) for d in depths], dim=0)
        with torch.no_grad():
            traj = self.model.generate_traj(traj_latents=lat, images_dp=rgb.float(), depths_dp=dep.float())
        B = len(env_ids)
        if self.continuous_traj:
            lists = batched_traj_to_actions(traj, B, max_actions=4)
        else:  # one sampled trajectory per environment, tokens -> ids (vln_utils.py L36-60)
            per = traj.shape[0] // B
            lists = [chunk_token(traj.view(B, per, *traj.shape[1:])[b, np.random.choice(per)]) for b in range(B)]
        return [SimpleNamespace(idx=s1_action_list(a)) for a in lists]
