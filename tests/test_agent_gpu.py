"""Closed loop on the GPU: batched agent -> batched policy -> n1b200 model (tiny Qwen config + full-size NavDP head).
Checks that the pieces compose (ragged prompts with growing image history, greedy decode + KV-reuse latents, goal-frame
memory into System 1, action queues) and that two runs give identical actions.  Parity of each piece is covered by
test_s1_gpu / test_s2_gpu / test_agent / test_policy."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ARROWS = "↑←→"


class SyntheticProcessor:
    """Deterministic stand-in for the HF processor with REAL shapes: characters -> text token ids, every image -> a
    [1, 8, 8] patch grid (16 image tokens) whose pixel rows are seeded by the frame's content."""

    class _Tok:
        padding_side = "left"

        def __init__(self, kind):
            self.kind, self.n = kind, 0

        def decode(self, ids, skip_special_tokens=True):
            """Random-weight token ids -> a well-formed answer: a pixel goal "y, x" or a run of arrows.  A random
            tiny model tends to repeat itself, so the KIND of answer is scripted: "pixel", "arrows", or alternating."""
            ids = [int(i) for i in ids if int(i) < 151643] or [0, 0]
            self.n += 1
            kind = self.kind if self.kind != "mixed" else ("pixel" if self.n % 3 else "arrows")
            if kind == "arrows":
                return "".join(ARROWS[i % 3] for i in ids[:3])
            return "%d, %d" % (ids[0] % 480, ids[-1] % 640)

    def __init__(self, kind="mixed"):
        self.tokenizer = self._Tok(kind)

    def apply_chat_template(self, conversation, tokenize=False, add_generation_prompt=True):
        parts = []
        for turn in conversation:
            parts.append(turn["role"] + ":" + "".join("<image>" if c["type"] == "image" else c["text"]
                                                      for c in turn["content"]))
        return "\n".join(parts)

    def __call__(self, text, images, return_tensors="pt"):
        ids, rows, grids = [], [], []
        pieces = text[0].split("<image>")
        for i, piece in enumerate(pieces):
            ids += [1000 + (ord(c) % 5000) for c in piece[-48:]]          # keep prompts short: last 48 chars per piece
            if i < len(pieces) - 1:
                tag = int(np.asarray(images[i]).reshape(-1)[0])
                g = torch.Generator().manual_seed(tag)
                rows.append(torch.randn(64, 1176, generator=g))
                grids.append(torch.tensor([1, 8, 8]))
                ids += [151652] + [151655] * 16 + [151653]
        return {"input_ids": torch.tensor([ids]), "pixel_values": torch.cat(rows), "image_grid_thw": torch.stack(grids)}


def _obs(k, e):
    rgb = np.full((48, 64, 3), (17 * e + k) % 256, dtype=np.uint8)
    depth = np.full((48, 64, 1), 0.05 + 0.01 * ((k + e) % 30), dtype=np.float32)
    return {"rgb": rgb, "depth": depth, "instruction": "walk to the door %d" % e}


def _rollout(model, B, frames, mode, kind="mixed", gpu_frames=False):
    from internnav_b200.agent import InternVLAN1Agent
    from internnav_b200.policy import InternVLAN1Policy
    from internnav_b200.preprocess import FramePreprocessor
    torch.manual_seed(0)
    pol = InternVLAN1Policy(model, SyntheticProcessor(kind), num_envs=B, num_history=4, resize_w=56, resize_h=56,
                            max_new_tokens=6)
    ag = InternVLAN1Agent(pol, num_envs=B, infer_mode=mode, sys2_max_forward_step=4,
                          preprocessor=FramePreprocessor("cuda:0") if gpu_frames else None)
    ag.reset()
    acts = []
    for k in range(frames):
        if k == 5:
            ag.reset([1])
        out = ag.step([_obs(k, e) for e in range(B)])
        acts.append([o["action"][0] for o in out])
    return acts, ag.calls


def test_closed_loop_batched_agent():
    from internnav_b200.internvla_n1 import InternVLAN1ForCausalLM
    from internnav_b200.manifest import random_navdp_state_dict
    from oracle import qwen_oracle as Q
    cfg = Q.tiny_cfg()
    s2_sd = Q.make_s2_state_dict(cfg, seed=3, lm_head=True)
    s1_sd = random_navdp_state_dict(seed=4, vlm_token_dim=cfg["hidden"])
    model = InternVLAN1ForCausalLM(cfg, device="cuda:0")
    model.load_parts(s2_sd, s1_sd)
    B, frames = 3, 9
    # System 1 draws its noise from the torch generators, which _rollout seeds: the runs are comparable
    a1, calls = _rollout(model, B, frames, "partial_async", "pixel")
    a2, calls2 = _rollout(model, B, frames, "partial_async", "pixel", gpu_frames=True)
    # (sync mode hands RAW frames to System 1 -- internvla_n1_agent.py L335 -- which the NavDP head does not take, in
    #  the reference either; it is exercised with discrete answers only)
    a3, calls3 = _rollout(model, B, frames, "sync", "arrows")
    a4, calls4 = _rollout(model, B, frames, "partial_async", "mixed")
    print("actions", a1, calls, "| sync/arrows", a3, calls3, "| mixed", a4, calls4)
    # Pillow on the host and the resize kernels give bit-identical System-1 inputs, hence identical rollouts
    assert a1 == a2 and calls == calls2, "GPU frame preprocessing changed the rollout"
    for acts in (a1, a3, a4):
        assert len(acts) == frames and all(len(r) == B and all(a in (-1, 0, 1, 2, 3) for a in r) for r in acts)
    assert calls["s1"] >= 2 and calls["s1_envs"] > calls["s1"], "System-1 calls were not exercised / batched"
    assert calls["s2_envs"] > calls["s2"], "System-2 calls were not batched across environments"
    assert calls3["s2"] >= 2 and calls3["s1"] == 0 and calls4["s1"] >= 1
    with pytest.raises(ValueError):      # raw frames into the NavDP head are refused, not read out of bounds
        model.model.navdp.rgbd_encoder(torch.zeros(1, 48, 64, 3), torch.zeros(1, 48, 64, 1))
