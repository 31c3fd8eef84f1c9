"""Scripted processor / language model for driving the System-2 host logic of the policy -- TEST INFRASTRUCTURE.

The policy's `s2_step` (internvla_n1_policy.py L113-195) is string and list work around three opaque collaborators: the
HF processor (chat template, tokeniser, image patches), `model.generate` and `model.generate_latents`.  The stand-ins
here make that work observable and reproducible without a checkpoint:

  FakeProcessor.apply_chat_template  renders the conversation as plain text with <image> placeholders kept;
  FakeProcessor.__call__             "tokenises": every character -> 1000 + ord, every <image> -> <vision_start>,
                                     n x <image_pad>, <vision_end> with n from the image size; pixel_values rows carry
                                     the frame tag of the image they came from (agent_script.make_obs frames);
  FakeTokenizer.decode               inverts the character code (special ids dropped);
  ScriptedLLM                        answers with the next scripted string.
"""
import numpy as np
import torch

IMAGE_PAD, VISION_START, VISION_END = 151655, 151652, 151653
CHAR0 = 1000


def encode(text):
    return [CHAR0 + ord(c) for c in text]


class _Batch(dict):
    """BatchFeature stand-in: mapping (for **inputs) with attribute access and .to()."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        return self


class FakeTokenizer:
    padding_side = "left"

    def decode(self, ids, skip_special_tokens=True):
        return "".join(chr(int(i) - CHAR0) for i in ids if CHAR0 <= int(i) < 151643)  # special ids dropped


class FakeProcessor:
    def __init__(self):
        self.tokenizer = FakeTokenizer()
        self.log = []

    def apply_chat_template(self, conversation, tokenize=False, add_generation_prompt=True):
        out = []
        for turn in conversation:
            parts = []
            for c in turn["content"]:
                parts.append("<image>" if c["type"] == "image" else c["text"])
            out.append("<|%s|>%s" % (turn["role"], "".join(parts)))
        return "\n".join(out) + ("\n<|assistant|>" if add_generation_prompt else "")

    @staticmethod
    def image_tag(img):
        return int(np.asarray(img).reshape(-1)[0])

    def __call__(self, text, images, return_tensors="pt"):
        assert len(text) == 1
        ids, k = [], 0
        grids, rows = [], []
        pieces = text[0].split("<image>")
        for i, piece in enumerate(pieces):
            ids += encode(piece)
            if i < len(pieces) - 1:
                w, h = images[k].size
                gh, gw = max(2, (h // 28) * 2), max(2, (w // 28) * 2)
                grids.append(torch.tensor([1, gh, gw]))
                rows.append(torch.full((gh * gw, 4), float(self.image_tag(images[k]))))
                ids += [VISION_START] + [IMAGE_PAD] * (gh * gw // 4) + [VISION_END]
                k += 1
        assert k == len(images), "placeholders and images do not match"
        self.log.append({"text": text[0], "images": [self.image_tag(im) for im in images],
                         "sizes": [list(im.size) for im in images]})
        return _Batch(input_ids=torch.tensor([ids]), attention_mask=torch.ones(1, len(ids), dtype=torch.int64),
                      pixel_values=torch.cat(rows) if rows else torch.zeros(0, 4), image_grid_thw=torch.stack(grids))


class ScriptedLLM:
    """`model.generate(...).sequences`, `model.generate_latents(...)`, `model.generate_traj(...)` of the reference."""

    def __init__(self, answers, trajs=None):
        self.answers, self.n = answers, 0
        self.trajs, self.n_traj = trajs or [], 0
        self.log = []

    def generate(self, input_ids=None, max_new_tokens=128, do_sample=False, use_cache=True, past_key_values=None,
                 return_dict_in_generate=True, **inputs):
        ans = self.answers[self.n % len(self.answers)]
        self.n += 1
        seq = torch.cat([input_ids, torch.tensor([encode(ans) + [151645]])], dim=1)
        self.log.append(["generate", int(input_ids.shape[1]), max_new_tokens, bool(do_sample)])
        return type("Out", (), {"sequences": seq})()

    def generate_latents(self, output_ids, pixel_values, image_grid_thw):
        self.log.append(["generate_latents", int(output_ids.shape[1]), int(pixel_values.shape[0]),
                         [int(v) for v in image_grid_thw.reshape(-1)]])
        return torch.tensor([float(self.n)])

    def generate_traj(self, traj_latents=None, images_dp=None, depths_dp=None, **kw):
        t = self.trajs[self.n_traj % len(self.trajs)]
        self.n_traj += 1
        self.log.append(["generate_traj", float(traj_latents.reshape(-1)[0])])
        return torch.tensor(t, dtype=torch.float32)


def random_answers(rng, n=30):
    """Pixel goals ("123, 456" style and wordier), arrows, STOP, look-down, and mixtures the regexes must sort out."""
    out = []
    for _ in range(n):
        u = rng.random()
        if u < 0.45:
            y, x = int(rng.integers(0, 480)), int(rng.integers(0, 640))
            out.append(rng.choice(["%d, %d", "(%d, %d)", "The next waypoint is at %d %d."]) % (y, x))
        elif u < 0.6:
            out.append("↓")
        elif u < 0.7:
            out.append("STOP")
        else:
            k = int(rng.integers(1, 5))
            out.append("".join(rng.choice(["↑", "←", "→"], size=k)) + ("STOP" if rng.random() < 0.1 else ""))
    return out


def random_trajs(rng, n=12, ns=4, T=32):
    """Small trajectory batches [ns, T, 3] of per-step deltas (the layout generate_traj returns)."""
    out = []
    for _ in range(n):
        fwd = rng.uniform(0.0, 0.35, size=(ns, T, 1))
        lat = rng.normal(0.0, 0.12, size=(ns, T, 1)) + rng.choice([-0.15, 0.0, 0.15])
        yaw = rng.normal(0.0, 0.05, size=(ns, T, 1))
        out.append(np.concatenate([fwd, lat, yaw], axis=-1).round(4).tolist())
    return out
