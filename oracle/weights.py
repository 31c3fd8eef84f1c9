"""Deterministic synthetic weights for the NavDP System-1 head -- TEST INFRASTRUCTURE.

No checkpoint is available offline, so every parity test runs on seeded random weights.  They are
generated with numpy's PCG64 (bit-reproducible across machines for a fixed numpy version) from the tensor manifest in
oracle/navdp_manifest.json, which was dumped from the reference class NavDP_Policy_DPT_CriticSum_DAT(memory_size=2,
navdp_version=0.1).state_dict() (navdp.py L16-114) by oracle/gen_golden.py.

Scales keep activations O(1) through the 12-block ViTs and the 16-layer decoder; the tensors the reference
zero-initialises (former_query, former_pe, cond_pos_embed, out_pos_embed: navdp_backbone.py L140, L147; navdp.py
L69-70) get non-zero values so their code paths are exercised.
"""
import json
import os
import zlib

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))


def manifest():
    with open(os.path.join(_HERE, "navdp_manifest.json")) as fh:
        return json.load(fh)


def _std(name, shape):
    last = name.split(".")[-1]
    if "norm" in name and last == "weight" and len(shape) == 1:
        return ("affine", 0.1)
    if name.endswith("layernorm.weight"):
        return ("affine", 0.1)
    if last == "gamma":
        return ("affine", 0.1)
    if last in ("bias", "in_proj_bias"):
        return ("normal", 0.02)
    if name.endswith("pos_embed") or "cls_token" in name or "mask_token" in name:
        return ("normal", 0.02 if "rgbd_encoder" in name else 0.2)
    if "former_query" in name or "former_pe" in name or "target_embedding" in name or "position_embedding" in name:
        return ("normal", 0.2)
    if name.endswith("positional_encoding.pe"):
        return ("keep", 0.0)
    if len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        return ("normal", 1.0 / np.sqrt(fan_in))
    return ("normal", 0.02)


def make_state_dict(seed=0, dtype=torch.float32):
    """name -> tensor for every entry of the manifest (reference key names, fp32)."""
    out = {}
    for name, (shape, _dt) in manifest().items():
        kind, s = _std(name, shape)
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        if kind == "keep":
            # PositionalEncoding.pe buffer (navdp_backbone.py L27-34); never read by forward, regenerated for completeness
            max_len, dim = shape
            pe = np.zeros((max_len, dim), dtype=np.float32)
            pos = np.arange(max_len, dtype=np.float32)[:, None]
            div = np.exp(np.arange(0, dim, 2, dtype=np.float32) * (-np.log(10000.0) / dim))
            pe[:, 0::2] = np.sin(pos * div)
            pe[:, 1::2] = np.cos(pos * div)
            arr = pe
        else:
            arr = rng.standard_normal(shape, dtype=np.float32) * np.float32(s)
            if kind == "affine":
                arr = arr + np.float32(1.0)
        out[name] = torch.from_numpy(np.ascontiguousarray(arr)).to(dtype)
    return out


def make_inputs(seed, B, T=32, Ns=32, K=20, frames=2, vlm_dim=3584, n_query=4):
    """Seeded synthetic inputs of the shapes in SURVEY.md §8d."""
    rng = np.random.Generator(np.random.PCG64([seed, 12345]))
    f32 = np.float32
    return {
        "latents": torch.from_numpy(rng.standard_normal((B, n_query, vlm_dim), dtype=f32)),
        "rgb": torch.from_numpy(rng.random((B, frames, 224, 224, 3), dtype=f32)),
        "depth": torch.from_numpy(rng.random((B, frames, 224, 224, 1), dtype=f32) * f32(5.0)),
        "x_init": torch.from_numpy(rng.standard_normal((B * Ns, T, 3), dtype=f32)),
        "step_noise": torch.from_numpy(rng.standard_normal((max(K - 1, 0), B * Ns, T, 3), dtype=f32)),
        "goal": torch.from_numpy(rng.standard_normal((B, 1, 384), dtype=f32)),
        "rgbd": torch.from_numpy(rng.standard_normal((B, 16 * frames, 384), dtype=f32)),
    }
