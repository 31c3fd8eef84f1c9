"""Reading an InternVLA-N1 checkpoint directory the way `InternVLAN1ForCausalLM.from_pretrained(model_path, ...)` is used
by the reference (internvla_n1_policy.py L33-38, habitat_vln_evaluator.py L117-122): `config.json` + sharded
`*.safetensors` (or `pytorch_model*.bin`).  Host-only; the tensors go to the library through load_state_dict.

config.json is accepted in both layouts seen in the wild: the flat transformers-4.51 one the reference was released with
(`hidden_size`, ..., `rope_scaling.mrope_section`, `vision_config`) and the nested 5.x one (`text_config`).
"""
import glob
import json
import os

import torch

from .qwen import QWEN25VL_7B


def cfg_from_hf(conf):
    """HF Qwen2.5-VL / InternVLA-N1 config dict -> the n1b200 dimension dict (keys of qwen.QWEN25VL_7B)."""
    t = conf.get("text_config") or conf
    v = conf.get("vision_config") or {}
    rope = t.get("rope_scaling") or t.get("rope_parameters") or conf.get("rope_scaling") or {}
    heads = int(t["num_attention_heads"])
    cfg = dict(QWEN25VL_7B)
    cfg.update(
        layers=int(t["num_hidden_layers"]), hidden=int(t["hidden_size"]), heads=heads,
        kv_heads=int(t.get("num_key_value_heads", heads)), head_dim=int(t.get("head_dim") or t["hidden_size"] // heads),
        inter=int(t["intermediate_size"]), vocab=int(t["vocab_size"]), rms_eps=float(t.get("rms_norm_eps", 1e-6)),
        rope_theta=float(rope.get("rope_theta") or t.get("rope_theta", 1000000.0)),
        mrope=[int(x) for x in rope.get("mrope_section", cfg["mrope"])])
    if v:
        cfg.update(
            v_depth=int(v.get("depth", cfg["v_depth"])), v_hidden=int(v.get("hidden_size", cfg["v_hidden"])),
            v_heads=int(v.get("num_heads", cfg["v_heads"])), v_inter=int(v.get("intermediate_size", cfg["v_inter"])),
            v_patch=int(v.get("patch_size", cfg["v_patch"])), v_tpatch=int(v.get("temporal_patch_size", cfg["v_tpatch"])),
            v_merge=int(v.get("spatial_merge_size", cfg["v_merge"])), v_window=int(v.get("window_size", cfg["v_window"])),
            v_out=int(v.get("out_hidden_size", cfg["hidden"])),
            fullatt=[int(x) for x in v.get("fullatt_block_indexes", cfg["fullatt"])])
    cfg["n_query"] = int(conf.get("n_query", cfg["n_query"]))
    return cfg


def read_checkpoint(path):
    """-> (n1b200 cfg dict, HF config dict, state_dict of CPU tensors)."""
    with open(os.path.join(path, "config.json")) as fh:
        conf = json.load(fh)
    sd = {}
    index = os.path.join(path, "model.safetensors.index.json")
    shards = sorted(glob.glob(os.path.join(path, "*.safetensors")))
    if os.path.exists(index):
        with open(index) as fh:
            shards = sorted({os.path.join(path, f) for f in json.load(fh)["weight_map"].values()})
    if shards:
        from safetensors.torch import load_file
        for f in shards:
            sd.update(load_file(f, device="cpu"))
    else:
        bins = sorted(glob.glob(os.path.join(path, "pytorch_model*.bin")))
        if not bins:
            raise FileNotFoundError("no *.safetensors / pytorch_model*.bin under %s" % path)
        for f in bins:
            sd.update(torch.load(f, map_location="cpu", weights_only=True))
    return cfg_from_hf(conf), conf, sd
