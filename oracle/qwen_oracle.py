"""CPU restatement of the System-2 forward used by InternVLA-N1 -- TEST INFRASTRUCTURE, not product code.

The arithmetic of Qwen2.5-VL (vision tower + decoder) is NOT under /root/reference: the reference subclasses
`transformers==4.51.0` (requirements/internvla_n1.txt L7; call sites internvla_n1.py L9-14, L39-48, L132, L185, L206,
L330-344).  This file restates the published algorithm of modeling_qwen2_5_vl.py in plain fp32 PyTorch functions over a
state_dict with the 4.51 checkpoint key layout (`visual.*`, `model.layers.*`, `model.embed_tokens.weight`,
`model.norm.weight`, plus InternVLA-N1's `model.latent_queries`, internvla_n1_arch.py L123), and restates the N1 glue
`generate_latents` (internvla_n1.py L320-347) and `get_rope_index_25` (internnav/dataset/rope2d.py L6-181).

Pinning: tests/test_oracle_s2.py checks (a) the blocks against the container's transformers 5.5 implementation of the
same modules (same math, newer packaging) on seeded tiny configs, (b) `rope_index` bit-exactly against the reference's
own rope2d.get_rope_index_25 where /root/reference exists, and against tests/golden/rope_index.json everywhere.
The reference itself holds no golden vector for this path: beyond those two anchors parity is unpinned.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

IMAGE_TOKEN_INDEX = 151655   # internvla_n1.py L19
TRAJ_TOKEN_INDEX = 151667    # internvla_n1.py L18
VISION_START = 151652

QWEN25VL_7B = dict(
    v_depth=32, v_hidden=1280, v_heads=16, v_inter=3420, v_patch=14, v_tpatch=2, v_merge=2, v_window=112, v_out=3584,
    fullatt=[7, 15, 23, 31],
    layers=28, hidden=3584, heads=28, kv_heads=4, head_dim=128, inter=18944, vocab=152064,
    rms_eps=1e-6, rope_theta=1000000.0, mrope=[16, 24, 24], n_query=4)


def tiny_cfg(**over):
    c = dict(v_depth=3, v_hidden=160, v_heads=2, v_inter=216, v_patch=14, v_tpatch=2, v_merge=2, v_window=112,
             v_out=256, fullatt=[1], layers=2, hidden=256, heads=2, kv_heads=1, head_dim=128, inter=512, vocab=152064,
             rms_eps=1e-6, rope_theta=1000000.0, mrope=[16, 24, 24], n_query=4)
    c.update(over)
    return c


def _rms(x, w, eps):
    """Qwen2_5_VLRMSNorm.forward (modeling_qwen2_5_vl.py: fp32 variance, weight * normalised)."""
    xf = x.float()
    xf = xf * torch.rsqrt(xf.pow(2).mean(-1, keepdim=True) + eps)
    return w.to(x.dtype) * xf.to(x.dtype)


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


# ------------------------------------------------------------------------------------------------ integer planning
def vit_indices(grid_thw, merge=2, patch=14, window=112):
    """rot_pos_emb ids + get_window_index + cu_seqlens of Qwen2_5_VisionTransformerPretrainedModel (restated).
    Returns pos_ids [N,2] (processor patch order), window_index [N/4], cu_window (unique_consecutive), cu_full."""
    pos_ids, window_index, cu_window, cu_full = [], [], [0], [0]
    win = window // merge // patch
    wid = 0
    for t, h, w in [tuple(int(v) for v in g) for g in grid_thw]:
        hp = torch.arange(h).unsqueeze(1).expand(-1, w).reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        wp = torch.arange(w).unsqueeze(0).expand(h, -1).reshape(h // merge, merge, w // merge, merge).permute(0, 2, 1, 3).flatten()
        pos_ids.append(torch.stack([hp, wp], dim=-1).repeat(t, 1))
        gh, gw = h // merge, w // merge
        index = torch.arange(t * gh * gw).reshape(t, gh, gw)
        pad_h, pad_w = win - gh % win, win - gw % win
        nh, nw = (gh + pad_h) // win, (gw + pad_w) // win
        ip = F.pad(index, (0, pad_w, 0, pad_h), "constant", -100).reshape(t, nh, win, nw, win)
        ip = ip.permute(0, 1, 3, 2, 4).reshape(t, nh * nw, win, win)
        seqlens = (ip != -100).sum([2, 3]).reshape(-1)
        ip = ip.reshape(-1)
        window_index.append(ip[ip != -100] + wid)
        cu_window.extend((seqlens.cumsum(0) * merge * merge + cu_window[-1]).tolist())
        wid += t * gh * gw
        for _ in range(t):
            cu_full.append(cu_full[-1] + h * w)
    cu_window = torch.unique_consecutive(torch.tensor(cu_window, dtype=torch.int32))
    return torch.cat(pos_ids), torch.cat(window_index), cu_window, torch.tensor(cu_full, dtype=torch.int32)


def rope_index(input_ids, image_grid_thw, merge=2):
    """get_rope_index_25 restricted to images and no padding mask (rope2d.py L67-157): input_ids [B,S] int64,
    image_grid_thw [n,3] -> position_ids [3,B,S] int64, deltas [B,1]."""
    B, S = input_ids.shape
    pos = torch.ones(3, B, S, dtype=torch.long)
    deltas = []
    img = 0
    for i in range(B):
        ids = input_ids[i].tolist()
        starts = [j for j in range(S - 1) if ids[j] == VISION_START]
        n_img = sum(1 for j in starts if ids[j + 1] == IMAGE_TOKEN_INDEX)
        chunks, st = [], 0
        for _ in range(n_img):
            ed = ids.index(IMAGE_TOKEN_INDEX, st)
            t, h, w = (int(v) for v in image_grid_thw[img])
            img += 1
            gt, gh, gw = t, h // merge, w // merge
            text_len = ed - st
            st_idx = int(chunks[-1].max()) + 1 if chunks else 0
            chunks.append(torch.arange(text_len).view(1, -1).expand(3, -1) + st_idx)
            t_index = (torch.arange(gt).view(-1, 1).expand(-1, gh * gw) * 0 * 2).long().flatten()  # second_per_grid_t = 0
            h_index = torch.arange(gh).view(1, -1, 1).expand(gt, -1, gw).flatten()
            w_index = torch.arange(gw).view(1, 1, -1).expand(gt, gh, -1).flatten()
            chunks.append(torch.stack([t_index, h_index, w_index]) + text_len + st_idx)
            st = ed + gt * gh * gw
        if st < S:
            st_idx = int(chunks[-1].max()) + 1 if chunks else 0
            chunks.append(torch.arange(S - st).view(1, -1).expand(3, -1) + st_idx)
        p = torch.cat(chunks, dim=1).reshape(3, -1)
        pos[:, i] = p
        deltas.append(int(p.max()) + 1 - S)
    return pos, torch.tensor(deltas).unsqueeze(1)


# ------------------------------------------------------------------------------------------------ vision tower
def vit_forward(sd, cfg, pixel_values, grid_thw, p="visual."):
    """Qwen2_5_VisionTransformerPretrainedModel.forward (4.51: returns the merged tokens in original order)."""
    Hv, heads, unit = cfg["v_hidden"], cfg["v_heads"], cfg["v_merge"] ** 2
    hd = Hv // heads
    dt = pixel_values.dtype
    x = F.linear(pixel_values, sd[p + "patch_embed.proj.weight"].reshape(Hv, -1).to(dt))  # Conv3d, stride = kernel
    pos_ids, window_index, cu_window, cu_full = vit_indices(grid_thw, cfg["v_merge"], cfg["v_patch"], cfg["v_window"])
    pos_ids, window_index = pos_ids.to(x.device), window_index.to(x.device)
    N = x.shape[0]
    dim = hd // 2
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, dim, 2, dtype=torch.float, device=x.device) / dim))
    freqs = torch.outer(torch.arange(int(max(max(g[1], g[2]) for g in grid_thw)), device=x.device, dtype=torch.float), inv_freq)
    rot = freqs[pos_ids].flatten(1)  # [N, dim]
    x = x.reshape(N // unit, unit, -1)[window_index].reshape(N, -1)
    rot = rot.reshape(N // unit, unit, -1)[window_index].reshape(N, -1)
    emb = torch.cat((rot, rot), dim=-1)
    cos, sin = emb.cos().unsqueeze(-2), emb.sin().unsqueeze(-2)
    for l in range(cfg["v_depth"]):
        b = "%sblocks.%d." % (p, l)
        cu = cu_full if l in cfg["fullatt"] else cu_window
        h = _rms(x, sd[b + "norm1.weight"], 1e-6)
        qkv = F.linear(h, sd[b + "attn.qkv.weight"].to(dt), sd[b + "attn.qkv.bias"].to(dt)).reshape(N, 3, heads, hd)
        q, k, v = qkv[:, 0], qkv[:, 1], qkv[:, 2]
        q = (q.float() * cos + _rot_half(q.float()) * sin).to(dt)
        k = (k.float() * cos + _rot_half(k.float()) * sin).to(dt)
        outs = []
        for s0, s1 in zip(cu[:-1].tolist(), cu[1:].tolist()):
            qq, kk, vv = (t[s0:s1].transpose(0, 1) for t in (q, k, v))
            a = torch.softmax((qq @ kk.transpose(-1, -2)) * hd ** -0.5, dim=-1, dtype=torch.float32).to(dt)
            outs.append((a @ vv).transpose(0, 1))
        a = torch.cat(outs, dim=0).reshape(N, Hv)
        x = x + F.linear(a, sd[b + "attn.proj.weight"].to(dt), sd[b + "attn.proj.bias"].to(dt))
        h = _rms(x, sd[b + "norm2.weight"], 1e-6)
        g = F.linear(h, sd[b + "mlp.gate_proj.weight"].to(dt), sd[b + "mlp.gate_proj.bias"].to(dt))
        u = F.linear(h, sd[b + "mlp.up_proj.weight"].to(dt), sd[b + "mlp.up_proj.bias"].to(dt))
        x = x + F.linear(F.silu(g) * u, sd[b + "mlp.down_proj.weight"].to(dt), sd[b + "mlp.down_proj.bias"].to(dt))
    m = _rms(x, sd[p + "merger.ln_q.weight"], 1e-6).view(-1, Hv * unit)
    m = F.linear(F.gelu(F.linear(m, sd[p + "merger.mlp.0.weight"].to(dt), sd[p + "merger.mlp.0.bias"].to(dt))),
                 sd[p + "merger.mlp.2.weight"].to(dt), sd[p + "merger.mlp.2.bias"].to(dt))
    return m[torch.argsort(window_index)]


# ------------------------------------------------------------------------------------------------ decoder
def text_forward(sd, cfg, inputs_embeds, position_ids, p="model.", key_mask=None):
    """Qwen2_5_VLTextModel.forward, full causal attention, no cache; returns the final-norm hidden states [B,S,H].
    `key_mask` [B,S] bool (the 2-D attention_mask of a padded batch): masked keys are invisible to every query."""
    x = inputs_embeds
    dt = x.dtype
    B, S, H = x.shape
    heads, kvh, hd = cfg["heads"], cfg["kv_heads"], cfg["head_dim"]
    inv_freq = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2, dtype=torch.int64, device=x.device).float() / hd))
    freqs = position_ids[:, :, :, None].float() * inv_freq[None, None, None, :]  # [3,B,S,hd/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    cos, sin = emb.cos().to(dt), emb.sin().to(dt)
    sec = cfg["mrope"] * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(cos.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    sin = torch.cat([m[i % 3] for i, m in enumerate(sin.split(sec, dim=-1))], dim=-1).unsqueeze(1)
    mask = torch.triu(torch.ones(S, S, dtype=torch.bool, device=x.device), diagonal=1)
    for l in range(cfg["layers"]):
        b = "%slayers.%d." % (p, l)
        h = _rms(x, sd[b + "input_layernorm.weight"], cfg["rms_eps"])
        q = F.linear(h, sd[b + "self_attn.q_proj.weight"].to(dt), sd[b + "self_attn.q_proj.bias"].to(dt))
        k = F.linear(h, sd[b + "self_attn.k_proj.weight"].to(dt), sd[b + "self_attn.k_proj.bias"].to(dt))
        v = F.linear(h, sd[b + "self_attn.v_proj.weight"].to(dt), sd[b + "self_attn.v_proj.bias"].to(dt))
        q = q.view(B, S, heads, hd).transpose(1, 2)
        k = k.view(B, S, kvh, hd).transpose(1, 2)
        v = v.view(B, S, kvh, hd).transpose(1, 2)
        q = q * cos + _rot_half(q) * sin
        k = k * cos + _rot_half(k) * sin
        k = k.repeat_interleave(heads // kvh, dim=1)
        v = v.repeat_interleave(heads // kvh, dim=1)
        s = (q @ k.transpose(2, 3)) * hd ** -0.5
        s = s.masked_fill(mask, float("-inf"))
        if key_mask is not None:
            s = s.masked_fill(~key_mask.to(s.device)[:, None, None, :], float("-inf"))
            s = s.masked_fill(~key_mask.to(s.device)[:, None, :, None], 0.0)  # padded queries: any finite row will do
        a = torch.softmax(s, dim=-1, dtype=torch.float32).to(dt) @ v
        a = a.transpose(1, 2).reshape(B, S, heads * hd)
        x = x + F.linear(a, sd[b + "self_attn.o_proj.weight"].to(dt))
        h = _rms(x, sd[b + "post_attention_layernorm.weight"], cfg["rms_eps"])
        g = F.linear(h, sd[b + "mlp.gate_proj.weight"].to(dt))
        u = F.linear(h, sd[b + "mlp.up_proj.weight"].to(dt))
        x = x + F.linear(F.silu(g) * u, sd[b + "mlp.down_proj.weight"].to(dt))
    return _rms(x, sd[p + "norm.weight"], cfg["rms_eps"])


def generate_latents(sd, cfg, input_ids, pixel_values, image_grid_thw):
    """InternVLAN1ForCausalLM.generate_latents (internvla_n1.py L320-347) for one prompt: input_ids [1,S]."""
    dt = pixel_values.dtype
    emb = sd["model.embed_tokens.weight"]
    text = emb[input_ids.to(emb.device)].to(dt)
    lat = sd["model.latent_queries"].to(dt).repeat(text.shape[0], 1, 1)
    image_idx = input_ids == IMAGE_TOKEN_INDEX
    nq = cfg["n_query"]
    ids = torch.cat([input_ids, torch.tensor([[TRAJ_TOKEN_INDEX] * nq])], dim=1)
    img = vit_forward(sd, cfg, pixel_values, image_grid_thw).unsqueeze(0)
    text[image_idx.to(text.device)] = img[0, : int(image_idx.sum())]
    text = torch.cat([text, lat], dim=1)
    pos, _ = rope_index(ids, image_grid_thw, cfg["v_merge"])
    hs = text_forward(sd, cfg, text, pos.to(text.device))
    return hs[:, -nq:, :]


def next_token_logits(sd, cfg, input_ids, image_feats, image_grid_thw):
    """Logits of the token after `input_ids` [1,S]: Qwen2_5_VLForConditionalGeneration.forward without a cache (embed,
    splice image features, get_rope_index, decoder, lm_head on the last position).  Returns fp32 [vocab]."""
    emb = sd["model.embed_tokens.weight"]
    dt = image_feats.dtype
    text = emb[input_ids.to(emb.device)].to(dt)
    image_idx = input_ids == IMAGE_TOKEN_INDEX
    text[image_idx.to(text.device)] = image_feats[: int(image_idx.sum())]
    pos, _ = rope_index(input_ids, image_grid_thw, cfg["v_merge"])
    hs = text_forward(sd, cfg, text, pos.to(text.device))
    return F.linear(hs[0, -1], sd["lm_head.weight"].to(dt)).float()


def greedy_generate(sd, cfg, input_ids, pixel_values, image_grid_thw, max_new_tokens=128, eos_token_ids=(151645, 151643),
                    return_logits=False):
    """`model.generate(**inputs, max_new_tokens=..., do_sample=False)` for one prompt (internvla_n1_policy.py L169-176):
    GenerationMixin greedy search -- next = argmax(logits[:, -1]); the eos id is appended, then generation stops.  No KV
    cache here: every step re-runs the full sequence, which is the same function of the same inputs.  Returns the list
    of generated ids (and, optionally, the fp32 logits each one was chosen from)."""
    feats = vit_forward(sd, cfg, pixel_values, image_grid_thw)
    ids = input_ids.clone()
    out, logs = [], []
    for _ in range(max_new_tokens):
        lg = next_token_logits(sd, cfg, ids, feats, image_grid_thw)
        tok = int(torch.argmax(lg))
        out.append(tok)
        logs.append(lg)
        ids = torch.cat([ids, torch.tensor([[tok]], dtype=ids.dtype)], dim=1)
        if tok in eos_token_ids:
            break
    return (out, logs) if return_logits else out


def training_traj_states(sd, cfg, input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos):
    """The System-2 half of the training forward (internvla_n1.py L128-235) on a collated batch
    (internvla_n1_lerobot_dataset.py L1155-1277): input_ids [B,S] right-padded, each sample ending with n_query TRAJ
    tokens at t_s_pos[b]; attention_mask = input_ids != pad.  Embeds, splices image features (masked_scatter order) and
    `latent_queries` at the TRAJ positions (L166-172), get_rope_index with the mask (masked positions keep 1), runs the
    decoder over the padded batch and gathers hidden[b, t_s_pos[b] : t_s_pos[b] + n_query].  -> [B, n_query, H]"""
    dt = pixel_values.dtype
    emb = sd["model.embed_tokens.weight"]
    x = emb[input_ids.to(emb.device)].to(dt)
    feats = vit_forward(sd, cfg, pixel_values, image_grid_thw)
    x[(input_ids == IMAGE_TOKEN_INDEX).to(x.device)] = feats
    nq = cfg["n_query"]
    lat = sd["model.latent_queries"].to(dt).reshape(nq, -1)
    traj_idx = input_ids == TRAJ_TOKEN_INDEX
    x[traj_idx.to(x.device)] = lat.repeat(input_ids.shape[0], 1)
    B, S = input_ids.shape
    pos = torch.ones(3, B, S, dtype=torch.long)
    img = 0
    grids = torch.as_tensor(image_grid_thw).reshape(-1, 3)
    for b in range(B):
        keep = attention_mask[b].bool()
        ids_b = input_ids[b][keep].unsqueeze(0)
        n_img = int(((ids_b[0, :-1] == VISION_START) & (ids_b[0, 1:] == IMAGE_TOKEN_INDEX)).sum())
        pb, _ = rope_index(ids_b, grids[img:img + n_img], cfg["v_merge"])
        img += n_img
        pos[:, b, keep] = pb[:, 0]
    hs = text_forward(sd, cfg, x, pos.to(x.device), key_mask=attention_mask.bool())
    return torch.stack([hs[b, t_s_pos[b]: t_s_pos[b] + nq] for b in range(B)])


def latent_query_grads(sd, cfg, input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos, grad_states):
    """Backward of the System-2 half of the training step (row a13): the LLM is frozen, the only trainable tensor on
    this side is `model.latent_queries` (internvla_n1_arch.py L123; written into the TRAJ positions at internvla_n1.py
    L166-172).  Given d loss / d traj_hidden_states [B, n_query, H] (from navdp_oracle.s1_training_grads) returns
    d loss / d latent_queries [1, n_query, H] by autograd through `training_traj_states`."""
    leaf = sd["model.latent_queries"].detach().clone().requires_grad_(True)
    full = dict(sd)
    full["model.latent_queries"] = leaf
    hs = training_traj_states(full, cfg, input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos)
    (hs * grad_states.to(hs.dtype)).sum().backward()
    return leaf.grad


# ------------------------------------------------------------------------------------------------ synthetic weights
def s2_shapes(cfg, lm_head=False):
    Hv, H = cfg["v_hidden"], cfg["hidden"]
    unit = cfg["v_merge"] ** 2
    out = [("visual.patch_embed.proj.weight", (Hv, 3, cfg["v_tpatch"], cfg["v_patch"], cfg["v_patch"]))]
    for i in range(cfg["v_depth"]):
        b = "visual.blocks.%d." % i
        out += [(b + "norm1.weight", (Hv,)), (b + "norm2.weight", (Hv,)), (b + "attn.qkv.weight", (3 * Hv, Hv)),
                (b + "attn.qkv.bias", (3 * Hv,)), (b + "attn.proj.weight", (Hv, Hv)), (b + "attn.proj.bias", (Hv,)),
                (b + "mlp.gate_proj.weight", (cfg["v_inter"], Hv)), (b + "mlp.gate_proj.bias", (cfg["v_inter"],)),
                (b + "mlp.up_proj.weight", (cfg["v_inter"], Hv)), (b + "mlp.up_proj.bias", (cfg["v_inter"],)),
                (b + "mlp.down_proj.weight", (Hv, cfg["v_inter"])), (b + "mlp.down_proj.bias", (Hv,))]
    out += [("visual.merger.ln_q.weight", (Hv,)), ("visual.merger.mlp.0.weight", (Hv * unit, Hv * unit)),
            ("visual.merger.mlp.0.bias", (Hv * unit,)), ("visual.merger.mlp.2.weight", (cfg["v_out"], Hv * unit)),
            ("visual.merger.mlp.2.bias", (cfg["v_out"],))]
    out += [("model.embed_tokens.weight", (cfg["vocab"], H)), ("model.latent_queries", (1, cfg["n_query"], H))]
    qd, kd = cfg["heads"] * cfg["head_dim"], cfg["kv_heads"] * cfg["head_dim"]
    for i in range(cfg["layers"]):
        b = "model.layers.%d." % i
        out += [(b + "input_layernorm.weight", (H,)), (b + "post_attention_layernorm.weight", (H,)),
                (b + "self_attn.q_proj.weight", (qd, H)), (b + "self_attn.q_proj.bias", (qd,)),
                (b + "self_attn.k_proj.weight", (kd, H)), (b + "self_attn.k_proj.bias", (kd,)),
                (b + "self_attn.v_proj.weight", (kd, H)), (b + "self_attn.v_proj.bias", (kd,)),
                (b + "self_attn.o_proj.weight", (H, qd)), (b + "mlp.gate_proj.weight", (cfg["inter"], H)),
                (b + "mlp.up_proj.weight", (cfg["inter"], H)), (b + "mlp.down_proj.weight", (H, cfg["inter"]))]
    out += [("model.norm.weight", (H,))]
    if lm_head:
        out += [("lm_head.weight", (cfg["vocab"], H))]
    return out


def make_s2_state_dict(cfg, seed=0, device="cpu", dtype=torch.float32, vocab_rows=None, lm_head=False):
    """Seeded synthetic weights (numpy PCG64, reproducible across machines).  `vocab_rows` truncates the random part of
    the embedding table for big configs (rows beyond it repeat) to keep generation fast."""
    import zlib
    sd = {}
    for name, shape in s2_shapes(cfg, lm_head=lm_head):
        rng = np.random.Generator(np.random.PCG64([seed, zlib.crc32(name.encode())]))
        if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name.endswith("layernorm.weight") \
                or name.endswith("norm.weight") or name.endswith("ln_q.weight"):
            a = 1.0 + 0.1 * rng.standard_normal(shape, dtype=np.float32)
        elif name.endswith(".bias"):
            a = 0.02 * rng.standard_normal(shape, dtype=np.float32)
        elif name == "model.embed_tokens.weight" and vocab_rows and vocab_rows < shape[0]:
            base = rng.standard_normal((vocab_rows, shape[1]), dtype=np.float32)
            a = np.tile(base, (shape[0] // vocab_rows + 1, 1))[: shape[0]]
        elif name in ("model.embed_tokens.weight", "model.latent_queries"):
            a = rng.standard_normal(shape, dtype=np.float32)
        else:
            fan_in = int(np.prod(shape[1:]))
            a = rng.standard_normal(shape, dtype=np.float32) / np.float32(math.sqrt(fan_in))
        sd[name] = torch.from_numpy(np.ascontiguousarray(a.astype(np.float32))).to(device=device, dtype=dtype)
    return sd


def make_prompt(rng, n_text_pre, grids, n_text_post, merge=2, vocab_text=151643):
    """Token ids of one prompt: text, then per image <vision_start> + image pads (+ <vision_end> 151653), then text."""
    ids = rng.integers(0, vocab_text, n_text_pre).tolist()
    for t, h, w in grids:
        ids += [VISION_START] + [IMAGE_TOKEN_INDEX] * (t * h * w // (merge * merge)) + [151653]
    ids += rng.integers(0, vocab_text, n_text_post).tolist()
    return ids
