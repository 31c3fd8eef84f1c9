"""Hand-written backward of the System-2 half of the training step -- TEST INFRASTRUCTURE: the executable
specification of the kernels that carry d loss / d traj_hidden_states back to `latent_queries`.

The decoder is frozen and causal, and `latent_queries` only enter at the n_query TRAJ positions of each sample, so
  * no weight gradient exists on this side,
  * the hidden states of all other positions do not depend on `latent_queries`: their keys / values are constants,
  * the backward therefore runs on the TRAJ rows alone (n_query rows per sample and layer) against the per-layer K / V
    cache the forward already holds (the decode path's slotted cache): RMSNorm, SwiGLU, o_proj / q / k / v dgrads,
    rotate-half RoPE transposed, and softmax-attention backward of n_query queries over the visible keys, where only
    the TRAJ columns of dK / dV are kept.
No autograd in this file; tests/test_oracle_s2.py compares with `qwen_oracle.latent_query_grads` (autograd), which is
pinned to transformers' autograd.
"""
import torch
import torch.nn.functional as F

from . import qwen_oracle as Q


def _rms_fwd(x, w, eps):
    r = torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps)
    return w * (x * r), r


def _rms_bwd(x, w, r, dy):
    dxh = dy * w
    return r * (dxh - x * r * r * (dxh * x).mean(-1, keepdim=True))


def latent_query_backward(sd, cfg, input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos, grad_states):
    """Same contract as qwen_oracle.latent_query_grads: -> d loss / d latent_queries [1, n_query, H]."""
    nq, heads, kvh, hd, eps = cfg["n_query"], cfg["heads"], cfg["kv_heads"], cfg["head_dim"], cfg["rms_eps"]
    rep = heads // kvh
    B, S = input_ids.shape
    # ---- forward over the padded batch (what the prefill computes anyway), keeping per layer: the TRAJ-row inputs and
    #      the rotated K / V of every position
    emb = sd["model.embed_tokens.weight"]
    x = emb[input_ids].float()
    x[input_ids == Q.IMAGE_TOKEN_INDEX] = Q.vit_forward(sd, cfg, pixel_values, image_grid_thw)
    x[input_ids == Q.TRAJ_TOKEN_INDEX] = sd["model.latent_queries"].reshape(nq, -1).repeat(B, 1)
    pos = torch.ones(3, B, S, dtype=torch.long)
    grids = torch.as_tensor(image_grid_thw).reshape(-1, 3)
    img = 0
    for b in range(B):
        keep = attention_mask[b].bool()
        ids_b = input_ids[b][keep].unsqueeze(0)
        n_img = int(((ids_b[0, :-1] == Q.VISION_START) & (ids_b[0, 1:] == Q.IMAGE_TOKEN_INDEX)).sum())
        pb, _ = Q.rope_index(ids_b, grids[img:img + n_img], cfg["v_merge"])
        img += n_img
        pos[:, b, keep] = pb[:, 0]
    inv_freq = 1.0 / (cfg["rope_theta"] ** (torch.arange(0, hd, 2).float() / hd))
    fr = pos[:, :, :, None].float() * inv_freq
    e = torch.cat((fr, fr), dim=-1)
    sec = cfg["mrope"] * 2
    cos = torch.cat([m[i % 3] for i, m in enumerate(e.cos().split(sec, dim=-1))], dim=-1)      # [B, S, hd]
    sin = torch.cat([m[i % 3] for i, m in enumerate(e.sin().split(sec, dim=-1))], dim=-1)
    rows = torch.stack([torch.arange(t, t + nq) for t in t_s_pos])                               # [B, nq]
    bidx = torch.arange(B)[:, None]
    causal = torch.triu(torch.ones(S, S, dtype=torch.bool), diagonal=1)
    kmask = ~attention_mask.bool()
    tape = []
    for l in range(cfg["layers"]):
        p = "model.layers.%d." % l
        h, r1 = _rms_fwd(x, sd[p + "input_layernorm.weight"], eps)
        q = F.linear(h, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]).view(B, S, heads, hd)
        k = F.linear(h, sd[p + "self_attn.k_proj.weight"], sd[p + "self_attn.k_proj.bias"]).view(B, S, kvh, hd)
        v = F.linear(h, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"]).view(B, S, kvh, hd)
        c, s_ = cos[:, :, None], sin[:, :, None]
        q = q * c + Q._rot_half(q) * s_
        k = k * c + Q._rot_half(k) * s_
        kk, vv = k.repeat_interleave(rep, dim=2), v.repeat_interleave(rep, dim=2)
        sc = torch.einsum("bqhd,bkhd->bhqk", q, kk) * hd ** -0.5
        sc = sc.masked_fill(causal, float("-inf")).masked_fill(kmask[:, None, None, :], float("-inf"))
        sc = sc.masked_fill(kmask[:, None, :, None], 0.0)
        pr = sc.softmax(-1)
        a = torch.einsum("bhqk,bkhd->bqhd", pr, vv).reshape(B, S, heads * hd)
        x_in = x
        x = x + F.linear(a, sd[p + "self_attn.o_proj.weight"])
        h2, r2 = _rms_fwd(x, sd[p + "post_attention_layernorm.weight"], eps)
        gt, up = F.linear(h2, sd[p + "mlp.gate_proj.weight"]), F.linear(h2, sd[p + "mlp.up_proj.weight"])
        x_mid = x
        x = x + F.linear(F.silu(gt) * up, sd[p + "mlp.down_proj.weight"])
        # what the backward reads: TRAJ-row slices only, plus the K / V cache of this layer
        tape.append(dict(x_in=x_in[bidx, rows], r1=r1[bidx, rows], q=q[bidx, rows], a=a[bidx, rows], k=k, v=v,
                         p=pr[bidx, :, rows], x_mid=x_mid[bidx, rows], r2=r2[bidx, rows], gt=gt[bidx, rows],
                         up=up[bidx, rows]))
    xf, rf = x[bidx, rows], torch.rsqrt(x[bidx, rows].pow(2).mean(-1, keepdim=True) + eps)

    # ---- backward on the TRAJ rows [B, nq, H]
    d = _rms_bwd(xf, sd["model.norm.weight"], rf, grad_states.float())
    cq, sq = cos[bidx, rows][:, :, None], sin[bidx, rows][:, :, None]                            # [B, nq, 1, hd]

    def rope_t(g):  # transpose of y = x c + rot_half(x) s :  x_bar = y_bar c - rot_half(y_bar s)
        return g * cq - Q._rot_half(g * sq)

    for l in reversed(range(cfg["layers"])):
        p = "model.layers.%d." % l
        t = tape[l]
        act = F.silu(t["gt"]) * t["up"]
        dact = d @ sd[p + "mlp.down_proj.weight"]
        sg = torch.sigmoid(t["gt"])
        dgt = dact * t["up"] * (sg * (1 + t["gt"] * (1 - sg)))
        dup = dact * F.silu(t["gt"])
        dh2 = dgt @ sd[p + "mlp.gate_proj.weight"] + dup @ sd[p + "mlp.up_proj.weight"]
        d = d + _rms_bwd(t["x_mid"], sd[p + "post_attention_layernorm.weight"], t["r2"], dh2)
        da = (d @ sd[p + "self_attn.o_proj.weight"]).view(B, nq, heads, hd)
        kk, vv = t["k"].repeat_interleave(rep, dim=2), t["v"].repeat_interleave(rep, dim=2)      # [B, S, heads, hd]
        pr = t["p"].permute(0, 2, 1, 3)                                                            # [B, heads, nq, S]
        dp = torch.einsum("bqhd,bkhd->bhqk", da, vv)
        ds = pr * (dp - (dp * pr).sum(-1, keepdim=True))
        dq = torch.einsum("bhqk,bkhd->bqhd", ds, kk) * hd ** -0.5
        dk_all = torch.einsum("bhqk,bqhd->bkhd", ds, t["q"]) * hd ** -0.5                         # [B, S, heads, hd]
        dv_all = torch.einsum("bhqk,bqhd->bkhd", pr, da)
        # only the TRAJ columns of dK / dV lead back to latent_queries; query heads of a GQA group share one kv head
        dk = dk_all[bidx, rows].view(B, nq, kvh, rep, hd).sum(3)
        dv = dv_all[bidx, rows].view(B, nq, kvh, rep, hd).sum(3)
        dq, dk = rope_t(dq), rope_t(dk)
        dh = (dq.reshape(B, nq, -1) @ sd[p + "self_attn.q_proj.weight"] + dk.reshape(B, nq, -1) @ sd[p + "self_attn.k_proj.weight"]
              + dv.reshape(B, nq, -1) @ sd[p + "self_attn.v_proj.weight"])
        d = d + _rms_bwd(t["x_in"], sd[p + "input_layernorm.weight"], t["r1"], dh)
    return d.sum(0, keepdim=True)
