"""Hand-written backward of the System-1 training loss -- TEST INFRASTRUCTURE: the executable specification of the
backward kernels of SURVEY.md §8 row a13 (not built yet).

Every forward primitive of oracle/navdp_oracle.py (linear, LayerNorm, exact GELU, ReLU, packed-projection multi-head
attention, layer scale, conv-as-GEMM patch embedding, bicubic position-table resample) gets an explicit backward with
the tensors it needs saved -- i.e. what a fused kernel has to keep or recompute -- and the model-level functions chain
them in reverse.  No autograd anywhere in this file.  tests/test_oracle_s1.py checks every parameter gradient and the
gradient w.r.t. the latent tokens against autograd through the restated forward (`navdp_oracle.s1_training_grads`), which
is itself pinned to the reference module's autograd (tests/golden/s1_training_reference.npz).

Reference lines: navdp.py L291-312 (forward_vlm_traj), L165-175 (sample_noise), L177-195 (decoder call),
navdp_backbone.py L79-99, L151-202, dinov2.py L180-232, L272-322, internvla_n1.py L287-303 (masked MSE).
"""
import math

import torch
import torch.nn.functional as F

from . import ddpm
from . import navdp_oracle as O


class Grads(dict):
    def add(self, name, g):
        self[name] = self[name] + g if name in self else g


# ------------------------------------------------------------------------------------------------ primitives
def lin_fwd(sd, p, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias")), x


def lin_bwd(sd, p, x, dy, g, need_dx=True):
    """dgrad: dx = dy W; wgrad: dW = dy^T x (contraction over all leading dims); db = column sums of dy."""
    dy2, x2 = dy.reshape(-1, dy.shape[-1]), x.reshape(-1, x.shape[-1])
    g.add(p + ".weight", dy2.t() @ x2)
    if (p + ".bias") in sd:
        g.add(p + ".bias", dy2.sum(0))
    return dy @ sd[p + ".weight"] if need_dx else None


def ln_fwd(sd, p, x, eps):
    mu = x.mean(-1, keepdim=True)
    rstd = torch.rsqrt(x.var(-1, unbiased=False, keepdim=True) + eps)
    xhat = (x - mu) * rstd
    return xhat * sd[p + ".weight"] + sd[p + ".bias"], (xhat, rstd)


def ln_bwd(sd, p, saved, dy, g):
    xhat, rstd = saved
    g.add(p + ".weight", (dy * xhat).reshape(-1, xhat.shape[-1]).sum(0))
    g.add(p + ".bias", dy.reshape(-1, xhat.shape[-1]).sum(0))
    dxh = dy * sd[p + ".weight"]
    return rstd * (dxh - dxh.mean(-1, keepdim=True) - xhat * (dxh * xhat).mean(-1, keepdim=True))


def gelu_bwd(x, dy):
    """d/dx [x Phi(x)] = Phi(x) + x phi(x)   (exact erf GELU, as F.gelu default)."""
    phi = torch.exp(-0.5 * x * x) / math.sqrt(2.0 * math.pi)
    return dy * (0.5 * (1.0 + torch.erf(x / math.sqrt(2.0))) + x * phi)


def attn_core_fwd(q, k, v, causal, scale):
    """q [B,h,Sq,d], k/v [B,h,Sk,d] -> o, saved probabilities."""
    s = (q @ k.transpose(-1, -2)) * scale
    if causal:
        s = s.masked_fill(torch.triu(torch.ones(s.shape[-2], s.shape[-1], dtype=torch.bool), diagonal=1), float("-inf"))
    p = s.softmax(-1)
    return p @ v, p


def attn_core_bwd(q, k, v, p, do, scale):
    dv = p.transpose(-1, -2) @ do
    dp = do @ v.transpose(-1, -2)
    ds = p * (dp - (dp * p).sum(-1, keepdim=True))      # softmax backward; masked entries have p = 0
    return (ds @ k) * scale, (ds.transpose(-1, -2) @ q) * scale, dv


def mha_fwd(sd, p, q_in, k_in, v_in, heads, causal=False):
    """nn.MultiheadAttention with packed in_proj (navdp_oracle._mha)."""
    D = q_in.shape[-1]
    w, b = sd[p + ".in_proj_weight"], sd[p + ".in_proj_bias"]
    q, k, v = F.linear(q_in, w[:D], b[:D]), F.linear(k_in, w[D:2 * D], b[D:2 * D]), F.linear(v_in, w[2 * D:], b[2 * D:])
    B, Sq, Sk, hd = q.shape[0], q.shape[1], k.shape[1], D // heads
    qh, kh, vh = (t.view(B, -1, heads, hd).transpose(1, 2) for t in (q, k, v))
    oh, prob = attn_core_fwd(qh, kh, vh, causal, 1.0 / math.sqrt(hd))
    o = oh.transpose(1, 2).reshape(B, Sq, D)
    y = F.linear(o, sd[p + ".out_proj.weight"], sd[p + ".out_proj.bias"])
    return y, (q_in, k_in, v_in, qh, kh, vh, prob, o)


def mha_bwd(sd, p, saved, dy, g, heads):
    q_in, k_in, v_in, qh, kh, vh, prob, o = saved
    D = q_in.shape[-1]
    hd = D // heads
    dy2 = dy.reshape(-1, D)
    g.add(p + ".out_proj.weight", dy2.t() @ o.reshape(-1, D))
    g.add(p + ".out_proj.bias", dy2.sum(0))
    do = (dy @ sd[p + ".out_proj.weight"]).view(dy.shape[0], -1, heads, hd).transpose(1, 2)
    dqh, dkh, dvh = attn_core_bwd(qh, kh, vh, prob, do, 1.0 / math.sqrt(hd))
    dq, dk, dv = (t.transpose(1, 2).reshape(t.shape[0], -1, D) for t in (dqh, dkh, dvh))
    w = sd[p + ".in_proj_weight"]
    dW = torch.cat([dq.reshape(-1, D).t() @ q_in.reshape(-1, D), dk.reshape(-1, D).t() @ k_in.reshape(-1, D),
                    dv.reshape(-1, D).t() @ v_in.reshape(-1, D)])
    g.add(p + ".in_proj_weight", dW)
    g.add(p + ".in_proj_bias", torch.cat([dq.reshape(-1, D).sum(0), dk.reshape(-1, D).sum(0), dv.reshape(-1, D).sum(0)]))
    return dq @ w[:D], dk @ w[D:2 * D], dv @ w[2 * D:]


# ------------------------------------------------------------------------------------------------ DINOv2 ViT-S (depth)
def _resample_matrix(n_src_side, n_dst_side, offset=0.1):
    """The bicubic position-table resample of dinov2.py L180-211 is linear in the table: R [dst^2, src^2]."""
    n = n_src_side * n_src_side
    eye = torch.eye(n).reshape(n, 1, n_src_side, n_src_side)
    s = float(n_dst_side + offset) / n_src_side
    out = F.interpolate(eye, scale_factor=(s, s), mode="bicubic", antialias=False)
    return out.reshape(n, -1).t()


def vit_fwd(sd, p, x):
    """navdp_oracle.dinov2_vits with everything a backward needs saved."""
    n = x.shape[0]
    Wp = sd[p + "patch_embed.proj.weight"]
    patches = F.unfold(x, kernel_size=14, stride=14).transpose(1, 2)                      # [n, 256, 3*14*14]
    t = patches @ Wp.reshape(Wp.shape[0], -1).t() + sd[p + "patch_embed.proj.bias"]
    side = int(math.isqrt(t.shape[1]))
    pe = sd[p + "pos_embed"]
    src_side = int(math.isqrt(pe.shape[1] - 1))
    R = None if src_side == side else _resample_matrix(src_side, side)
    pe_patch = pe[0, 1:] if R is None else R @ pe[0, 1:]
    t = torch.cat((sd[p + "cls_token"].expand(n, -1, -1), t), dim=1) + torch.cat((pe[:, :1], pe_patch.unsqueeze(0)), dim=1)
    tape = []
    heads, C = 6, t.shape[-1]
    for i in range(12):
        b = "%sblocks.%d." % (p, i)
        h, s1 = ln_fwd(sd, b + "norm1", t, 1e-6)
        qkv, _ = lin_fwd(sd, b + "attn.qkv", h)
        B, N = h.shape[:2]
        qkv_h = qkv.reshape(B, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
        scale = (C // heads) ** -0.5
        oh, prob = attn_core_fwd(qkv_h[0], qkv_h[1], qkv_h[2], False, scale)
        o = oh.transpose(1, 2).reshape(B, N, C)
        a, _ = lin_fwd(sd, b + "attn.proj", o)
        t = t + a * sd[b + "ls1.gamma"]
        h2, s2 = ln_fwd(sd, b + "norm2", t, 1e-6)
        f1, _ = lin_fwd(sd, b + "mlp.fc1", h2)
        f2, _ = lin_fwd(sd, b + "mlp.fc2", F.gelu(f1))
        t = t + f2 * sd[b + "ls2.gamma"]
        tape.append((s1, h, qkv_h, prob, o, a, s2, h2, f1, f2, scale))
    y, sn = ln_fwd(sd, p + "norm", t, 1e-6)
    return y[:, 1:], (patches, R, tape, sn, n)


def vit_bwd(sd, p, saved, dy, g):
    patches, R, tape, sn, n = saved
    heads = 6
    dt = ln_bwd(sd, p + "norm", sn, torch.cat((torch.zeros_like(dy[:, :1]), dy), dim=1), g)
    C = dt.shape[-1]
    for i in reversed(range(12)):
        b = "%sblocks.%d." % (p, i)
        s1, h, qkv_h, prob, o, a, s2, h2, f1, f2, scale = tape[i]
        g.add(b + "ls2.gamma", (dt * f2).reshape(-1, C).sum(0))
        df2 = dt * sd[b + "ls2.gamma"]
        dact = lin_bwd(sd, b + "mlp.fc2", F.gelu(f1), df2, g)
        dh2 = lin_bwd(sd, b + "mlp.fc1", h2, gelu_bwd(f1, dact), g)
        dt = dt + ln_bwd(sd, b + "norm2", s2, dh2, g)
        g.add(b + "ls1.gamma", (dt * a).reshape(-1, C).sum(0))
        da = dt * sd[b + "ls1.gamma"]
        do = lin_bwd(sd, b + "attn.proj", o, da, g)
        B, N = do.shape[:2]
        doh = do.view(B, N, heads, C // heads).transpose(1, 2)
        dq, dk, dv = attn_core_bwd(qkv_h[0], qkv_h[1], qkv_h[2], prob, doh, scale)
        dqkv = torch.stack((dq, dk, dv)).permute(1, 3, 0, 2, 4).reshape(B, N, 3 * C)
        dh = lin_bwd(sd, b + "attn.qkv", h, dqkv, g)
        dt = dt + ln_bwd(sd, b + "norm1", s1, dh, g)
    # token assembly: cls token, position table (through the resample), patch embedding (conv as GEMM)
    g.add(p + "cls_token", dt[:, :1].sum(0, keepdim=True))
    dpe_patch = dt[:, 1:].sum(0)
    dpe = torch.cat((dt[:, :1].sum(0), dpe_patch if R is None else R.t() @ dpe_patch), dim=0).unsqueeze(0)
    g.add(p + "pos_embed", dpe)
    dpatch = dt[:, 1:].reshape(-1, C)
    Wp = sd[p + "patch_embed.proj.weight"]
    g.add(p + "patch_embed.proj.weight", (dpatch.t() @ patches.reshape(-1, patches.shape[-1])).reshape(Wp.shape))
    g.add(p + "patch_embed.proj.bias", dpatch.sum(0))


# ------------------------------------------------------------------------------------------------ model pieces
def _post_layer_fwd(sd, p, x, mem, heads):
    a1, m1 = mha_fwd(sd, p + "self_attn", x, x, x, heads)
    x1, n1 = ln_fwd(sd, p + "norm1", x + a1, 1e-5)
    a2, m2 = mha_fwd(sd, p + "multihead_attn", x1, mem, mem, heads)
    x2, n2 = ln_fwd(sd, p + "norm2", x1 + a2, 1e-5)
    f1, _ = lin_fwd(sd, p + "linear1", x2)
    f2, _ = lin_fwd(sd, p + "linear2", F.relu(f1))
    x3, n3 = ln_fwd(sd, p + "norm3", x2 + f2, 1e-5)
    return x3, (m1, n1, m2, n2, x2, f1, n3)


def _post_layer_bwd(sd, p, saved, dy, g, heads):
    m1, n1, m2, n2, x2, f1, n3 = saved
    d = ln_bwd(sd, p + "norm3", n3, dy, g)
    dact = lin_bwd(sd, p + "linear2", F.relu(f1), d, g)
    dx2 = d + lin_bwd(sd, p + "linear1", x2, dact * (f1 > 0), g)
    d = ln_bwd(sd, p + "norm2", n2, dx2, g)
    dq, dk, dv = mha_bwd(sd, p + "multihead_attn", m2, d, g, heads)
    dx1, dmem = d + dq, dk + dv
    d = ln_bwd(sd, p + "norm1", n1, dx1, g)
    dq, dk, dv = mha_bwd(sd, p + "self_attn", m1, d, g, heads)
    return d + dq + dk + dv, dmem


def rgbd_fwd(sd, images, depths, frames=2, p="rgbd_encoder."):
    B, T = images.shape[:2]
    with torch.no_grad():  # the RGB tokens are detached in the reference (navdp_backbone.py L170-171)
        mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.bfloat16).float().reshape(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.bfloat16).float().reshape(1, 3, 1, 1)
        ti = images.permute(0, 1, 4, 2, 3).reshape(-1, 3, 224, 224)
        image_token = O.dinov2_vits(sd, p + "rgb_model.", (ti - mean) / std).reshape(B, T * 256, -1)
    td = depths.permute(0, 1, 4, 2, 3).reshape(-1, 1, 224, 224)
    dtok, vsave = vit_fwd(sd, p + "depth_model.", torch.cat([td, td, td], dim=1))
    token = torch.cat((image_token, dtok.reshape(B, T * 256, -1)), dim=1) + sd[p + "former_pe.weight"][: frames * 512]
    x = sd[p + "former_query.weight"][: frames * 16].unsqueeze(0).expand(B, -1, -1)
    tape = []
    for i in range(2):
        x, s = _post_layer_fwd(sd, "%sformer_net.layers.%d." % (p, i), x, token, 8)
        tape.append(s)
    y, _ = lin_fwd(sd, p + "project_layer", x)
    return y, (vsave, tape, x, B, T, frames)


def rgbd_bwd(sd, saved, dy, g, p="rgbd_encoder."):
    vsave, tape, x_last, B, T, frames = saved
    d = lin_bwd(sd, p + "project_layer", x_last, dy, g)
    dtoken = 0
    for i in reversed(range(2)):
        d, dm = _post_layer_bwd(sd, "%sformer_net.layers.%d." % (p, i), tape[i], d, g, 8)
        dtoken = dtoken + dm
    gq = torch.zeros_like(sd[p + "former_query.weight"])
    gq[: frames * 16] = d.sum(0)
    g.add(p + "former_query.weight", gq)
    gpe = torch.zeros_like(sd[p + "former_pe.weight"])
    gpe[: frames * 512] = dtoken.sum(0)
    g.add(p + "former_pe.weight", gpe)
    ddepth = dtoken[:, T * 256:].reshape(B * T, 256, -1)      # the first T*256 tokens are the detached RGB ones
    vit_bwd(sd, p + "depth_model.", vsave, ddepth, g)


def goal_fwd(sd, vlm_tokens):
    h0, _ = lin_fwd(sd, "vlm_embed_mlp.0", vlm_tokens)
    h1, _ = lin_fwd(sd, "vlm_embed_mlp.2", F.relu(h0))
    h2, _ = lin_fwd(sd, "vlm_embed_mlp.4", F.relu(h1))
    B, n, _ = h2.shape
    c = "goal_compressor."
    x = h2 + sd[c + "token_positional_encoding.position_embedding.weight"][:n]
    q = sd[c + "target_embedding.weight"].unsqueeze(0).expand(B, -1, -1)
    q = q + sd[c + "query_positional_encoding.position_embedding.weight"][: q.shape[1]]
    y, m = mha_fwd(sd, c + "cross_attention", q, x, x, 8)
    return y, (vlm_tokens, h0, h1, m, n)


def goal_bwd(sd, saved, dy, g):
    vlm_tokens, h0, h1, m, n = saved
    c = "goal_compressor."
    dq, dk, dv = mha_bwd(sd, c + "cross_attention", m, dy, g, 8)
    nq = dq.shape[1]
    g.add(c + "target_embedding.weight", dq.sum(0))
    gqp = torch.zeros_like(sd[c + "query_positional_encoding.position_embedding.weight"])
    gqp[:nq] = dq.sum(0)
    g.add(c + "query_positional_encoding.position_embedding.weight", gqp)
    dx = dk + dv
    gtp = torch.zeros_like(sd[c + "token_positional_encoding.position_embedding.weight"])
    gtp[:n] = dx.sum(0)
    g.add(c + "token_positional_encoding.position_embedding.weight", gtp)
    d = lin_bwd(sd, "vlm_embed_mlp.4", F.relu(h1), dx, g) * (h1 > 0)
    d = lin_bwd(sd, "vlm_embed_mlp.2", F.relu(h0), d, g) * (h0 > 0)
    return lin_bwd(sd, "vlm_embed_mlp.0", vlm_tokens, d, g)


def decoder_fwd(sd, noisy, timestep, goal, rgbd, layers=16, heads=8):
    R, T, _ = noisy.shape
    B = goal.shape[0]
    Ns = R // B
    x, _ = lin_fwd(sd, "input_embed", noisy)
    time_emb = O.sinusoidal_pos_emb(timestep).unsqueeze(1)
    M = 2 + rgbd.shape[1]
    cond = (torch.cat([time_emb, goal, rgbd], dim=1) + sd["cond_pos_embed"][:, :M]).repeat_interleave(Ns, dim=0)
    x = x + sd["out_pos_embed"][:, :T]
    tape = []
    for i in range(layers):
        p = "decoder.layers.%d." % i
        h1, n1 = ln_fwd(sd, p + "norm1", x, 1e-5)
        a1, m1 = mha_fwd(sd, p + "self_attn", h1, h1, h1, heads, causal=True)
        x = x + a1
        h2, n2 = ln_fwd(sd, p + "norm2", x, 1e-5)
        a2, m2 = mha_fwd(sd, p + "multihead_attn", h2, cond, cond, heads)
        x = x + a2
        h3, n3 = ln_fwd(sd, p + "norm3", x, 1e-5)
        f1, _ = lin_fwd(sd, p + "linear1", h3)
        f2, _ = lin_fwd(sd, p + "linear2", F.gelu(f1))
        x = x + f2
        tape.append((n1, m1, n2, m2, n3, h3, f1))
    hN, nN = ln_fwd(sd, "layernorm", x, 1e-5)
    y, _ = lin_fwd(sd, "action_head", hN)
    return y, (noisy, tape, nN, hN, B, Ns, M, T)


def decoder_bwd(sd, saved, dy, g, layers=16, heads=8):
    noisy, tape, nN, hN, B, Ns, M, T = saved
    dx = ln_bwd(sd, "layernorm", nN, lin_bwd(sd, "action_head", hN, dy, g), g)
    dcond = 0
    for i in reversed(range(layers)):
        p = "decoder.layers.%d." % i
        n1, m1, n2, m2, n3, h3, f1 = tape[i]
        dact = lin_bwd(sd, p + "linear2", F.gelu(f1), dx, g)
        dx = dx + ln_bwd(sd, p + "norm3", n3, lin_bwd(sd, p + "linear1", h3, gelu_bwd(f1, dact), g), g)
        dq, dk, dv = mha_bwd(sd, p + "multihead_attn", m2, dx, g, heads)
        dcond = dcond + dk + dv
        dx = dx + ln_bwd(sd, p + "norm2", n2, dq, g)
        dq, dk, dv = mha_bwd(sd, p + "self_attn", m1, dx, g, heads)
        dx = dx + ln_bwd(sd, p + "norm1", n1, dq + dk + dv, g)
    gop = torch.zeros_like(sd["out_pos_embed"])
    gop[:, :T] = dx.sum(0, keepdim=True)
    g.add("out_pos_embed", gop)
    lin_bwd(sd, "input_embed", noisy, dx, g, need_dx=False)
    dcond = dcond.reshape(B, Ns, M, -1).sum(1)                 # the Ns samples of an environment share its condition
    gcp = torch.zeros_like(sd["cond_pos_embed"])
    gcp[:, :M] = dcond.sum(0, keepdim=True)
    g.add("cond_pos_embed", gcp)
    return dcond[:, 1:2], dcond[:, 2:]                          # d goal, d rgbd (the time embedding has no parameter)


def s1_training_backward(sd, traj_hidden_states, traj_images, traj_depths, traj_poses, video_frame_num, noise, timesteps,
                         K=20):
    """Forward + hand-written backward of navdp_oracle.s1_training_loss.  -> (loss, Grads, d loss / d hidden states)"""
    sd = {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    Bb, f = traj_images.shape[:2]
    hs = traj_hidden_states.unsqueeze(1).repeat(1, f, 1, 1).flatten(0, 1)
    mask = (torch.arange(f).expand(Bb, f) < video_frame_num.unsqueeze(1)).flatten(0, 1)[:, None, None].float()
    cur_i, cur_d = traj_images.flatten(0, 1), traj_depths.flatten(0, 1)
    g_i = traj_images[:, 0:1].repeat(1, f, 1, 1, 1).flatten(0, 1)
    g_d = traj_depths[:, 0:1].repeat(1, f, 1, 1).flatten(0, 1)
    images_dp = torch.stack([g_i, cur_i], dim=1)
    depths_dp = torch.stack([g_d, cur_d], dim=1).unsqueeze(-1)
    goal, gsave = goal_fwd(sd, hs)
    noisy = ddpm.DDPMScheduler(num_train_timesteps=K).add_noise(traj_poses.flatten(0, 1), noise, timesteps)
    rgbd, rsave = rgbd_fwd(sd, images_dp, depths_dp)
    pred, dsave = decoder_fwd(sd, noisy, timesteps, goal, rgbd)
    err = pred - noise
    denom = mask.sum() * err.shape[1] * err.shape[2]
    loss = (err.square() * mask).sum() / denom
    g = Grads()
    dgoal, drgbd = decoder_bwd(sd, dsave, 2.0 * err * mask / denom, g)
    rgbd_bwd(sd, rsave, drgbd, g)
    dhs = goal_bwd(sd, gsave, dgoal, g)
    return loss, g, dhs.reshape(Bb, f, *dhs.shape[1:]).sum(1)
