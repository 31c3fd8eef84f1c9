"""CPU restatement of the reference's NextDiT System 1 (`system1 = "nextdit_async"`, the released DualVLN default) --
TEST INFRASTRUCTURE, not product code.

Follows, line by line:
  * internnav/model/basemodel/internvla_n1/internvla_n1.py L349-432 (`generate_traj`, nextdit branch: condition tokens,
    CFG batch, flow-matching Euler loop),
  * internvla_n1_arch.py L76-118 (SinusoidalPositionalEncoding, MemoryEncoder, QFormer) and L131-145 (the modules),
  * nextdit_crossattn_traj.py L46-95 (NextDiTCrossAttn: 12 layers, dim 384, 6 heads, latent_embedding_size 768),
  * nextdit_traj.py L39-178 (LuminaNextDiTBlock.forward) and L296-368 (LuminaNextDiT2DModel.forward).

PARITY STATUS.  The block wiring above is the reference's own source and is pinned: tests/test_oracle_nextdit.py imports
the reference's nextdit_traj.py / nextdit_crossattn_traj.py classes (with stand-ins for the `diffusers` leaf modules,
oracle/diffusers_standin.py) and compares them with this file.  The LEAF modules live in an un-vendored third-party
dependency that is absent from this image -- `diffusers==0.33.1` (requirements/internvla_n1.txt) -- and are restated
from that release's published source: Attention + LuminaAttnProcessor2_0, LuminaFeedForward, LuminaRMSNormZero,
RMSNorm, LuminaLayerNormContinuous, LuminaCombinedTimestepCaptionEmbedding, PixArtAlphaTextProjection,
FlowMatchEulerDiscreteScheduler.  Those leaves are PARITY-UNPINNED (no diffusers install to run them against); the
stand-ins and this file restate them independently of each other's code paths only in the sense of two spellings of the
same published algorithm.

State-dict keys are the reference's attribute paths below `InternVLAN1ForCausalLM.model` (`cond_projector.0.weight`,
`rgb_model.blocks.0.attn.qkv.weight`, `memory_encoder.encoder.layers.0...`, `rgb_resampler.decoder.layers.0...`,
`traj_dit.model.layers.0.attn1.to_q.weight`, ...)."""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import navdp_oracle as N

DIM, HEADS, LAYERS, LATENT = 384, 6, 12, 768
RESNET_MEAN = (0.485, 0.456, 0.406)
RESNET_STD = (0.229, 0.224, 0.225)


def _w(sd, k, x):
    return sd[k].to(x.dtype)


def _lin(sd, p, x):
    b = sd.get(p + ".bias")
    return F.linear(x, _w(sd, p + ".weight", x), None if b is None else b.to(x.dtype))


def rms_norm(x, weight, eps):
    """diffusers.models.normalization.RMSNorm.forward: variance in fp32, cast to the weight dtype before the product."""
    var = x.float().pow(2).mean(-1, keepdim=True)
    y = x * torch.rsqrt(var + eps)
    if weight is not None:
        if weight.dtype in (torch.float16, torch.bfloat16):
            y = y.to(weight.dtype)
        y = y * weight
    return y.to(x.dtype) if weight is None else y


# ------------------------------------------------------------------------------------------------ condition tokens
def encoder_layer_post(sd, p, x, heads):
    """nn.TransformerEncoderLayer(d_model, nhead) defaults: post-norm, ReLU, dim_feedforward 2048, eps 1e-5
    (internvla_n1_arch.py L79-83)."""
    x = N._ln(sd, p + "norm1", x + N._mha(sd, p + "self_attn", x, x, x, heads), 1e-5)
    ff = N._lin(sd, p + "linear2", F.relu(N._lin(sd, p + "linear1", x)))
    return N._ln(sd, p + "norm2", x + ff, 1e-5)


def memory_encoder(sd, memory, p="memory_encoder."):
    """MemoryEncoder.forward, internvla_n1_arch.py L86-95: learned positions + 3 post-norm encoder layers (6 heads)."""
    x = memory + sd[p + "memory_pos"][: memory.shape[1]].to(memory.dtype).unsqueeze(0)
    for i in range(3):
        x = encoder_layer_post(sd, "%sencoder.layers.%d." % (p, i), x, 6)
    return x


def qformer(sd, visual_feats, p="rgb_resampler."):
    """QFormer.forward, internvla_n1_arch.py L112-118: 32 learned queries through 3 post-norm decoder layers (12 heads,
    d = 768); `visual_proj` is constructed but never applied."""
    B = visual_feats.shape[0]
    q = (sd[p + "query_tokens"] + sd[p + "query_pos"]).to(visual_feats.dtype).unsqueeze(0).expand(B, -1, -1)
    for i in range(3):
        q = N._decoder_layer_post(sd, "%sdecoder.layers.%d." % (p, i), q, visual_feats, 12)
    return q


def condition_tokens(sd, traj_latents, images_dp):
    """internvla_n1.py L363-382: [B, 4, 3584] latents + [B, 2, 224, 224, 3] frames -> [B, 36, 768].  The reference
    unflattens the ViT features with sizes=(1, -1) (one environment per call, L371-374); per environment that is the
    training branch's `sizes=(bsz, -1)` (L241-245), which is what a batch means here."""
    dtype = traj_latents.dtype
    lat = _lin(sd, "cond_projector.2", F.gelu(_lin(sd, "cond_projector.0", traj_latents), approximate="tanh"))
    B = images_dp.shape[0]
    img = images_dp.permute(0, 1, 4, 2, 3)
    mean = torch.tensor(RESNET_MEAN, dtype=torch.float32, device=img.device).view(1, 1, 3, 1, 1)
    std = torch.tensor(RESNET_STD, dtype=torch.float32, device=img.device).view(1, 1, 3, 1, 1)
    img = ((img - mean) / std).flatten(0, 1).to(dtype)
    feat = N.dinov2_vits(sd, "rgb_model.", img).unflatten(0, (B, -1)).flatten(1, 2)      # [B, 512, 384]
    mem = memory_encoder(sd, feat)
    mem = torch.cat((feat, mem), dim=-1)                                                     # [B, 512, 768]
    tokens = qformer(sd, mem)                                                                # [B, 32, 768]
    return torch.cat((tokens, lat), dim=1)


# ------------------------------------------------------------------------------------------------ DiT leaves (diffusers 0.33.1)
def timestep_sinusoid(t, dim=256, max_period=10000):
    """get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0.0, scale=1)."""
    half = dim // 2
    exponent = -math.log(max_period) * torch.arange(half, dtype=torch.float32, device=t.device) / half
    emb = t[:, None].float() * torch.exp(exponent)[None, :]
    return torch.cat((torch.cos(emb), torch.sin(emb)), dim=-1)


def time_caption_embed(sd, p, timestep, caption, mask):
    """LuminaCombinedTimestepCaptionEmbedding.forward: TimestepEmbedding(256 -> 384, SiLU) + Linear(LayerNorm(masked mean
    of the caption))."""
    tf = timestep_sinusoid(timestep).to(caption.dtype)
    te = _lin(sd, p + "timestep_embedder.linear_2", F.silu(_lin(sd, p + "timestep_embedder.linear_1", tf)))
    m = mask.float().unsqueeze(-1)
    pool = ((caption * m).sum(dim=1) / m.sum(dim=1)).to(caption.dtype)
    ce = _lin(sd, p + "caption_embedder.1", N._ln(sd, p + "caption_embedder.0", pool, 1e-5))
    return te + ce


def lumina_attention(sd, p, hidden, encoder, heads):
    """Attention(qk_norm="layer_norm_across_heads", bias=False) through LuminaAttnProcessor2_0 with no rotary embedding and
    an all-ones mask (nextdit_crossattn_traj.py L84-93 passes image_rotary_emb=None): q / k LayerNorm over the full
    width, softmax(q k^T / sqrt(hd)) v, result left as [B, S, heads, hd] (to_out is applied by the block)."""
    q = N._ln(sd, p + "norm_q", _lin(sd, p + "to_q", hidden), 1e-5)
    k = N._ln(sd, p + "norm_k", _lin(sd, p + "to_k", encoder), 1e-5)
    v = _lin(sd, p + "to_v", encoder)
    B, S, D = q.shape
    hd = D // heads
    q = q.view(B, S, heads, hd).transpose(1, 2)
    k = k.view(B, -1, heads, hd).transpose(1, 2)
    v = v.view(B, -1, heads, hd).transpose(1, 2)
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    return (s.softmax(-1) @ v).transpose(1, 2)


def dit_block(sd, p, x, enc, temb, heads=HEADS, eps=1e-5):
    """LuminaNextDiTBlock.forward, nextdit_traj.py L125-178."""
    mod = _lin(sd, p + "norm1.linear", F.silu(temb))                                       # LuminaRMSNormZero
    scale_msa, gate_msa, scale_mlp, gate_mlp = mod.chunk(4, dim=1)
    h = rms_norm(x, _w(sd, p + "norm1.norm.weight", x), eps) * (1 + scale_msa[:, None])
    self_out = lumina_attention(sd, p + "attn1.", h, h, heads)
    cross_out = lumina_attention(sd, p + "attn2.", h, rms_norm(enc, _w(sd, p + "norm1_context.weight", enc), eps), heads)
    cross_out = cross_out * _w(sd, p + "gate", x).tanh().view(1, 1, -1, 1)
    mixed = (self_out + cross_out).flatten(-2)
    hidden = _lin(sd, p + "attn2.to_out.0", mixed)
    x = x + gate_msa.unsqueeze(1).tanh() * rms_norm(hidden, _w(sd, p + "norm2.weight", x), eps)
    m = rms_norm(x, _w(sd, p + "ffn_norm1.weight", x), eps) * (1 + scale_mlp.unsqueeze(1))
    a = _lin(sd, p + "feed_forward.linear_1", m)
    ff = _lin(sd, p + "feed_forward.linear_2", F.silu(a.float()).to(a.dtype) * _lin(sd, p + "feed_forward.linear_3", m))
    return x + gate_mlp.unsqueeze(1).tanh() * rms_norm(ff, _w(sd, p + "ffn_norm2.weight", x), eps)


def traj_dit(sd, x, timestep, z_latents, p="traj_dit.model."):
    """NextDiTCrossAttn.forward (nextdit_crossattn_traj.py L84-95) -> LuminaNextDiT2DModel.forward (nextdit_traj.py
    L296-368): x [n, 32, 384], timestep [n], z_latents [n, 36, 768] -> [n, 32, 384]."""
    enc = _lin(sd, p + "caption_projection.linear_2",
               F.gelu(_lin(sd, p + "caption_projection.linear_1", z_latents), approximate="tanh"))
    mask = torch.ones(z_latents.shape[0], z_latents.shape[1], device=z_latents.device)
    temb = time_caption_embed(sd, p + "time_caption_embed.", timestep, enc, mask)
    for i in range(LAYERS):
        x = dit_block(sd, "%slayers.%d." % (p, i), x, enc, temb)
    # LuminaLayerNormContinuous(elementwise_affine=False, eps=1e-6, out_dim=384)
    scale = _lin(sd, p + "norm_out.linear_1", F.silu(temb).to(x.dtype))
    x = F.layer_norm(x, (x.shape[-1],), None, None, 1e-6) * (1 + scale)[:, None, :]
    return _lin(sd, p + "norm_out.linear_2", x)


# ------------------------------------------------------------------------------------------------ sampler
def flow_match_schedule(num_inference_steps=10, num_train_timesteps=1000):
    """FlowMatchEulerDiscreteScheduler() (shift 1.0, no dynamic shifting) after
    set_timesteps(n, sigmas=np.linspace(1.0, 1 / n, n)) (internvla_n1.py L395-396): float32 sigmas with a trailing 0,
    timesteps = sigmas * 1000 (float32)."""
    sig = np.linspace(1.0, 1 / num_inference_steps, num_inference_steps).astype(np.float32)
    sig = torch.from_numpy(sig).to(torch.float32)
    return sig * num_train_timesteps, torch.cat((sig, torch.zeros(1)))


def action_features(sd, latents):
    """action_encoder + SinusoidalPositionalEncoding(384) of the step index (internvla_n1.py L401-409,
    internvla_n1_arch.py L52-73: sin | cos halves, frequencies exp(-i ln(1e4) / half))."""
    T = latents.shape[1]
    half = DIM // 2
    freqs = torch.arange(T, dtype=torch.float32)[:, None] * torch.exp(
        -torch.arange(half, dtype=torch.float) * (torch.log(torch.tensor(10000.0)) / half))[None, :]
    pos = torch.cat((torch.sin(freqs), torch.cos(freqs)), dim=-1).to(latents.device)
    f = _lin(sd, "action_encoder", latents)
    return f + pos.to(f.dtype)[None]


def generate_traj(sd, traj_latents, images_dp, x_init, guidance_scale=1.0, num_inference_steps=10, num_sample_trajs=32):
    """internvla_n1.py L349-432.  x_init replaces `randn_tensor` (L389-394): [B * Ns, 32, 3] in traj_latents' dtype."""
    dtype = traj_latents.dtype
    hidden = condition_tokens(sd, traj_latents, images_dp)
    hidden_in = torch.cat((torch.zeros_like(hidden), hidden), 0).repeat_interleave(num_sample_trajs, dim=0)
    timesteps, sigmas = flow_match_schedule(num_inference_steps)
    latents = x_init.to(dtype)
    for i, t in enumerate(timesteps):
        feats = action_features(sd, latents)
        inp = feats.repeat(2, 1, 1)
        tt = t.unsqueeze(0).expand(inp.shape[0]).to(inp.device, torch.long)
        pred = _lin(sd, "action_decoder", traj_dit(sd, inp, tt, hidden_in))
        uncond, cond = pred.chunk(2)
        pred = uncond + guidance_scale * (cond - uncond)
        # FlowMatchEulerDiscreteScheduler.step: fp32 Euler update, cast back to the model dtype
        latents = (latents.to(torch.float32) + (sigmas[i + 1] - sigmas[i]).to(pred.device) * pred).to(pred.dtype)
    return latents
