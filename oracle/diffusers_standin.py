"""Stand-ins for the `diffusers==0.33.1` leaf modules that the reference's NextDiT files import -- TEST INFRASTRUCTURE.

`diffusers` is not installed in this image (and cannot be: no network), so the reference's
internnav/model/basemodel/internvla_n1/nextdit_traj.py and nextdit_crossattn_traj.py cannot be imported as they are.
`install()` registers a minimal `diffusers` package in sys.modules whose classes have the constructor signatures, the
parameter names and the forward semantics of the 0.33.1 release, written as nn.Modules.  With it the reference's OWN
classes (LuminaNextDiTBlock, LuminaNextDiT2DModel, NextDiTCrossAttn) import and run, which pins the block / model
wiring of oracle/nextdit_oracle.py against the reference source.  The leaves themselves remain a restatement of a
third-party dependency (parity-unpinned, see the oracle's header)."""
import math
import sys
import types
from dataclasses import dataclass

import torch
import torch.nn as nn
import torch.nn.functional as F


class RMSNorm(nn.Module):
    def __init__(self, dim, eps, elementwise_affine=True, bias=False):
        super().__init__()
        self.eps = eps
        self.weight = nn.Parameter(torch.ones(dim)) if elementwise_affine else None

    def forward(self, x):
        dt = x.dtype
        var = x.to(torch.float32).pow(2).mean(-1, keepdim=True)
        x = x * torch.rsqrt(var + self.eps)
        if self.weight is not None:
            if self.weight.dtype in (torch.float16, torch.bfloat16):
                x = x.to(self.weight.dtype)
            return x * self.weight
        return x.to(dt)


class LuminaRMSNormZero(nn.Module):
    def __init__(self, embedding_dim, norm_eps, norm_elementwise_affine):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear = nn.Linear(min(embedding_dim, 1024), 4 * embedding_dim, bias=True)
        self.norm = RMSNorm(embedding_dim, eps=norm_eps, elementwise_affine=norm_elementwise_affine)

    def forward(self, x, emb=None):
        emb = self.linear(self.silu(emb))
        scale_msa, gate_msa, scale_mlp, gate_mlp = emb.chunk(4, dim=1)
        return self.norm(x) * (1 + scale_msa[:, None]), gate_msa, scale_mlp, gate_mlp


class LuminaLayerNormContinuous(nn.Module):
    def __init__(self, embedding_dim, conditioning_embedding_dim, elementwise_affine=True, eps=1e-5, bias=True,
                 norm_type="layer_norm", out_dim=None):
        super().__init__()
        self.silu = nn.SiLU()
        self.linear_1 = nn.Linear(conditioning_embedding_dim, embedding_dim, bias=bias)
        self.norm = nn.LayerNorm(embedding_dim, eps, elementwise_affine, bias)
        self.linear_2 = nn.Linear(embedding_dim, out_dim, bias=bias) if out_dim is not None else None

    def forward(self, x, conditioning_embedding):
        scale = self.linear_1(self.silu(conditioning_embedding).to(x.dtype))
        x = self.norm(x) * (1 + scale)[:, None, :]
        return self.linear_2(x) if self.linear_2 is not None else x


class LuminaFeedForward(nn.Module):
    def __init__(self, dim, inner_dim, multiple_of=256, ffn_dim_multiplier=None):
        super().__init__()
        inner_dim = int(2 * inner_dim / 3)
        if ffn_dim_multiplier is not None:
            inner_dim = int(ffn_dim_multiplier * inner_dim)
        inner_dim = multiple_of * ((inner_dim + multiple_of - 1) // multiple_of)
        self.linear_1 = nn.Linear(dim, inner_dim, bias=False)
        self.linear_2 = nn.Linear(inner_dim, dim, bias=False)
        self.linear_3 = nn.Linear(dim, inner_dim, bias=False)

    def forward(self, x):
        a = self.linear_1(x)
        return self.linear_2(F.silu(a.float()).to(a.dtype) * self.linear_3(x))


class Attention(nn.Module):
    def __init__(self, query_dim, cross_attention_dim=None, heads=8, kv_heads=None, dim_head=64, bias=False, qk_norm=None,
                 eps=1e-5, out_bias=True, processor=None, **kw):
        super().__init__()
        self.heads, self.inner_dim = heads, dim_head * heads
        kv_heads = heads if kv_heads is None else kv_heads
        self.inner_kv_dim = dim_head * kv_heads
        self.scale = dim_head ** -0.5
        cross = query_dim if cross_attention_dim is None else cross_attention_dim
        self.to_q = nn.Linear(query_dim, self.inner_dim, bias=bias)
        self.to_k = nn.Linear(cross, self.inner_kv_dim, bias=bias)
        self.to_v = nn.Linear(cross, self.inner_kv_dim, bias=bias)
        assert qk_norm in (None, "layer_norm_across_heads")
        self.norm_q = nn.LayerNorm(dim_head * heads, eps=eps) if qk_norm else None
        self.norm_k = nn.LayerNorm(dim_head * kv_heads, eps=eps) if qk_norm else None
        self.to_out = nn.ModuleList([nn.Linear(self.inner_dim, query_dim, bias=out_bias), nn.Dropout(0.0)])
        self.processor = processor

    def forward(self, hidden_states, encoder_hidden_states=None, attention_mask=None, **kw):
        return self.processor(self, hidden_states, encoder_hidden_states=encoder_hidden_states, attention_mask=attention_mask, **kw)


class LuminaAttnProcessor2_0:
    def __call__(self, attn, hidden_states, encoder_hidden_states, attention_mask=None, query_rotary_emb=None,
                 key_rotary_emb=None, base_sequence_length=None):
        assert query_rotary_emb is None and key_rotary_emb is None, "the trajectory DiT passes no rotary embedding"
        B, S, _ = hidden_states.shape
        q, k, v = attn.to_q(hidden_states), attn.to_k(encoder_hidden_states), attn.to_v(encoder_hidden_states)
        hd = q.shape[-1] // attn.heads
        kvh = k.shape[-1] // hd
        dtype = q.dtype
        if attn.norm_q is not None:
            q = attn.norm_q(q)
        if attn.norm_k is not None:
            k = attn.norm_k(k)
        q, k, v = q.view(B, -1, attn.heads, hd), k.view(B, -1, kvh, hd), v.view(B, -1, kvh, hd)
        q, k = q.to(dtype), k.to(dtype)
        rep = attn.heads // kvh
        if rep >= 1:
            k = k.unsqueeze(3).repeat(1, 1, 1, rep, 1).flatten(2, 3)
            v = v.unsqueeze(3).repeat(1, 1, 1, rep, 1).flatten(2, 3)
        mask = attention_mask.bool().view(B, 1, 1, -1).expand(-1, attn.heads, S, -1)
        o = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), attn_mask=mask, scale=None)
        return o.transpose(1, 2).to(dtype)


class PixArtAlphaTextProjection(nn.Module):
    def __init__(self, in_features, hidden_size, out_features=None, act_fn="gelu_tanh"):
        super().__init__()
        self.linear_1 = nn.Linear(in_features, hidden_size, bias=True)
        self.act_1 = nn.GELU(approximate="tanh")
        self.linear_2 = nn.Linear(hidden_size, hidden_size if out_features is None else out_features, bias=True)

    def forward(self, caption):
        return self.linear_2(self.act_1(self.linear_1(caption)))


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels, time_embed_dim):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.act = nn.SiLU()
        self.linear_2 = nn.Linear(time_embed_dim, time_embed_dim)

    def forward(self, x):
        return self.linear_2(self.act(self.linear_1(x)))


class LuminaCombinedTimestepCaptionEmbedding(nn.Module):
    def __init__(self, hidden_size=4096, cross_attention_dim=2048, frequency_embedding_size=256):
        super().__init__()
        self.freq = frequency_embedding_size
        self.timestep_embedder = TimestepEmbedding(frequency_embedding_size, hidden_size)
        self.caption_embedder = nn.Sequential(nn.LayerNorm(cross_attention_dim), nn.Linear(cross_attention_dim, hidden_size, bias=True))

    def forward(self, timestep, caption_feat, caption_mask):
        half = self.freq // 2
        exponent = -math.log(10000) * torch.arange(half, dtype=torch.float32, device=timestep.device) / half
        emb = timestep[:, None].float() * torch.exp(exponent)[None, :]
        time_freq = torch.cat((torch.cos(emb), torch.sin(emb)), dim=-1)       # flip_sin_to_cos=True
        time_embed = self.timestep_embedder(time_freq.to(dtype=self.timestep_embedder.linear_1.weight.dtype))
        m = caption_mask.float().unsqueeze(-1)
        pool = ((caption_feat * m).sum(dim=1) / m.sum(dim=1)).to(caption_feat)
        return time_embed + self.caption_embedder(pool)


class LuminaPatchEmbed(nn.Module):   # constructed by the reference model, never called by its forward
    def __init__(self, patch_size=2, in_channels=4, embed_dim=768, bias=True):
        super().__init__()
        self.proj = nn.Linear(patch_size * patch_size * in_channels, embed_dim, bias=bias)


@dataclass
class Transformer2DModelOutput:
    sample: torch.Tensor


class ConfigMixin:
    pass


class ModelMixin(nn.Module):
    def enable_gradient_checkpointing(self):
        self.gradient_checkpointing = True


def register_to_config(fn):
    return fn


class FlowMatchEulerDiscreteScheduler:
    """Only what internvla_n1.py L360, L395-396, L427 uses: defaults (shift 1.0), set_timesteps(n, sigmas=...), step()."""
    def __init__(self, num_train_timesteps=1000, shift=1.0):
        self.num_train_timesteps, self.shift = num_train_timesteps, shift

    def set_timesteps(self, num_inference_steps=None, device=None, sigmas=None):
        import numpy as np
        sigmas = np.array(sigmas).astype(np.float32)
        sigmas = self.shift * sigmas / (1 + (self.shift - 1) * sigmas)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = sigmas * self.num_train_timesteps
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._i = 0

    def step(self, model_output, timestep, sample):
        sample = sample.to(torch.float32)
        prev = sample + (self.sigmas[self._i + 1] - self.sigmas[self._i]) * model_output
        self._i += 1
        return types.SimpleNamespace(prev_sample=prev.to(model_output.dtype))


def install():
    """Register the stand-in package (idempotent; extends the bare `diffusers` that oracle/ref_loader.py registers for the
    DDPM scheduler).  Refuses to shadow a real diffusers install."""
    real = sys.modules.get("diffusers")
    if real is not None and getattr(real, "__file__", None):
        raise RuntimeError("a real diffusers is importable: use it instead of the stand-in")

    def mod(name, **attrs):
        m = sys.modules.get(name)
        if m is None:
            m = types.ModuleType(name)
            sys.modules[name] = m
            parent, _, leaf = name.rpartition(".")
            if parent:
                setattr(sys.modules[parent], leaf, m)
        m.__dict__.update(attrs)
        m._n1_standin = True
        if not hasattr(m, "__path__"):
            m.__path__ = []
        return m

    quiet = types.SimpleNamespace(warning=lambda *a, **k: None, info=lambda *a, **k: None)
    logging = types.SimpleNamespace(get_logger=lambda name: quiet)
    mod("diffusers")
    mod("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
    mod("diffusers.models")
    mod("diffusers.models.attention", LuminaFeedForward=LuminaFeedForward)
    mod("diffusers.models.attention_processor", Attention=Attention, LuminaAttnProcessor2_0=LuminaAttnProcessor2_0)
    mod("diffusers.models.embeddings", LuminaCombinedTimestepCaptionEmbedding=LuminaCombinedTimestepCaptionEmbedding,
        LuminaPatchEmbed=LuminaPatchEmbed, PixArtAlphaTextProjection=PixArtAlphaTextProjection,
        get_2d_rotary_pos_embed_lumina=lambda *a, **k: None)
    mod("diffusers.models.modeling_outputs", Transformer2DModelOutput=Transformer2DModelOutput)
    mod("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    mod("diffusers.models.normalization", LuminaLayerNormContinuous=LuminaLayerNormContinuous,
        LuminaRMSNormZero=LuminaRMSNormZero, RMSNorm=RMSNorm)
    mod("diffusers.utils", is_torch_version=lambda *a, **k: True, logging=logging)
    mod("diffusers.schedulers", FlowMatchEulerDiscreteScheduler=FlowMatchEulerDiscreteScheduler)
