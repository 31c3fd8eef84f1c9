"""Host-side pieces of the NextDiT System 1 (internnav_b200/nextdit.py) that run at load / call time on the CPU: the
flow-matching schedule, the position-table resampling and the algebraic foldings of constants into the packed weights --
each against the oracle's unfused arithmetic (oracle/nextdit_oracle.py, oracle/navdp_oracle.py) in fp32."""
import torch
import torch.nn.functional as F

from internnav_b200 import nextdit as N
from internnav_b200.manifest import random_nextdit_state_dict
from oracle import navdp_oracle as NO, nextdit_oracle as O


def _rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-12))


def test_schedule_equals_the_oracle():
    for n in (10, 4, 25):
        ts, sig = N.flow_match_schedule(n)
        ts_o, sig_o = O.flow_match_schedule(n)
        assert torch.equal(ts, ts_o.to(torch.long)) and torch.equal(sig, sig_o)
    assert N.flow_match_schedule(10)[0].tolist() == [1000, 900, 800, 700, 600, 500, 400, 300, 200, 100]


def test_position_table_resampling():
    pe = torch.randn(1, 37 * 37 + 1, 384, generator=torch.Generator().manual_seed(0))
    assert torch.allclose(N._resample_pos_embed(pe), NO.interpolate_pos_encoding(pe, 16), atol=1e-6)
    pe16 = torch.randn(1, 257, 384)
    assert torch.equal(N._resample_pos_embed(pe16), pe16)


def test_patch_embedding_fold():
    """im2col per channel (196 taps + 4 zero columns, the layout of n1_op_patchify_depth) @ W'^T + b' == the DINOv2 patch
    convolution applied to the ImageNet-normalised frame (internvla_n1.py L367-368, patch_embed.py L69-81)."""
    g = torch.Generator().manual_seed(1)
    Wp, bp = torch.randn(384, 3, 14, 14, generator=g) * 0.05, torch.randn(384, generator=g) * 0.1
    img = torch.rand(2, 224, 224, 3, generator=g)
    Wf, bf = N.fold_patch_embed(Wp, bp)
    assert Wf.shape == (384, 600) and bf.shape == (384,)
    cols = []
    for c in range(3):
        p = img[..., c].reshape(2, 16, 14, 16, 14).permute(0, 1, 3, 2, 4).reshape(2 * 256, 196)
        cols.append(F.pad(p, (0, 4)))
    mine = (torch.cat(cols, dim=1) @ Wf.t() + bf).view(2, 256, 384)
    mean = torch.tensor(N.RESNET_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(N.RESNET_STD).view(1, 3, 1, 1)
    ref = F.conv2d((img.permute(0, 3, 1, 2) - mean) / std, Wp, bp, stride=14).flatten(2).transpose(1, 2)
    assert _rel(mine, ref) < 1e-5


def test_cross_attention_and_head_folds():
    sd = random_nextdit_state_dict(2)
    p = "traj_dit.model.layers.3."
    g = torch.Generator().manual_seed(3)
    enc = torch.randn(5, 36, 384, generator=g)
    Wk, Wv = N.fold_cross_kv(sd[p + "attn2.to_k.weight"], sd[p + "attn2.to_v.weight"], sd[p + "norm1_context.weight"], sd[p + "gate"])
    e = enc * torch.rsqrt(enc.pow(2).mean(-1, keepdim=True) + 1e-5)                      # RMSNorm without its weight
    ctx = O.rms_norm(enc, sd[p + "norm1_context.weight"], 1e-5)                          # the reference's operand
    assert _rel(e @ Wk.t(), F.linear(ctx, sd[p + "attn2.to_k.weight"])) < 1e-5
    v_ref = F.linear(ctx, sd[p + "attn2.to_v.weight"]).view(5, 36, 6, 64) * sd[p + "gate"].tanh().view(1, 1, 6, 1)
    assert _rel((e @ Wv.t()).view(5, 36, 6, 64), v_ref) < 1e-5
    # softmax(.) (V * g) == (softmax(.) V) * g: the gate commutes with the attention average -- checked on one head
    a = torch.softmax(torch.randn(32, 36, generator=g), -1)
    assert torch.allclose(a @ v_ref[0, :, 2], (a @ (v_ref[0, :, 2] / sd[p + "gate"][2].tanh())) * sd[p + "gate"][2].tanh(), atol=1e-5)
    q = "traj_dit.model.norm_out.linear_2."
    Wh, bh = N.fold_head(sd[q + "weight"], sd[q + "bias"], sd["action_decoder.weight"], sd["action_decoder.bias"])
    y = torch.randn(7, 384, generator=g)
    ref = F.linear(F.linear(y, sd[q + "weight"], sd[q + "bias"]), sd["action_decoder.weight"], sd["action_decoder.bias"])
    assert Wh.shape == (8, 384) and _rel((y @ Wh.t() + bh)[:, :3], ref) < 1e-5 and float((y @ Wh.t() + bh)[:, 3:].abs().max()) == 0.0


def test_missing_tensors_are_reported():
    import pytest
    sd = random_nextdit_state_dict(0)
    del sd["traj_dit.model.layers.0.attn1.to_q.weight"]
    m = object.__new__(N.NextDiTSystem1)        # no CUDA device here: exercise the key check only
    m.device = torch.device("cpu")
    with pytest.raises(KeyError):
        m.load_state_dict(sd)
