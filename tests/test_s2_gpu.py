"""System-2 parity on the GPU: CUDA path (through the C ABI) vs the fp32 oracle (oracle/qwen_oracle.py) on seeded
weights.  Tolerance: rel-L2 vs the fp32 oracle <= 2e-2 and <= 2x the error of the reference-equivalent
bf16 PyTorch run; position ids bit-exact."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = 2e-2


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def _setup(cfg, seed, vocab_rows=None):
    from internnav_b200.qwen import System2
    from oracle import qwen_oracle as Q
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = Q.make_s2_state_dict(cfg, seed=seed, vocab_rows=vocab_rows)
    s2 = System2(cfg, device="cuda:0")
    s2.load_state_dict(sd)
    sd_gpu = {k: v.cuda() for k, v in sd.items()}
    return s2, sd_gpu


def _check_vit(s2, sd, cfg, grids, seed):
    from oracle import qwen_oracle as Q
    n_p = sum(t * h * w for t, h, w in grids)
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(n_p, 1176, generator=g).bfloat16().cuda()
    out = s2.visual(px, grids)
    with torch.no_grad():
        ref = Q.vit_forward(sd, cfg, px.float(), grids)
        eager = Q.vit_forward({k: v.bfloat16() for k, v in sd.items() if k.startswith("visual.")}, cfg, px, grids)
    e, ee = _rel(out, ref), _rel(eager, ref)
    print("vit", grids, "rel err", e, "bf16 eager", ee)
    assert e < TOL and e < 2 * ee + 2e-3, (e, ee)
    return out


def _check_latents(s2, sd, cfg, prompts, grids_per_prompt, seed):
    from oracle import qwen_oracle as Q
    all_grids = [g for gs in grids_per_prompt for g in gs]
    n_p = sum(t * h * w for t, h, w in all_grids)
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(n_p, 1176, generator=g).bfloat16().cuda()
    out = s2.generate_latents(prompts, px, all_grids)
    # position ids of the plan: bit-exact vs the oracle's rope_index
    plan = s2.llm_plan(prompts, all_grids)
    pos, delta = s2.positions(plan, len(prompts))
    off_t, off_p = 0, 0
    sdb = {k: v.bfloat16() for k, v in sd.items()}
    for b, (ids, gs) in enumerate(zip(prompts, grids_per_prompt)):
        npb = sum(t * h * w for t, h, w in gs)
        ids_t = torch.tensor([ids])
        full = torch.cat([ids_t, torch.full((1, cfg["n_query"]), Q.TRAJ_TOKEN_INDEX)], dim=1)
        rp, rd = Q.rope_index(full, torch.tensor(gs).reshape(-1, 3), cfg["v_merge"])
        L = full.shape[1]
        assert torch.equal(pos[:, off_t:off_t + L], rp[:, 0]) and int(delta[b]) == int(rd)
        with torch.no_grad():
            pxb = px[off_p:off_p + npb]
            ref = Q.generate_latents(sd, cfg, ids_t, pxb.float(), gs)
            eager = Q.generate_latents(sdb, cfg, ids_t, pxb, gs)
        e, ee = _rel(out[b], ref[0]), _rel(eager[0], ref[0])
        print("latents env", b, "S", L, "rel err", e, "bf16 eager", ee)
        assert e < TOL and e < 2 * ee + 2e-3, (e, ee)
        off_t += L
        off_p += npb


def test_tiny_config():
    from oracle import qwen_oracle as Q
    cfg = Q.tiny_cfg()
    s2, sd = _setup(cfg, 1, vocab_rows=512)
    _check_vit(s2, sd, cfg, [(1, 16, 20), (1, 28, 28)], 0)
    rng = np.random.Generator(np.random.PCG64(5))
    gpp = [[(1, 16, 16)], [(1, 8, 12), (1, 28, 28)], [(1, 4, 4)]]
    prompts = [Q.make_prompt(rng, 7 + 3 * i, gs, 20 - 5 * i) for i, gs in enumerate(gpp)]
    _check_latents(s2, sd, cfg, prompts, gpp, 1)


def test_real_width_shallow():
    """Qwen2.5-VL-7B widths (1280/16 heads/3420; 3584/28q/4kv/18944) with 2 + 2 layers: every GEMM shape, head layout and
    padding rule of the full model, at a depth the fp32 oracle finishes in seconds."""
    from oracle import qwen_oracle as Q
    cfg = dict(Q.QWEN25VL_7B)
    cfg.update(v_depth=2, fullatt=[1], layers=2)
    s2, sd = _setup(cfg, 2, vocab_rows=1024)
    _check_vit(s2, sd, cfg, [(1, 28, 28)], 3)
    rng = np.random.Generator(np.random.PCG64(9))
    gpp = [[(1, 28, 28)], [(1, 28, 28)]]
    prompts = [Q.make_prompt(rng, 24, gs, 80 - 10 * i) for i, gs in enumerate(gpp)]  # S = 304 / 294
    _check_latents(s2, sd, cfg, prompts, gpp, 4)


# ------------------------------------------------------------------------------------------------ greedy decode
MARGIN = 0.15  # logit units: bf16 logits (|x| ~ 4-5, ulp 2^-6) on a bf16 forward vs the fp32 oracle


def _check_generate(s2, sd, cfg, prompts, gpp, seed, max_new, eos):
    """Teacher-forced parity of the greedy decode: with OUR prefix, the token we picked must be the oracle's argmax or
    within MARGIN of it (bf16 near-ties); the stop rule and the KV-reuse latents are checked exactly / to TOL."""
    from oracle import qwen_oracle as Q
    all_grids = [g for gs in gpp for g in gs]
    g = torch.Generator().manual_seed(seed)
    px = torch.randn(sum(t * h * w for t, h, w in all_grids), 1176, generator=g).bfloat16().cuda()
    out, lat, passes = s2.generate(prompts, px, all_grids, max_new_tokens=max_new, eos_token_ids=eos, with_latents=True)
    assert passes == max(len(o) for o in out) - 1  # the pass after the last sampled token is never run
    off_p, exact, total = 0, 0, 0
    full_prompts = []
    for b, (ids, gs) in enumerate(zip(prompts, gpp)):
        npb = sum(t * h * w for t, h, w in gs)
        pxb = px[off_p:off_p + npb]
        off_p += npb
        assert 1 <= len(out[b]) <= max_new
        assert all(t not in eos for t in out[b][:-1]), "generation continued past an eos id"
        assert out[b][-1] in eos or len(out[b]) == max_new, "generation stopped early"
        cur = torch.tensor([ids])
        with torch.no_grad():
            feats = Q.vit_forward(sd, cfg, pxb.float(), gs)
            for tok in out[b]:
                lg = Q.next_token_logits(sd, cfg, cur, feats, gs)
                gap = float(lg.max() - lg[tok])
                assert gap <= MARGIN, ("env %d picked %d, oracle argmax %d, gap %.3f" % (b, tok, int(lg.argmax()), gap))
                exact += int(int(lg.argmax()) == tok)
                total += 1
                cur = torch.cat([cur, torch.tensor([[tok]])], dim=1)
            ref = Q.generate_latents(sd, cfg, cur, pxb.float(), gs)
        e = _rel(lat[b], ref[0])
        print("generate env", b, "tokens", out[b], "latent rel err (KV reuse)", e)
        assert e < TOL, e
        full_prompts.append(list(ids) + list(out[b]))
    assert exact >= 0.6 * total, (exact, total)
    # the cache-extension latents equal our own re-prefill of prompt + generated ids (the reference's route)
    lat2 = s2.generate_latents(full_prompts, px, all_grids)
    e2 = _rel(lat, lat2)
    print("KV-reuse vs re-prefill latents rel", e2, "exact", exact, "/", total, "passes", passes)
    assert e2 < 1e-2, e2
    return out


def _setup_gen(cfg, seed):
    from internnav_b200.qwen import System2
    from oracle import qwen_oracle as Q
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = Q.make_s2_state_dict(cfg, seed=seed, lm_head=True)
    s2 = System2(cfg, device="cuda:0")
    s2.load_state_dict(sd)
    return s2, {k: v.cuda() for k, v in sd.items()}


def test_generate_tiny():
    from oracle import qwen_oracle as Q
    cfg = Q.tiny_cfg()
    s2, sd = _setup_gen(cfg, 5)
    rng = np.random.Generator(np.random.PCG64(3))
    gpp = [[(1, 8, 12)], [(1, 16, 16), (1, 4, 4)], [(1, 4, 8)]]
    prompts = [Q.make_prompt(rng, 6 + 2 * i, gs, 10 + 3 * i) for i, gs in enumerate(gpp)]
    out = _check_generate(s2, sd, cfg, prompts, gpp, 0, 8, ())          # budget-limited, ragged prompts
    # golden: env 0 is the committed single-prompt case of tests/golden/greedy_generate.json (same seeds)
    import json
    import os
    with open(os.path.join(os.path.dirname(__file__), "golden", "greedy_generate.json")) as fh:
        gold = json.load(fh)["tokens"]
    agree = sum(int(a == b) for a, b in zip(out[0], gold))
    print("golden agreement", agree, "/", len(out[0]))
    stop = out[0][2]
    out2 = _check_generate(s2, sd, cfg, prompts, gpp, 0, 8, (stop,))     # env 0 stops after 3 tokens, the others go on
    assert out2[0] == out[0][:3]
    assert any(len(o) > 3 for o in out2[1:]) or all(stop in o for o in out2[1:])


def test_generate_real_width_shallow():
    from oracle import qwen_oracle as Q
    cfg = dict(Q.QWEN25VL_7B)
    cfg.update(v_depth=2, fullatt=[1], layers=2)
    s2, sd = _setup_gen(cfg, 6)
    rng = np.random.Generator(np.random.PCG64(11))
    gpp = [[(1, 28, 28)], [(1, 28, 28)]]
    prompts = [Q.make_prompt(rng, 24, gs, 80 - 10 * i) for i, gs in enumerate(gpp)]
    _check_generate(s2, sd, cfg, prompts, gpp, 7, 5, (151645, 151643))


def test_generate_requires_lm_head():
    from oracle import qwen_oracle as Q
    cfg = Q.tiny_cfg()
    s2, _ = _setup(cfg, 1, vocab_rows=512)
    with pytest.raises(RuntimeError):
        s2.generate([[1, 2, 3]], torch.zeros(0, 1176), [], max_new_tokens=2)


def test_from_pretrained_checkpoint_directory(tmp_path):
    """`InternVLAN1ForCausalLM.from_pretrained(dir, torch_dtype=bf16, attn_implementation=..., device_map={"": dev})`
    (internvla_n1_policy.py L33-38) on a checkpoint directory in the released layout (flat 4.51 config.json, sharded
    safetensors, `model.navdp.*` keys): same latents and trajectories as the model loaded from the in-memory parts."""
    import json
    from safetensors.torch import save_file
    from internnav_b200.internvla_n1 import InternVLAN1ForCausalLM
    from internnav_b200.manifest import random_navdp_state_dict
    from oracle import qwen_oracle as Q, weights
    cfg = Q.tiny_cfg()
    s2_sd = Q.make_s2_state_dict(cfg, seed=21, vocab_rows=512)
    s1_sd = random_navdp_state_dict(seed=22, vlm_token_dim=cfg["hidden"])
    full = {k: v.contiguous() for k, v in s2_sd.items()}
    full.update({"model.navdp." + k: v.detach().cpu().contiguous() for k, v in s1_sd.items()})
    names = sorted(full)
    shard_a = {k: full[k] for k in names[::2]}
    shard_b = {k: full[k] for k in names[1::2]}
    save_file(shard_a, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file(shard_b, str(tmp_path / "model-00002-of-00002.safetensors"))
    wm = {k: "model-00001-of-00002.safetensors" for k in shard_a}
    wm.update({k: "model-00002-of-00002.safetensors" for k in shard_b})
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": wm}))
    conf = dict(model_type="internvla_n1", hidden_size=cfg["hidden"], num_hidden_layers=cfg["layers"],
                num_attention_heads=cfg["heads"], num_key_value_heads=cfg["kv_heads"], intermediate_size=cfg["inter"],
                vocab_size=cfg["vocab"], rms_norm_eps=cfg["rms_eps"], rope_theta=cfg["rope_theta"],
                rope_scaling={"type": "mrope", "mrope_section": cfg["mrope"]}, n_query=4, system1="navdp_async",
                vision_config=dict(depth=cfg["v_depth"], hidden_size=cfg["v_hidden"], num_heads=cfg["v_heads"],
                                   intermediate_size=cfg["v_inter"], out_hidden_size=cfg["v_out"], patch_size=14,
                                   temporal_patch_size=2, spatial_merge_size=2, window_size=112,
                                   fullatt_block_indexes=cfg["fullatt"]))
    (tmp_path / "config.json").write_text(json.dumps(conf))
    a = InternVLAN1ForCausalLM.from_pretrained(str(tmp_path), torch_dtype=torch.bfloat16,
                                               attn_implementation="flash_attention_2", device_map={"": "cuda:0"})
    b = InternVLAN1ForCausalLM(cfg, device="cuda:0")
    b.load_parts(s2_sd, s1_sd)
    assert a.cfg == b.cfg and a.name_or_path == str(tmp_path)
    rng = np.random.Generator(np.random.PCG64(23))
    gs = [(1, 8, 12)]
    prompts = [Q.make_prompt(rng, 7, gs, 12)]
    px = torch.randn(96, 1176, generator=torch.Generator().manual_seed(1)).bfloat16().cuda()
    la, lb = a.generate_latents(prompts, px, gs), b.generate_latents(prompts, px, gs)
    assert torch.equal(la, lb)
    inp = weights.make_inputs(24, B=1, K=20)
    ta = a.generate_traj(la, inp["rgb"].cuda(), inp["depth"].cuda(), x_init=inp["x_init"].cuda(), step_noise=inp["step_noise"].cuda())
    tb = b.generate_traj(lb, inp["rgb"].cuda(), inp["depth"].cuda(), x_init=inp["x_init"].cuda(), step_noise=inp["step_noise"].cuda())
    assert torch.equal(ta, tb)
