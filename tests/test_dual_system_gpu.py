"""End-to-end dual-system step through the model-level mirror class (tiny Qwen config + full-size NavDP head):
System-2 latents -> System-1 trajectories -> action ids, against the oracle chain on the same weights and noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def test_dual_system_step_matches_oracle_chain():
    from internnav_b200.internvla_n1 import InternVLAN1ForCausalLM, InternVLAN1Net
    from internnav_b200.manifest import random_navdp_state_dict
    from oracle import navdp_oracle as O, qwen_oracle as Q, weights
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = Q.tiny_cfg()  # hidden 256 -> NavDP vlm_token_dim 256
    s2_sd = Q.make_s2_state_dict(cfg, seed=3, vocab_rows=512)
    s1_sd = random_navdp_state_dict(seed=4, vlm_token_dim=cfg["hidden"])
    model = InternVLAN1ForCausalLM(cfg, device="cuda:0")
    model.load_parts(s2_sd, s1_sd)
    assert model.get_n_query() == 4 and model.get_system1_type() == "navdp_async"
    with pytest.raises(NotImplementedError):
        model.forward()

    B = 2
    rng = np.random.Generator(np.random.PCG64(8))
    gpp = [[(1, 16, 16)], [(1, 8, 12), (1, 16, 16)]]
    prompts = [Q.make_prompt(rng, 9, gs, 21) for gs in gpp]
    grids = [g for gs in gpp for g in gs]
    n_p = sum(t * h * w for t, h, w in grids)
    g = torch.Generator().manual_seed(2)
    px = torch.randn(n_p, 1176, generator=g).bfloat16().cuda()
    with pytest.raises(RuntimeError):   # this state_dict has no lm_head.weight: greedy decode must refuse, not improvise
        model.generate(prompts, px, grids, max_new_tokens=4)
    inp = weights.make_inputs(9, B=B, K=20)
    rgb, dep = inp["rgb"].cuda(), inp["depth"].cuda()
    x0, nz = inp["x_init"].cuda(), inp["step_noise"].cuda()

    traj, acts = model.dual_system_step(prompts, px, grids, rgb, dep, x_init=x0, step_noise=nz)
    assert traj.shape == (B * 32, 32, 3) and len(acts) == B and all(len(a) <= 4 for a in acts)

    # oracle chain (fp32, on the GPU for speed): per-env generate_latents, then the batched System-1 oracle
    s2_gpu = {k: v.cuda() for k, v in s2_sd.items()}
    s1_gpu = {k: v.cuda().float() for k, v in s1_sd.items()}
    lat, off = [], 0
    with torch.no_grad():
        for ids, gs in zip(prompts, gpp):
            npb = sum(t * h * w for t, h, w in gs)
            lat.append(Q.generate_latents(s2_gpu, cfg, torch.tensor([ids]), px[off:off + npb].float(), gs))
            off += npb
        lat = torch.cat(lat)
        mine_lat = model.generate_latents(prompts, px, grids)
        assert _rel(mine_lat, lat) < 2e-2
        ref = O.predict_pointgoal_action_async(s1_gpu, lat, rgb, dep, x0, nz, K=20)
        # reference-equivalent run: the same oracle chain in bf16 (what the reference's eager bf16 model computes)
        s2_b = {k: v.bfloat16() for k, v in s2_gpu.items()}
        s1_b = {k: v.bfloat16() for k, v in s1_gpu.items()}
        lat_b, off = [], 0
        for ids, gs in zip(prompts, gpp):
            npb = sum(t * h * w for t, h, w in gs)
            lat_b.append(Q.generate_latents(s2_b, cfg, torch.tensor([ids]), px[off:off + npb], gs))
            off += npb
        eager = O.predict_pointgoal_action_async(s1_b, torch.cat(lat_b), rgb.bfloat16(), dep.bfloat16(), x0.bfloat16(),
                                                 nz.bfloat16(), K=20)
    e, e_eager = _rel(traj, ref), _rel(eager, ref)
    print("dual-system trajectories rel err vs oracle chain", e, "bf16 eager chain", e_eager)
    # end of a two-model chain (tiny random-weight Qwen -> goal token -> 20 sampler steps): the reference-equivalent bf16
    # run itself is at 2.7e-2 here (measured on B200; ours 3.3e-2), above the per-stage 2e-2 bar, so the binding bar for the
    # chained output is the relative one (<= 2x bf16 eager); the per-stage bars are asserted where the stages are tested
    assert e < 4e-2 and e < 2 * e_eager + 2e-3, (e, e_eager)
    # policy wrapper: same trajectories -> same ids as the batched tail
    pol = InternVLAN1Net(model)
    outs = pol.s1_step_latent(rgb, dep, mine_lat)
    assert len(outs) == B and all(isinstance(o.idx, list) for o in outs)


def test_training_forward_from_collated_batch():
    """SURVEY §8 row a13, forward: a collated dual-system batch (TRAJ tokens appended, right padding, ragged frame counts)
    through InternVLAN1ForCausalLM.forward vs the oracle chain -- padded-batch decoder with latent_queries at the TRAJ
    positions (qwen_oracle.training_traj_states) -> navdp_oracle.s1_training_loss -- with the noise draws injected."""
    from internnav_b200.internvla_n1 import InternVLAN1ForCausalLM
    from internnav_b200.manifest import random_navdp_state_dict
    from internnav_b200.training import collate_traj_batch
    from oracle import navdp_oracle as O, qwen_oracle as Q
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = Q.tiny_cfg()
    s2_sd = Q.make_s2_state_dict(cfg, seed=13, vocab_rows=512)
    s1_sd = random_navdp_state_dict(seed=14, vlm_token_dim=cfg["hidden"])
    model = InternVLAN1ForCausalLM(cfg, device="cuda:0")
    model.load_parts(s2_sd, s1_sd)
    rng = np.random.Generator(np.random.PCG64(15))
    g = torch.Generator().manual_seed(16)
    gpp = [[(1, 8, 12)], [(1, 16, 16), (1, 4, 8)], [(1, 4, 4)]]
    frames = [3, 2, 1]
    inst = []
    for gs, f in zip(gpp, frames):
        ids = torch.tensor([Q.make_prompt(rng, 6, gs, 15)])
        n_p = sum(t * h * w for t, h, w in gs)
        inst.append(dict(input_ids=ids, labels=torch.full_like(ids, -100), pixel_values=torch.randn(n_p, 1176, generator=g),
                         image_grid_thw=torch.tensor(gs), traj_images=torch.rand(f, 224, 224, 3, generator=g),
                         traj_depths=torch.rand(f, 224, 224, generator=g) * 5, traj_poses=torch.randn(f, 32, 3, generator=g) * 0.5))
    batch = collate_traj_batch(inst)
    B, fmax = len(inst), max(frames)
    assert batch["traj_images"].shape == (B, fmax, 224, 224, 3) and batch["video_frame_num"].tolist() == frames
    assert torch.equal(batch["traj_images"][2, 1], batch["traj_images"][2, 0])          # last frame repeated
    noise = torch.randn(B * fmax, 32, 3, generator=g).cuda()
    ts = torch.randint(0, 20, (B * fmax,), generator=g)
    dev = {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}
    dev["pixel_values"] = dev["pixel_values"].bfloat16()
    out = model.forward(noise=noise, timesteps=ts, **dev)
    s2_gpu = {k: v.cuda() for k, v in s2_sd.items()}
    s1_gpu = {k: v.cuda().float() for k, v in s1_sd.items()}
    with torch.no_grad():
        hs = Q.training_traj_states(s2_gpu, cfg, batch["input_ids"], batch["attention_mask"], dev["pixel_values"].float(),
                                    batch["image_grid_thw"], batch["t_s_pos"])
        ref = O.s1_training_loss(s1_gpu, hs, dev["traj_images"], dev["traj_depths"], dev["traj_poses"],
                                 dev["video_frame_num"], noise, ts.cuda())
    e_h = _rel(out.traj_hidden_states, hs)
    print("training forward: traj states rel err", e_h, "loss", float(out.loss), "oracle", float(ref))
    assert e_h < 2e-2
    assert abs(float(out.loss) - float(ref)) / float(ref) < 2e-2
    with pytest.raises(ValueError):   # t_s_pos must point at the TRAJ tokens
        model.forward(noise=noise, timesteps=ts, **{**dev, "t_s_pos": [p - 1 for p in batch["t_s_pos"]]})
