"""The batched eager-GPU BASELINE arm (oracle/eager_gpu.py, used only by bench.py) computes what the parity oracle
computes: checked here on the CPU in fp32 with a tiny config (SDPA's math path), so the arm bench.py times is the reference
algorithm and not something cheaper."""
import numpy as np
import torch


def test_batched_eager_equals_oracle():
    from oracle import eager_gpu as E, navdp_oracle as O, qwen_oracle as Q, weights
    cfg = Q.tiny_cfg()
    sd = Q.make_s2_state_dict(cfg, seed=3, vocab_rows=512)
    rng = np.random.Generator(np.random.PCG64(4))
    grid = (1, 12, 20)
    B = 3
    prompts = [Q.make_prompt(rng, 5, [grid], 9) for _ in range(B)]
    n_p = grid[0] * grid[1] * grid[2]
    px = torch.randn(B * n_p, 1176, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        got = E.generate_latents_batched(sd, cfg, torch.tensor(prompts), px, grid)
        ref = torch.cat([Q.generate_latents(sd, cfg, torch.tensor([p]), px[b * n_p:(b + 1) * n_p], [grid])
                         for b, p in enumerate(prompts)])
        v_got = E.vit_forward_batched(sd, cfg, px, grid, B)
        v_ref = Q.vit_forward(sd, cfg, px, [grid] * B)
    assert torch.allclose(v_got, v_ref, atol=2e-4, rtol=2e-4), float((v_got - v_ref).abs().max())
    assert torch.allclose(got, ref, atol=5e-4, rtol=5e-4), float((got - ref).abs().max())
    # System 1 with the attention switch: identical mathematics
    sd1 = weights.make_state_dict(0)
    inp = weights.make_inputs(3, B=1, K=3)
    k = torch.tensor([2])
    with torch.no_grad():
        a = O.predict_noise(sd1, inp["x_init"], k, inp["goal"], inp["rgbd"])
        r = O.rgbd_encoder(sd1, inp["rgb"], inp["depth"])
        O.ATTENTION = "sdpa"
        try:
            b = O.predict_noise(sd1, inp["x_init"], k, inp["goal"], inp["rgbd"])
            r2 = O.rgbd_encoder(sd1, inp["rgb"], inp["depth"])
        finally:
            O.ATTENTION = "math"
    assert torch.allclose(a, b, atol=2e-4, rtol=2e-4) and torch.allclose(r, r2, atol=2e-4, rtol=2e-4)
