"""Generate tests/golden/s1_training_reference.npz from the REFERENCE's own NavDP module -- build container only.

    python -m oracle.gen_golden_training

Row a13 (training branch, System-1 half): `NavDP_Policy_DPT_CriticSum_DAT.forward_vlm_traj` (navdp.py L291-312) of the
reference class itself, on seeded weights / inputs with the draws of `sample_noise` (L165-175) injected, followed by
the masked-MSE of `InternVLAN1ForCausalLM.forward` (internvla_n1.py L287-303, restated here in five lines because that
class cannot be imported without diffusers / NextDiT) and ONE autograd backward through the reference module.

Stored: the prediction, the loss, and for every parameter that receives a gradient its L2 norm and its dot product
with a seeded probe vector (two numbers per tensor instead of 77 M floats), plus the gradient with respect to the
latent tokens (the quantity that continues into the frozen LLM towards `latent_queries`).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, weights  # noqa: E402

CASE = dict(seed=201, B=2, f=2, frames=[2, 1], K=20)


def make_batch(case=CASE):
    rng = np.random.Generator(np.random.PCG64([case["seed"], 777]))
    f32 = np.float32
    B, f = case["B"], case["f"]
    return dict(
        hs=torch.from_numpy(rng.standard_normal((B, 4, 3584), dtype=f32)),
        traj_images=torch.from_numpy(rng.random((B, f, 224, 224, 3), dtype=f32)),
        traj_depths=torch.from_numpy(rng.random((B, f, 224, 224), dtype=f32) * f32(5.0)),
        traj_poses=torch.from_numpy(rng.standard_normal((B, f, 32, 3), dtype=f32) * f32(0.5)),
        video_frame_num=torch.tensor(case["frames"]),
        noise=torch.from_numpy(rng.standard_normal((B * f, 32, 3), dtype=f32)),
        timesteps=torch.from_numpy(rng.integers(0, case["K"], B * f)).long())


def probe(name, shape):
    import zlib
    rng = np.random.Generator(np.random.PCG64([zlib.crc32(name.encode()), 4242]))
    return torch.from_numpy(rng.standard_normal(shape, dtype=np.float32))


def dp_inputs(batch):
    """internvla_n1.py L289-296: [goal frame, current frame] pairs for every selected frame."""
    ti, td = batch["traj_images"], batch["traj_depths"]
    B, f = ti.shape[:2]
    cur_i, cur_d = ti.flatten(0, 1), td.flatten(0, 1)
    g_i = ti[:, 0:1].repeat(1, f, 1, 1, 1).flatten(0, 1)
    g_d = td[:, 0:1].repeat(1, f, 1, 1).flatten(0, 1)
    return torch.stack([g_i, cur_i], dim=1), torch.stack([g_d, cur_d], dim=1).unsqueeze(-1)


def main():
    torch.set_num_threads(os.cpu_count())
    m = ref_loader.build_reference_navdp(predict_size=32, memory_size=2, navdp_version=0.1)
    m.load_state_dict(weights.make_state_dict(0), strict=True)
    m.input_dtype = torch.float32
    for p in m.parameters():
        p.requires_grad_(True)
    batch = make_batch()
    B, f = batch["traj_images"].shape[:2]
    hs = batch["hs"].clone().requires_grad_(True)
    hs_rep = hs.unsqueeze(1).repeat(1, f, 1, 1).flatten(0, 1)                       # internvla_n1.py L236
    images_dp, depths_dp = dp_inputs(batch)
    real_randn, real_randint = torch.randn, torch.randint
    torch.randn = lambda *a, **k: batch["noise"].clone()
    torch.randint = lambda *a, **k: batch["timesteps"].clone()
    try:
        pred, eps = m.forward_vlm_traj(hs_rep, images_dp, depths_dp, tensor_label_actions=batch["traj_poses"])
    finally:
        torch.randn, torch.randint = real_randn, real_randint
    assert torch.equal(eps, batch["noise"])
    err = (pred - eps).square()                                                      # internvla_n1.py L299-303
    mask = (torch.arange(f).expand(B, f) < batch["video_frame_num"].unsqueeze(1)).flatten(0, 1)[:, None, None]
    loss = (err * mask).sum() / mask.sum() / (err.shape[1] * err.shape[2])
    loss.backward()
    gold = {"pred": pred.detach().numpy(), "loss": np.float32(loss.item()), "grad_hs": hs.grad.numpy()}
    names, norms, dots = [], [], []
    for name, p in m.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        norms.append(float(p.grad.norm()))
        dots.append(float((p.grad * probe(name, tuple(p.shape))).sum()))
    gold["grad_names"] = np.array(names)
    gold["grad_norms"] = np.array(norms, dtype=np.float64)
    gold["grad_dots"] = np.array(dots, dtype=np.float64)
    out = os.path.join(ROOT, "tests", "golden", "s1_training_reference.npz")
    np.savez_compressed(out, **gold)
    none = [n for n, p in m.named_parameters() if p.grad is None]
    print("loss", float(loss), "params with grad", len(names), "without", len(none), "bytes", os.path.getsize(out))
    print("no grad:", none[:12])


if __name__ == "__main__":
    main()
