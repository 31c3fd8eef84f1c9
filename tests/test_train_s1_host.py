"""The System-1 training SCHEDULE (internnav_b200/train_s1.py) on the CPU: the same sequence of kernel calls the GPU
backend will issue, executed by the fp32 reference backend of tests/ops_reference.py, against the oracle's gradients
(which are pinned to the reference module's autograd).  Validates operand layouts, saves and accumulation order; the
kernels themselves are covered by tests/test_bwd_ops_gpu.py."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ops_reference import TorchOps  # noqa: E402

from internnav_b200.train_s1 import S1TrainStep  # noqa: E402
from oracle import ddpm, gen_golden_training as G, navdp_oracle as O, weights  # noqa: E402


def test_schedule_reproduces_oracle_gradients():
    torch.set_num_threads(os.cpu_count())
    sd = weights.make_state_dict(0)
    b = G.make_batch(dict(G.CASE, seed=203))
    args = (b["hs"], b["traj_images"], b["traj_depths"], b["traj_poses"], b["video_frame_num"], b["noise"], b["timesteps"])
    loss_ref, grads_ref, dhs_ref = O.s1_training_grads(sd, *args)
    # the frozen RGB branch is an input of the schedule (on the GPU: the RGB ViT of n1_rgbd_encode)
    imgs, _ = G.dp_inputs(b)
    with torch.no_grad():
        mean = torch.tensor([0.485, 0.456, 0.406], dtype=torch.bfloat16).float().reshape(1, 3, 1, 1)
        std = torch.tensor([0.229, 0.224, 0.225], dtype=torch.bfloat16).float().reshape(1, 3, 1, 1)
        ti = imgs.permute(0, 1, 4, 2, 3).reshape(-1, 3, 224, 224)
        rgb_tokens = O.dinov2_vits(sd, "rgbd_encoder.rgb_model.", (ti - mean) / std).reshape(imgs.shape[0], 2 * 256, -1)
        ops = TorchOps()
        step = S1TrainStep({k: v.float() for k, v in sd.items() if v.is_floating_point()}, ops)
        acp = torch.as_tensor(ddpm.DDPMScheduler(num_train_timesteps=20).alphas_cumprod).float()
        loss, grads, dhs = step.forward_backward(b["hs"], rgb_tokens, b["traj_depths"], b["traj_poses"], b["video_frame_num"],
                                                 b["noise"], b["timesteps"], acp)
    assert abs(float(loss) - float(loss_ref)) < 1e-5 * max(1.0, abs(float(loss_ref)))
    assert sorted(grads) == sorted(grads_ref), set(grads) ^ set(grads_ref)
    worst = ("", 0.0)
    for k, gr in grads_ref.items():
        rel = float((grads[k].reshape(gr.shape) - gr).norm() / (gr.norm() + 1e-12))
        worst = max(worst, (k, rel), key=lambda t: t[1])
        assert rel < 2e-3 or float((grads[k].reshape(gr.shape) - gr).abs().max()) < 1e-8, (k, rel)
    rel_h = float((dhs - dhs_ref).norm() / dhs_ref.norm())
    print("schedule vs oracle: worst parameter gradient", worst, "latent gradient", rel_h, "kernel calls", ops.calls)
    assert rel_h < 2e-3
    # every heavy operation went through the backend
    assert ops.calls["mm_nt"] > 300 and ops.calls["wgrad"] > 150 and ops.calls["attention_bwd"] == 12 + 4 + 32 + 1 and ops.calls["norm_bwd"] > 70
