"""Generates tests/golden/navdp_policy_reference.npz: outputs of the REFERENCE's own stand-alone NavDPNet
(internnav/model/basemodel/navdp/navdp_policy.py, run here on the CPU in fp32 through oracle/ref_loader.py) on seeded
weights (internnav_b200.manifest.random_navdp_policy_state_dict(seed=7)) and seeded inputs.  Needs /root/reference.

    python -m oracle.gen_golden_navdp_policy
"""
import os

import numpy as np
import torch

from internnav_b200.manifest import random_navdp_policy_state_dict
from . import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def make_inputs(seed=11, B=1, Ns=8, T=24, K=10, m=8):
    g = torch.Generator().manual_seed(seed)
    return dict(images=torch.rand(B, m, 224, 224, 3, generator=g), depths=torch.rand(B, 1, 224, 224, 1, generator=g) * 5.0,
                goal=torch.randn(B, 3, generator=g), x_init=torch.randn(B * Ns, T, 3, generator=g),
                step_noise=torch.randn(K - 1, B * Ns, T, 3, generator=g))


def main():
    torch.set_num_threads(os.cpu_count())
    net = ref_loader.build_reference_navdp_policy()
    sd = random_navdp_policy_state_dict(seed=7)
    missing, unexpected = net.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith(("image_encoder.", "pixel_encoder.", "pixel_aux_head.", "image_aux_head.")) for k in missing), missing
    inp = make_inputs()
    out = {}
    with torch.no_grad():
        rgbd = net.rgbd_encoder(inp["images"], inp["depths"])
        out["rgbd"] = rgbd.numpy()
        goal = net.point_encoder(inp["goal"]).unsqueeze(1)
        out["eps"] = net.predict_noise(inp["x_init"], torch.tensor([7]), goal, rgbd).numpy()
        out["critic_of_x_init"] = net.predict_critic(inp["x_init"], rgbd).numpy()
        # full entry points with the sampler's draws injected
        real = torch.randn
        for name in ("pointgoal", "nogoal"):
            net.noise_scheduler.noise_queue = [inp["step_noise"][i] for i in range(inp["step_noise"].shape[0])]
            torch.randn = lambda *a, **k: inp["x_init"].clone()
            try:
                if name == "pointgoal":
                    neg, pos = net.predict_pointgoal_batch_action_vel(inp["goal"], inp["images"], inp["depths"], sample_num=8)
                else:
                    neg, pos = net.predict_nogoal_batch_action_vel(inp["images"], inp["depths"], sample_num=8)
            finally:
                torch.randn = real
                net.noise_scheduler.noise_queue = None
            out[name + "_negative"], out[name + "_positive"] = neg.numpy(), pos.numpy()
    path = os.path.join(ROOT, "tests", "golden", "navdp_policy_reference.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
