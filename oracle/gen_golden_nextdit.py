"""Generate tests/golden/nextdit_reference.npz by running the REFERENCE's own `generate_traj` (nextdit_async branch,
internvla_n1.py L349-432) -- TEST INFRASTRUCTURE, runs only in the container that has /root/reference.

What executes is the reference's code: InternVLAN1ForCausalLM.generate_traj (called unbound on a stand-in `self` that
carries the reference's own sub-modules: MemoryEncoder, QFormer, SinusoidalPositionalEncoding from internvla_n1_arch.py,
NextDiTCrossAttn from nextdit_crossattn_traj.py, DinoVisionTransformer (vits) from the reference's depth_anything tree).
Three things are stand-ins because `diffusers` is absent from this image: the Lumina leaf modules, the flow-matching Euler
scheduler (oracle/diffusers_standin.py) and `randn_tensor` (returns the seeded x_init so the run is reproducible).
Weights: internnav_b200.manifest.random_nextdit_state_dict(seed) loaded with strict=True into those modules.

Usage:  python -m oracle.gen_golden_nextdit"""
import importlib
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SEED, B, NS = 5, 1, 3


def build_reference_model(sd):
    """-> (reference generate_traj function, stand-in self) with the weights of `sd`."""
    from . import diffusers_standin, ref_loader
    ref_loader.load_reference_nextdit()
    tu = types.ModuleType("diffusers.utils.torch_utils")
    tu.randn_tensor = lambda *a, **k: None      # replaced per call below
    sys.modules["diffusers.utils.torch_utils"] = tu
    sys.modules["diffusers.utils"].torch_utils = tu
    main = importlib.import_module("internnav.model.basemodel.internvla_n1.internvla_n1")
    arch = importlib.import_module("internnav.model.basemodel.internvla_n1.internvla_n1_arch")
    cross = importlib.import_module("internnav.model.basemodel.internvla_n1.nextdit_crossattn_traj")
    dpt = importlib.import_module("internnav.model.encoder.depth_anything.depth_anything_v2.dpt")

    class Model(nn.Module):      # the attributes InternVLAN1MetaModel.__init__ creates for 'nextdit_async' (L131-145)
        def __init__(self):
            super().__init__()
            self.traj_dit = cross.NextDiTCrossAttn(cross.NextDiTCrossAttnConfig(latent_embedding_size=arch.LatentEmbSize,
                                                                                _gradient_checkpointing=False))
            self.action_encoder = nn.Linear(3, 384, bias=True)
            self.pos_encoding = arch.SinusoidalPositionalEncoding(384)
            self.action_decoder = nn.Linear(384, 3, bias=True)
            self.cond_projector = nn.Sequential(nn.Linear(3584, arch.LatentEmbSize), nn.GELU(approximate="tanh"),
                                                nn.Linear(arch.LatentEmbSize, arch.LatentEmbSize))
            cfg = {'encoder': 'vits', 'features': 64, 'out_channels': [48, 96, 192, 384]}
            self.rgb_model = dpt.DepthAnythingV2(**cfg).pretrained
            self.memory_encoder = arch.MemoryEncoder()
            self.rgb_resampler = arch.QFormer()

    model = Model().eval()
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and all("freqs_cis" in k for k in missing), (missing, unexpected)

    class Self:
        _resnet_mean = torch.FloatTensor(main._RESNET_MEAN).view(1, 1, 3, 1, 1)
        _resnet_std = torch.FloatTensor(main._RESNET_STD).view(1, 1, 3, 1, 1)

        def get_system1_type(self):
            return "nextdit_async"

        def get_model(self):
            return model

    return main, main.InternVLAN1ForCausalLM.generate_traj, Self(), model


def make_inputs(seed=SEED, batch=B, ns=NS):
    g = torch.Generator().manual_seed(seed)
    return dict(traj_latents=torch.randn(batch, 4, 3584, generator=g),
                images_dp=torch.rand(batch, 2, 224, 224, 3, generator=g),
                x_init=torch.randn(batch * ns, 32, 3, generator=g))


def main():
    sys.path.insert(0, ROOT)
    from internnav_b200.manifest import random_nextdit_state_dict
    sd = random_nextdit_state_dict(SEED)
    mod, generate_traj, self_, model = build_reference_model(sd)
    inp = make_inputs()
    out = {}
    for scale in (1.0, 2.5):
        mod.randn_tensor = lambda shape, generator=None, device=None, dtype=None: inp["x_init"].to(dtype).clone()
        with torch.no_grad():
            out[scale] = generate_traj(self_, inp["traj_latents"], inp["images_dp"], guidance_scale=scale,
                                       num_sample_trajs=NS)
    # intermediate: the condition tokens, recomputed with the same reference modules (internvla_n1.py L363-382)
    with torch.no_grad():
        lat = model.cond_projector(inp["traj_latents"])
        img = (inp["images_dp"].permute(0, 1, 4, 2, 3) - self_._resnet_mean) / self_._resnet_std
        feat = model.rgb_model.get_intermediate_layers(img.flatten(0, 1))[0].unflatten(dim=0, sizes=(1, -1))
        mem = model.memory_encoder(feat.flatten(1, 2))
        tokens = model.rgb_resampler(torch.cat([feat.flatten(1, 2), mem], dim=-1))
        cond = torch.cat([tokens, lat], dim=1)
    path = os.path.join(ROOT, "tests", "golden", "nextdit_reference.npz")
    np.savez_compressed(path, seed=SEED, batch=B, ns=NS, traj_scale_1=out[1.0].numpy(), traj_scale_2p5=out[2.5].numpy(),
                        condition_tokens=cond.numpy())
    print("wrote", path, {k: tuple(v.shape) for k, v in out.items()}, cond.shape)


if __name__ == "__main__":
    main()
