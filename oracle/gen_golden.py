"""Generate tests/golden/*.npz from the REFERENCE's own modules -- run in the build container only.

    python -m oracle.gen_golden

Imports /root/reference's NavDP / DINOv2 / vln_utils code through oracle/ref_loader.py, loads the deterministic
synthetic weights of oracle/weights.py (seed 0) into the reference class with strict=True, runs the reference forward
on seeded inputs with all randomness injected, and stores ONLY the outputs (inputs and weights are regenerated from
their seeds at test time).  Also refreshes oracle/navdp_manifest.json from the reference state_dict.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_loader, weights  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    m = ref_loader.build_reference_navdp(predict_size=32, memory_size=2, navdp_version=0.1)
    man = {k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in m.state_dict().items()}
    with open(os.path.join(ROOT, "oracle", "navdp_manifest.json"), "w") as fh:
        json.dump(man, fh, indent=0)
    sd = weights.make_state_dict(0)
    m.load_state_dict(sd, strict=True)
    vu = ref_loader.load_reference_vln_utils()
    gold = {}
    with torch.no_grad():
        # (1) RGB-D encoder, 2 environments                         navdp_backbone.py L151-202
        inp = weights.make_inputs(101, B=2)
        gold["rgbd_B2"] = m.rgbd_encoder(inp["rgb"], inp["depth"]).numpy()
        # (2) goal token, 3 environments (the reference loops bs=1; run per env)   navdp.py L237-238
        inp = weights.make_inputs(102, B=3)
        gold["goal_B3"] = torch.cat([m.goal_compressor(m.vlm_embed_mlp(inp["latents"][i:i + 1]), None)
                                     for i in range(3)]).numpy()
        # (3) predict_noise at the reference shape and the BASELINE configs[1] shape   navdp.py L177-195
        inp = weights.make_inputs(103, B=1, T=32, Ns=32)
        gold["eps_T32"] = m.predict_noise(inp["x_init"], torch.tensor([7]), inp["goal"], inp["rgbd"]).numpy()
        # horizon 8 needs a model built with predict_size=8 (tgt_mask / out_pos_embed are sized by it, navdp.py L70, L81);
        # it gets the same weights with out_pos_embed[:, :8]
        m8 = ref_loader.build_reference_navdp(predict_size=8, memory_size=2, navdp_version=0.1)
        sd8 = dict(sd)
        sd8["out_pos_embed"] = sd["out_pos_embed"][:, :8].clone()
        m8.load_state_dict(sd8, strict=True)
        inp = weights.make_inputs(103, B=1, T=8, Ns=32)
        gold["eps_T8"] = m8.predict_noise(inp["x_init"], torch.tensor([13]), inp["goal"], inp["rgbd"]).numpy()
        del m8
        # (4) full predict_pointgoal_action_async with injected noise   navdp.py L197-253
        inp = weights.make_inputs(104, B=1, K=20)
        m.noise_scheduler.noise_queue = [inp["step_noise"][i] for i in range(19)]
        real = torch.randn
        torch.randn = lambda *a, **k: inp["x_init"].clone()
        try:
            traj = m.predict_pointgoal_action_async(inp["latents"], inp["rgb"], inp["depth"])
        finally:
            torch.randn = real
            m.noise_scheduler.noise_queue = None
        gold["traj_full"] = traj.numpy()
        # (5) integer tail on seeded trajectories   vln_utils.py L63-136
        acts = []
        rng = np.random.Generator(np.random.PCG64(105))
        for i in range(12):
            tr = torch.from_numpy((rng.standard_normal((32, 32, 3), dtype=np.float32) * 0.15 +
                                   np.array([0.5 * np.cos(i), 0.5 * np.sin(i), 0.0], dtype=np.float32)))
            acts.append(vu.traj_to_actions(tr.clone()))
        acts.append(vu.traj_to_actions(traj.clone()))
    np.savez_compressed(os.path.join(out_dir, "s1_reference_outputs.npz"), **gold)
    with open(os.path.join(out_dir, "traj_to_actions.json"), "w") as fh:
        json.dump(acts, fh)
    print("wrote", {k: v.shape for k, v in gold.items()}, "and", len(acts), "action lists")


if __name__ == "__main__":
    main()
