"""One optimisation step of the dual-system training, navdp_async branch:

    collated batch -> System-2 TRAJ states (frozen decoder, K/V cache)        s2_train.cu
                   -> System-1 forward + backward (trainable)                train_s1.py over bwd_kernels.cu
                   -> d loss / d TRAJ states -> d latent_queries             s2_train.cu
                   -> bucketed all-reduce over the data-parallel group       ddp.py (NCCL)
                   -> fused AdamW on fp32 masters, bf16 working copies       bwd_kernels.cu

replacing `Trainer.training_step` -> `InternVLAN1ForCausalLM.forward(...).loss.backward()` -> DDP -> optimizer of the
reference (internnav/trainer/internvla_n1_trainer.py L206-217; scripts/train/base_train/train.py).  The trainable set is
the reference's: every System-1 tensor except the detached RGB ViT, plus `latent_queries`.

STATUS: assembled at the end of round 1 from pieces that are individually specified and CPU-checked (the schedule, the
bucket layout, the backward algorithms) but whose kernels have NOT yet run on a B200; tests/test_bwd_ops_gpu.py holds
the GPU parity tests, skipped until a run is on record.  There is no CPU path.
"""
from collections import OrderedDict

import torch

from . import _bwd, _lib
from .ddp import GradientBuckets
from .train_s1 import GpuOps, S1TrainStep


class DualSystemTrainer:
    def __init__(self, model, navdp_state_dict, latent_queries, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0,
                 bucket_cap_mb=100, process_group=None):
        """model: internnav_b200.internvla_n1.InternVLAN1ForCausalLM with weights loaded (the frozen parts are used from
        it); navdp_state_dict: {reference name: tensor} for `model.navdp.*`; latent_queries [1, n_query, H]."""
        self.model = model
        dev = model.device
        if dev.type != "cuda":
            raise RuntimeError("n1b200 has no CPU path")
        self.device = dev
        self.masters = OrderedDict()
        for k, v in navdp_state_dict.items():
            if v.is_floating_point():
                self.masters[k] = v.detach().to(dev, torch.float32).clone()
        self.latent = latent_queries.detach().to(dev, torch.float32).clone()
        self.s1 = S1TrainStep(self.masters, GpuOps(str(dev)))
        trainable = OrderedDict((k, (tuple(v.shape), torch.float32)) for k, v in self.masters.items()
                                if not k.startswith("rgbd_encoder.rgb_model."))
        trainable["model.latent_queries"] = (tuple(self.latent.shape), torch.float32)
        self.buckets = GradientBuckets(trainable, dev, bucket_cap_mb=bucket_cap_mb, process_group=process_group)
        self.opt = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.m = {k: torch.zeros_like(g) for k, g in self.buckets.grads.items()}
        self.v = {k: torch.zeros_like(g) for k, g in self.buckets.grads.items()}
        self.steps = 0
        import ctypes
        K = 20
        buf = (ctypes.c_float * (K * 5))()
        _lib.check(_lib.lib().n1_ddpm_tables(K, buf))
        t = torch.tensor(list(buf), dtype=torch.float32).view(K, 5)
        # column 0 = sqrt(1 - acp), so acp = 1 - col0^2   (n1_ddpm_tables layout, include/n1b200.h)
        self.alphas_cumprod = (1.0 - t[:, 0] ** 2).to(dev)

    def loss_and_grads(self, batch, noise, timesteps):
        """batch: the collated dict of internnav_b200.training.collate_traj_batch (tensors may live on the host)."""
        m = self.model
        nq = m.config.n_query
        prompts = []
        rows = m._prompts(batch["input_ids"])
        mask = batch["attention_mask"].tolist()
        for b, row in enumerate(rows):
            t = int(batch["t_s_pos"][b])
            prompts.append([tok for i, tok in enumerate(row[:t]) if mask[b][i]])
        grid = batch["image_grid_thw"].tolist()
        hs = m._s2.train_forward(prompts, batch["pixel_values"], grid)
        ti, td = batch["traj_images"].to(self.device), batch["traj_depths"].to(self.device)
        B, f = ti.shape[:2]
        goal_i = ti[:, 0:1].repeat(1, f, 1, 1, 1).flatten(0, 1)
        images_dp = torch.stack([goal_i, ti.flatten(0, 1)], dim=1)
        rgb_tokens = m.model.navdp.rgb_memory_tokens(images_dp)
        loss, grads, dhs = self.s1.forward_backward(hs, rgb_tokens, td, batch["traj_poses"], batch["video_frame_num"], noise,
                                                    timesteps, self.alphas_cumprod, rgb_has_pe=True)
        grads["model.latent_queries"] = m._s2.train_backward(dhs).reshape(self.latent.shape)
        return loss, grads, hs

    def step(self, batch, noise, timesteps):
        loss, grads, _ = self.loss_and_grads(batch, noise, timesteps)
        self.buckets.zero()
        for k, g in grads.items():
            if k in self.buckets.grads:
                self.buckets.grads[k].copy_(g.reshape(self.buckets.grads[k].shape))
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.buckets.all_reduce(async_op=True).wait()
        self.steps += 1
        o = self.opt
        for k, g in self.buckets.grads.items():
            if k not in grads:   # torch.optim skips parameters whose .grad is None (no update, no weight decay)
                continue
            master = self.latent if k == "model.latent_queries" else self.masters[k]
            _bwd.adamw(master.view(-1), None, g.view(-1), self.m[k].view(-1), self.v[k].view(-1), o["lr"], betas=o["betas"],
                       eps=o["eps"], weight_decay=o["weight_decay"], step=self.steps)
        self.s1.refresh()
        self.model._s2.set_latent_queries(self.latent)
        return loss
