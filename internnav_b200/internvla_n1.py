"""Model-level mirror of the reference InternVLA-N1 dual-system model and its policy wrapper.

`InternVLAN1ForCausalLM` keeps the hot-path API surface of the reference class
(internnav/model/basemodel/internvla_n1/internvla_n1.py L39-441): `generate_latents(input_ids, pixel_values,
image_grid_thw)`, `generate_traj(traj_latents, images_dp, depths_dp, ...)`, attributes `.visual`, `.model.navdp`,
`.config.{system1, n_query}`, `.get_model()`, `.device`.  Both systems execute in libn1b200.so; PyTorch carries the
tensors.  Everything is batched over B independent environments, each treated exactly like the reference's single one.

`generate()` is the greedy decode of System 2 (`model.generate(do_sample=False)`, internvla_n1_policy.py L169-176) on a
KV cache; `generate_with_latents()` additionally extends that cache by the TRAJ tokens, which replaces the reference's
second full prefill in `generate_latents(output_ids, ...)` (policy L187-190).

`forward()` is the training forward of the navdp_async branch on a collated batch (the backward lives in train_step.py).

`system1 = "nextdit_async"` (the released DualVLN head: trajectory DiT + flow matching, nextdit.py) is served through the
same `generate_traj`; its training branch is not built (the training step is the navdp_async one, train_step.py).
"""
from types import SimpleNamespace

import numpy as np
import torch

from .navdp import NavDP_Policy_DPT_CriticSum_DAT
from .postprocess import batched_traj_to_actions, s1_action_list
from .qwen import EOS_TOKEN_IDS, PAD_TOKEN_ID, QWEN25VL_7B, System2

TRAJ_TOKEN_INDEX = 151667
IMAGE_TOKEN_INDEX = 151655


class _Model:
    """Stands in for `InternVLAN1Model` (= `.model` / `.get_model()` of the reference class)."""

    def __init__(self, navdp, s2, config, nextdit=None):
        self.navdp = navdp          # system1 = "navdp_async"
        self.nextdit = nextdit      # system1 = "nextdit_async": traj_dit, cond_projector, rgb_model, memory_encoder, ... in one object
        self._s2 = s2
        self.config = config
        self.device = s2.device

    @property
    def latent_queries(self):
        raise AttributeError("latent_queries live inside the n1b200 handle; pass them through load_state_dict")


class InternVLAN1ForCausalLM:
    def __init__(self, cfg=None, device="cuda:0", system1="navdp_async", predict_size=32, memory_size=2):
        if system1 not in ("navdp_async", "nextdit_async"):
            raise NotImplementedError("n1b200 implements system1 = 'navdp_async' and 'nextdit_async' (the reference's two "
                                      "asynchronous System-1 heads); got %r" % (system1,))
        self.cfg = dict(QWEN25VL_7B if cfg is None else cfg)
        self.device = torch.device(device)
        self.config = SimpleNamespace(system1=system1, n_query=self.cfg["n_query"], use_cache=True,
                                      hidden_size=self.cfg["hidden"], image_token_id=IMAGE_TOKEN_INDEX)
        self._s2 = System2(self.cfg, device=device)
        navdp = nextdit = None
        if "navdp" in system1:
            navdp = NavDP_Policy_DPT_CriticSum_DAT(memory_size=memory_size, predict_size=predict_size,
                                                   vlm_token_dim=self.cfg["hidden"], navdp_version=0.1, device=device,
                                                   n_query=self.cfg["n_query"])
        else:
            from .nextdit import NextDiTSystem1
            nextdit = NextDiTSystem1(device=device)
        self.model = _Model(navdp, self._s2, self.config, nextdit)

    @classmethod
    def from_pretrained(cls, model_path, torch_dtype=None, attn_implementation=None, device_map=None, device=None, **kw):
        """`InternVLAN1ForCausalLM.from_pretrained(model_path, torch_dtype=torch.bfloat16, attn_implementation=
        "flash_attention_2", device_map={"": device})` (internvla_n1_policy.py L33-38): reads config.json and the weight
        shards of a checkpoint directory and packs them into the library.  `torch_dtype` / `attn_implementation` are
        accepted for signature compatibility (the kernels are bf16 with fp32 accumulation; attention is the library's)."""
        from .checkpoint import read_checkpoint
        if device is None:
            device = (device_map or {}).get("", "cuda:0") if isinstance(device_map, dict) else (device_map or "cuda:0")
        cfg, conf, sd = read_checkpoint(model_path)
        model = cls(cfg, device=str(device), system1=conf.get("system1", "navdp_async"),
                    predict_size=int(conf.get("predict_step_nums", 32)))
        model.load_state_dict(sd)
        model.name_or_path = model_path
        return model

    # ------------------------------------------------------------------ reference-shaped accessors
    def get_model(self):
        return self.model

    def get_n_query(self):
        return self.config.n_query

    def get_system1_type(self):
        return self.config.system1

    def eval(self):
        return self

    def visual(self, pixel_values, grid_thw):
        grid = grid_thw.tolist() if torch.is_tensor(grid_thw) else grid_thw
        return self._s2.visual(pixel_values, grid)

    def load_state_dict(self, state_dict, strict=True):
        """Full-model state_dict in the reference checkpoint layout: `visual.*`, `model.*` (incl. `model.latent_queries`
        and `model.navdp.*`).  `lm_head.*` is ignored (generate_latents never uses it)."""
        navdp_sd = {k[len("model.navdp."):]: v for k, v in state_dict.items() if k.startswith("model.navdp.")}
        rest = {k: v for k, v in state_dict.items() if not k.startswith("model.navdp.")}
        if self.model.nextdit is not None:
            # the NextDiT modules hang directly off `.model` in the reference (internvla_n1_arch.py L131-145)
            heads = ("cond_projector.", "rgb_model.", "memory_encoder.", "rgb_resampler.", "action_encoder.", "action_decoder.",
                     "traj_dit.")
            dit_sd = {k[len("model."):]: v for k, v in rest.items() if k.startswith("model.") and k[len("model."):].startswith(heads)}
            rest = {k: v for k, v in rest.items() if not (k.startswith("model.") and k[len("model."):].startswith(heads))}
            self.model.nextdit.load_state_dict(dit_sd)
        self._s2.load_state_dict(rest)
        if navdp_sd:
            self.model.navdp.load_state_dict(navdp_sd)

    def load_parts(self, s2_state_dict, system1_state_dict):
        self._s2.load_state_dict(s2_state_dict)
        (self.model.navdp if self.model.nextdit is None else self.model.nextdit).load_state_dict(system1_state_dict)

    # ------------------------------------------------------------------ hot path
    @staticmethod
    def _prompts(input_ids):
        if torch.is_tensor(input_ids):
            return [row.tolist() for row in input_ids]
        return [list(map(int, p)) for p in input_ids]

    def generate_latents(self, input_ids, pixel_values, image_grid_thw):
        """internvla_n1.py L320-347.  input_ids: [B, S] tensor (equal lengths) or a list of B token-id lists (ragged);
        pixel_values [sum patches, 1176]; image_grid_thw [n_img, 3] in prompt order -> [B, n_query, hidden]."""
        grid = image_grid_thw.tolist() if torch.is_tensor(image_grid_thw) else image_grid_thw
        with torch.no_grad():
            return self._s2.generate_latents(self._prompts(input_ids), pixel_values, grid)

    def generate_traj(self, traj_latents, images_dp, depths_dp=None, predict_step_nums=32, guidance_scale=1.0,
                      num_inference_steps=10, num_sample_trajs=32, x_init=None, step_noise=None):
        """internvla_n1.py L349-441: the nextdit branch (L359-432, flow-matching Euler over the trajectory DiT with
        classifier-free guidance) or the navdp branch (L434-441) -> [num_sample_trajs * B, predict_size, 3]."""
        if self.model.nextdit is not None:
            return self.model.nextdit.generate_traj(traj_latents.to(self.device), images_dp, depths_dp,
                                                    predict_step_nums=predict_step_nums, guidance_scale=guidance_scale,
                                                    num_inference_steps=num_inference_steps,
                                                    num_sample_trajs=num_sample_trajs, x_init=x_init)
        return self.model.navdp.predict_pointgoal_action_async(
            traj_latents.to(self.device), images_dp, depths_dp, sample_num=num_sample_trajs, x_init=x_init,
            step_noise=step_noise)

    def dual_system_step(self, input_ids, pixel_values, image_grid_thw, images_dp, depths_dp, x_init=None,
                         step_noise=None):
        """One full policy step for B environments: System-2 latent plan -> System-1 trajectories -> action ids.
        Returns (trajectories [32B, T, 3], list of B action lists (<= 4 non-zero ids each, [] => action -1))."""
        lat = self.generate_latents(input_ids, pixel_values, image_grid_thw)
        traj = self.generate_traj(lat, images_dp, depths_dp, x_init=x_init, step_noise=step_noise)
        acts = [s1_action_list(a) for a in batched_traj_to_actions(traj, lat.shape[0], max_actions=4)]
        return traj, acts

    def s1_training_loss(self, traj_hidden_states, traj_images, traj_depths, traj_poses, video_frame_num, noise=None,
                         timesteps=None):
        """The navdp_async branch of the training forward, from the gathered latent states on (internvla_n1.py
        L231-303): traj_hidden_states [B, n_query, H] (the 4 states at t_s_pos); traj_images [B, f, 224, 224, 3];
        traj_depths [B, f, 224, 224]; traj_poses [B, f, 32, 3]; video_frame_num [B] -> scalar loss (the forward; the backward lives in train_step.py)."""
        B, f = traj_images.shape[:2]
        hs = traj_hidden_states.unsqueeze(1).repeat(1, f, 1, 1).flatten(0, 1)
        loss_mask = torch.arange(f, device=traj_images.device).expand(B, f) < video_frame_num.to(traj_images.device).unsqueeze(1)
        cur_images, cur_depths = traj_images.flatten(0, 1), traj_depths.flatten(0, 1)
        goal_images = traj_images[:, 0:1].repeat(1, f, 1, 1, 1).flatten(0, 1)
        goal_depths = traj_depths[:, 0:1].repeat(1, f, 1, 1).flatten(0, 1)
        images_dp = torch.stack([goal_images, cur_images], dim=1)
        depths_dp = torch.stack([goal_depths, cur_depths], dim=1).unsqueeze(-1)
        pred, eps = self.model.navdp.forward_vlm_traj(hs, images_dp, depths_dp, traj_poses, noise=noise,
                                                      timesteps=timesteps)
        err = (pred.float() - eps.float()).square()
        mask = loss_mask.flatten(0, 1)[:, None, None].to(err.device)
        return (err * mask).sum() / mask.sum() / (err.shape[1] * err.shape[2])

    def generate(self, input_ids=None, pixel_values=None, image_grid_thw=None, max_new_tokens=128, do_sample=False,
                 eos_token_id=None, pad_token_id=None, return_dict_in_generate=False, with_latents=False, **hf_kwargs):
        """`model.generate(**inputs, max_new_tokens=128, do_sample=False, use_cache=True, past_key_values=None,
        return_dict_in_generate=True).sequences` (internvla_n1_policy.py L169-176) for B prompts.  `sequences` is
        [B, S_max + longest generation] int64: prompt, generated ids (the eos id included), then `pad_token_id`; ragged
        prompts are right-aligned the way the HF processor pads them (left padding).  Other HF keyword arguments
        (attention_mask, use_cache, past_key_values, ...) are accepted and have no effect on greedy search."""
        if do_sample or hf_kwargs.get("num_beams", 1) != 1:
            raise NotImplementedError("n1b200 implements greedy search only (the reference calls do_sample=False)")
        eos = EOS_TOKEN_IDS if eos_token_id is None else \
            (tuple(eos_token_id) if isinstance(eos_token_id, (list, tuple)) else (int(eos_token_id),))
        pad = PAD_TOKEN_ID if pad_token_id is None else int(pad_token_id)
        prompts = self._prompts(input_ids)
        grid = image_grid_thw.tolist() if torch.is_tensor(image_grid_thw) else image_grid_thw
        with torch.no_grad():
            toks, lat, passes = self._s2.generate(prompts, pixel_values, grid, max_new_tokens=int(max_new_tokens),
                                                  eos_token_ids=eos, pad_token_id=pad, with_latents=with_latents)
        s_max = max(len(p) for p in prompts)
        g_max = max(len(t) for t in toks)
        seq = torch.full((len(prompts), s_max + g_max), pad, dtype=torch.int64)
        for b, (p, t) in enumerate(zip(prompts, toks)):
            seq[b, s_max - len(p):s_max] = torch.tensor(p, dtype=torch.int64)
            seq[b, s_max:s_max + len(t)] = torch.tensor(t, dtype=torch.int64)
        out = SimpleNamespace(sequences=seq.to(self.device), generated=toks, latents=lat, decode_passes=passes)
        return out if (return_dict_in_generate or with_latents) else out.sequences

    def generate_with_latents(self, input_ids, pixel_values, image_grid_thw, max_new_tokens=128, **kw):
        """One System-2 call of the dual-system policy (internvla_n1_policy.py L166-195): greedy answer tokens AND the
        latent plan `generate_latents(output_ids, pixel_values, image_grid_thw)`, sharing one vision pass, one prefill
        and the decode's K/V cache.  -> namespace(sequences, generated, latents [B, n_query, hidden], decode_passes)."""
        return self.generate(input_ids, pixel_values, image_grid_thw, max_new_tokens=max_new_tokens, with_latents=True,
                             return_dict_in_generate=True, **kw)

    def traj_hidden_states(self, input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos):
        """System-2 half of the training forward (internvla_n1.py L128-235) for a collated batch
        (internvla_n1_lerobot_dataset.py L1155-1277: every sample ends with n_query TRAJ tokens at t_s_pos[b], then
        padding; attention_mask = input_ids != pad).  Causal attention makes the states at the TRAJ positions a function
        of the unmasked tokens before them only, and masked tokens are invisible as keys and skipped by get_rope_index,
        so this is the latent-plan prefill over the unmasked prefix of each sample.  -> [B, n_query, hidden]"""
        rows = self._prompts(input_ids)
        nq = self.config.n_query
        mask = None if attention_mask is None else \
            (attention_mask.tolist() if torch.is_tensor(attention_mask) else attention_mask)
        prompts = []
        for b, row in enumerate(rows):
            t = int(t_s_pos[b])
            if row[t:t + nq] != [TRAJ_TOKEN_INDEX] * nq:
                raise ValueError("sample %d: input_ids[t_s_pos : t_s_pos + n_query] are not the TRAJ tokens" % b)
            keep = [True] * len(row) if mask is None else [bool(v) for v in mask[b]]
            if any(tok == IMAGE_TOKEN_INDEX and keep[i] for i, tok in enumerate(row) if i >= t + nq):
                raise NotImplementedError("image tokens after the TRAJ tokens (not produced by the reference collator)")
            prompts.append([tok for i, tok in enumerate(row[:t]) if keep[i]])
        return self.generate_latents(prompts, pixel_values, image_grid_thw)

    def forward(self, input_ids=None, attention_mask=None, labels=None, t_s_pos=None, pixel_values=None,
                image_grid_thw=None, traj_images=None, traj_depths=None, video_frame_num=None, traj_poses=None,
                noise=None, timesteps=None, **hf_kwargs):
        """Training forward of the navdp_async branch (internvla_n1.py L58-318) on a collated batch -> namespace(loss,
        logits=None, traj_hidden_states).  This call is the forward (loss value, no autograd graph); the optimisation step --
        backward kernels, gradient exchange, AdamW -- is `train_step.DualSystemTrainer.step`; `logits` (computed but unused by this branch in the reference, L229) are not produced.
        `noise` / `timesteps` inject the draws of `sample_noise` (navdp.py L165-175) for parity tests."""
        if labels is None or t_s_pos is None or traj_images is None:
            raise NotImplementedError("forward() implements the training branch (labels + t_s_pos + traj_* given); for "
                                      "inference use generate() / generate_latents() / generate_traj()")
        hs = self.traj_hidden_states(input_ids, attention_mask, pixel_values, image_grid_thw, t_s_pos)
        loss = self.s1_training_loss(hs, traj_images, traj_depths, traj_poses, video_frame_num, noise=noise,
                                     timesteps=timesteps)
        return SimpleNamespace(loss=loss, logits=None, traj_hidden_states=hs)

    __call__ = forward


class S1Output(SimpleNamespace):
    pass


class InternVLAN1Net:
    """Policy wrapper with the reference's System-1 entry point (internvla_n1_policy.py L200-215)."""

    def __init__(self, model, continuous_traj=True):
        self.model = model
        self.continuous_traj = continuous_traj

    def eval(self):
        return self

    def reset(self):
        pass

    def s1_step_latent(self, rgb, depth, latent):
        with torch.no_grad():
            dp_actions = self.model.generate_traj(traj_latents=latent, images_dp=rgb, depths_dp=depth)
        B = latent.shape[0]
        lists = batched_traj_to_actions(dp_actions, B, max_actions=4) if self.continuous_traj else None
        if lists is None:
            raise NotImplementedError("chunk_token sampling path: use postprocess.chunk_token on a chosen sample")
        outs = [S1Output(idx=s1_action_list(a)) for a in lists]
        return outs[0] if B == 1 else outs


def preprocess_s1_inputs(rgb_u8, depth_m, goal_rgb_u8, goal_depth_m, depth_clip=5.0):
    """S1 input preprocessing of the agent (internvla_n1_agent.py L308-334): uint8 RGB -> [0,1], depth metres clipped,
    stacked as [goal frame, current frame].  Inputs are already 224x224 numpy arrays (resize happens upstream, PIL)."""
    def f(x):
        return np.asarray(x, dtype=np.float32) / 255.0
    d0 = np.minimum(np.asarray(goal_depth_m, dtype=np.float32), depth_clip)
    d1 = np.minimum(np.asarray(depth_m, dtype=np.float32), depth_clip)
    rgbs = torch.from_numpy(np.stack([f(goal_rgb_u8), f(rgb_u8)])).unsqueeze(0)
    depths = torch.from_numpy(np.stack([d0, d1])).unsqueeze(0).unsqueeze(-1)
    return rgbs, depths
