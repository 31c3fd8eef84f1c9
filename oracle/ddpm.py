"""DDPM scheduler restatement -- TEST INFRASTRUCTURE.

The reference calls `diffusers.schedulers.scheduling_ddpm.DDPMScheduler` (diffusers==0.33.1 pinned in
/root/reference/requirements/internvla_n1.txt L3; call sites: internnav/model/basemodel/internvla_n1/navdp.py L3,
L74-76 constructor, L173 add_noise, L247 set_timesteps, L250 step).  diffusers is NOT vendored under /root/reference
and is not installed in this image, so this file restates the published algorithm of that class for the exact
constructor arguments the reference uses:

    DDPMScheduler(num_train_timesteps=N, beta_schedule='squaredcos_cap_v2', clip_sample=True,
                  prediction_type='epsilon')          # defaults: variance_type='fixed_small', clip_sample_range=1.0,
                                                      # timestep_spacing='leading', steps_offset=0, thresholding=False

PARITY UNPINNED against diffusers itself: the reference holds no test / golden vector for the scheduler
and the real package cannot be executed here (not in the image, not in /opt/wheelhouse, no network).  What anchors it
instead (tests/test_oracle_s1.py): (1) the constructor arguments are the ones the reference's own tree documents for this
class (internnav/model/encoder/diffusion_policy/config/*.yaml `noise_scheduler`: squaredcos_cap_v2, fixed_small,
clip_sample True, epsilon); (2) an independent float64 derivation -- the Gaussian posterior q(x_{t-1} | x_t, x_0) of the
forward process by the product-of-Gaussians rule, using none of the formulas below -- reproduces `step` (mean and noise
scale) and `add_noise` at every t; (3) cosine-schedule end points and the t = 0 behaviour.  That pins the arithmetic to
the published algorithm (Ho et al. 2020 eq. 7, Nichol & Dhariwal 2021 eq. 17); what remains unpinned are diffusers'
*defaults* not visible at the call site (timestep_spacing 'leading', steps_offset 0), which for set_timesteps(N) with
N == num_train_timesteps all reduce to t = N-1 .. 0 anyway.  The arithmetic below follows diffusers' scheduling_ddpm.py:
  betas_for_alpha_bar (cosine), set_timesteps ('leading' spacing), _get_variance ('fixed_small', clamp 1e-20),
  step (epsilon prediction, clip to [-1, 1], posterior mean coefficients (formula 7 of Ho et al. 2020)), add_noise.
"""
import math
from types import SimpleNamespace

import numpy as np
import torch


def betas_for_alpha_bar(num_diffusion_timesteps, max_beta=0.999):
    def alpha_bar_fn(t):
        return math.cos((t + 0.008) / 1.008 * math.pi / 2) ** 2

    betas = []
    for i in range(num_diffusion_timesteps):
        t1 = i / num_diffusion_timesteps
        t2 = (i + 1) / num_diffusion_timesteps
        betas.append(min(1 - alpha_bar_fn(t2) / alpha_bar_fn(t1), max_beta))
    return torch.tensor(betas, dtype=torch.float32)


class _StepOutput(SimpleNamespace):
    pass


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_schedule="squaredcos_cap_v2", clip_sample=True,
                 prediction_type="epsilon", variance_type="fixed_small", clip_sample_range=1.0):
        assert beta_schedule == "squaredcos_cap_v2" and prediction_type == "epsilon" and variance_type == "fixed_small"
        self.config = SimpleNamespace(num_train_timesteps=num_train_timesteps, clip_sample=clip_sample,
                                      clip_sample_range=clip_sample_range, prediction_type=prediction_type)
        self.betas = betas_for_alpha_bar(num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())
        # deterministic-noise hook for tests: list of tensors consumed front-to-back by step() when t > 0
        self.noise_queue = None

    def set_timesteps(self, num_inference_steps, device=None):
        assert num_inference_steps <= self.config.num_train_timesteps
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps  # 'leading' spacing, steps_offset 0
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def previous_timestep(self, timestep):
        n = self.num_inference_steps if self.num_inference_steps else self.config.num_train_timesteps
        return timestep - self.config.num_train_timesteps // n

    def _get_variance(self, t):
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        current_beta_t = 1 - alpha_prod_t / alpha_prod_t_prev
        variance = (1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t
        return torch.clamp(variance, min=1e-20)

    def step(self, model_output, timestep, sample, generator=None, return_dict=True):
        t = int(timestep)
        prev_t = self.previous_timestep(t)
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        if self.config.clip_sample:
            pred_original_sample = pred_original_sample.clamp(-self.config.clip_sample_range,
                                                              self.config.clip_sample_range)
        pred_original_sample_coeff = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        current_sample_coeff = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        pred_prev_sample = pred_original_sample_coeff * pred_original_sample + current_sample_coeff * sample
        variance = 0
        if t > 0:
            if self.noise_queue is not None:
                variance_noise = self.noise_queue.pop(0).to(model_output.device, model_output.dtype)
            else:
                variance_noise = torch.randn(model_output.shape, generator=generator, device=model_output.device,
                                             dtype=model_output.dtype)
            variance = (self._get_variance(t) ** 0.5) * variance_noise
        pred_prev_sample = pred_prev_sample + variance
        return _StepOutput(prev_sample=pred_prev_sample, pred_original_sample=pred_original_sample)

    def add_noise(self, original_samples, noise, timesteps):
        acp = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        timesteps = timesteps.to(original_samples.device)
        s1 = acp[timesteps] ** 0.5
        s2 = (1 - acp[timesteps]) ** 0.5
        while s1.dim() < original_samples.dim():
            s1, s2 = s1.unsqueeze(-1), s2.unsqueeze(-1)
        return s1 * original_samples + s2 * noise

    def coef_table(self):
        """[N, 5] float32 {sqrt(1-acp_t), 1/sqrt(acp_t), c0, c1, sigma} -- the table the CUDA path consumes."""
        out = []
        for t in range(self.config.num_train_timesteps):
            a_t = self.alphas_cumprod[t]
            a_p = self.alphas_cumprod[t - 1] if t > 0 else self.one
            cur_a = a_t / a_p
            out.append([float((1 - a_t) ** 0.5), float(1.0 / a_t ** 0.5), float(a_p ** 0.5 * (1 - cur_a) / (1 - a_t)),
                        float(cur_a ** 0.5 * (1 - a_p) / (1 - a_t)),
                        float(self._get_variance(t) ** 0.5) if t > 0 else 0.0])
        return np.asarray(out, dtype=np.float32)
