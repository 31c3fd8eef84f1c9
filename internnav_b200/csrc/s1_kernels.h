// Launchers for the System-1 element-wise / row-wise kernels (s1_kernels.cu).
#pragma once
#include "n1_ops.h"

namespace n1 {

// Per-step DDPM constants: x0 = clamp((x_t - sqrt(1-acp_t) eps) / sqrt(acp_t), -1, 1);
// x_{t-1} = c0 x0 + c1 x_t + sigma z.
struct DdpmCoef {
  float sqrt_one_minus_acp, inv_sqrt_acp, c0, c1, sigma;
};

// exact_norm: fp32 ImageNet mean / std (the stand-alone policy's RGBDBackbone, navdp_backbone.py L233-234) instead of the
// bf16-rounded constants the InternVLA-N1 head holds (navdp_backbone.py L126-127)
void patchify_rgb(const float* img, bf16* out, int n_img, int ldk, cudaStream_t s, bool exact_norm = false);
void patchify_depth(const float* img, bf16* out, int n_img, int ldk, cudaStream_t s);
void fill_cls(bf16* x, const float* cls_pos, int n_img, int tokens, int D, cudaStream_t s);
void vit_out(const bf16* x, bf16* mem, const float* w, const float* b, const float* pe, int n_img, int frames,
             int slot_base, int slots, cudaStream_t s);
void bcast_rows(const bf16* src, bf16* dst, long rows, int period, int D, cudaStream_t s);
void embed_actions(const float* xt, const float* w, const float* bias, const float* pos, bf16* tgt, long rows, int T,
                   cudaStream_t s);
// condition rows [time | goal x goal_slots | memory tokens] + position table; goal == nullptr: zero goal embedding
void build_cond(const int* tsteps, int t_scalar, const bf16* goal, const bf16* rgbd, const float* cpe, bf16* cond, int B,
                int Mtok, int first_slot, int num_slots, cudaStream_t s, int goal_slots = 1);
// critic head of the stand-alone policy: out[r] = critic_w . mean_t LayerNorm(h[r, t, :]) + critic_b   (one warp per sample)
void critic_head(const bf16* h, const float* lw, const float* lb, const float* cw, const float* cb, long samples, int T,
                 float* out, cudaStream_t s);
void head_ddpm(const bf16* h, const float* lw, const float* lb, const float* hw, const float* hb, long rows, int mode,
               float* x, const float* noise, float* eps_out, const DdpmCoef& cf, cudaStream_t s);
void fill_int(int* p, int value, int n, cudaStream_t s);
void f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t s);
void bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t s);

}  // namespace n1
