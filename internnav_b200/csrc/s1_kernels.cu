// Small fused element-wise / row-wise kernels of the System-1 (NavDP) path: everything around the GEMMs that
// the reference runs as strings of tiny PyTorch launches.
#include "n1_ops.h"
#include "n1_ptx.cuh"
#include "s1_kernels.h"

namespace n1 {
namespace {

// ---------------------------------------------------------------------------------- DINOv2 patchify
// navdp_backbone.py L159-166 (HWC -> CHW, ImageNet normalise) + patch_embed.py L69-81 (Conv2d k=14,s=14 == GEMM
// over im2col rows).  Output row = img * 256 + py * 16 + px, column = c * 196 + ky * 14 + kx, zero padded to ldk.
__global__ void patchify_rgb_kernel(const float* __restrict__ img, bf16* __restrict__ out, int n_img, int ldk, int exact) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n_img * 256 * ldk;
  if (idx >= total) return;
  const int col = idx % ldk;
  const long row = idx / ldk;
  float v = 0.f;
  if (col < 588) {
    const int c = col / 196, k = col % 196, ky = k / 14, kx = k % 14;
    const int im = row / 256, p = row % 256, py = p / 16, px = p % 16;
    // bf16 roundings of the ImageNet mean/std: the reference holds these constants in bf16 (navdp_backbone.py L126-127)
    float mean = c == 0 ? 0.484375f : (c == 1 ? 0.455078125f : 0.40625f);
    float stdv = c == 0 ? 0.228515625f : (c == 1 ? 0.2236328125f : 0.224609375f);
    if (exact) {  // fp32 constants of the stand-alone RGBDBackbone
      mean = c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f);
      stdv = c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f);
    }
    const float x = img[(((long)im * 224 + py * 14 + ky) * 224 + px * 14 + kx) * 3 + c];
    v = (x - mean) / stdv;
  }
  out[idx] = __float2bfloat16(v);
}
// Depth is replicated to 3 channels with no normalisation (navdp_backbone.py L176-181); the three identical channels
// are folded into the weight (sum over c at pack time), so the im2col row has 196 columns.
__global__ void patchify_depth_kernel(const float* __restrict__ img, bf16* __restrict__ out, int n_img, int ldk) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long total = (long)n_img * 256 * ldk;
  if (idx >= total) return;
  const int col = idx % ldk;
  const long row = idx / ldk;
  float v = 0.f;
  if (col < 196) {
    const int ky = col / 14, kx = col % 14;
    const int im = row / 256, p = row % 256, py = p / 16, px = p % 16;
    v = img[((long)im * 224 + py * 14 + ky) * 224 + px * 14 + kx];
  }
  out[idx] = __float2bfloat16(v);
}

// x[img, 0, :] = cls + pos[0]   (dinov2.py L219-220)
__global__ void fill_cls_kernel(bf16* __restrict__ x, const float* __restrict__ cls_pos, int n_img, int tokens, int D) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_img * D) return;
  const int im = idx / D, d = idx % D;
  x[(long)im * tokens * D + d] = __float2bfloat16(cls_pos[d]);
}

// Final ViT LayerNorm (eps 1e-6), drop the cls token, scatter into the Q-former memory [B, slots*256, D] and add
// former_pe (dinov2.py L312-314, navdp_backbone.py L166, L181, L191-192).  One warp per output row; D = 384.
__global__ void __launch_bounds__(256) vit_out_kernel(const bf16* __restrict__ x, bf16* __restrict__ mem,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      const float* __restrict__ pe, int n_img, int frames,
                                                      int slot_base, int slots) {
  constexpr int D = 384;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_img * 256) return;
  const int im = warp / 256, p = warp % 256;
  const int env = im / frames, slot = slot_base + im % frames;
  const bf16* xr = x + ((long)im * 257 + 1 + p) * D;
  float v[12];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint2 q = *reinterpret_cast<const uint2*>(xr + (lane + i * 32) * 4);
    v[i * 4 + 0] = bf16_lo(q.x), v[i * 4 + 1] = bf16_hi(q.x), v[i * 4 + 2] = bf16_lo(q.y), v[i * 4 + 3] = bf16_hi(q.y);
    s += v[i * 4] + v[i * 4 + 1] + v[i * 4 + 2] + v[i * 4 + 3];
  }
  const float mean = warp_sum(s) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(sq) / D + 1e-6f);
  const long orow = (long)env * slots * 256 + slot * 256 + p;
  const float* per = pe ? pe + (long)(slot * 256 + p) * D : nullptr;  // null: tokens without former_pe (training)
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const int c = (lane + i * 32) * 4;
    float o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j] = (v[i * 4 + j] - mean) * rstd * w[c + j] + b[c + j] + (per ? per[c + j] : 0.f);
    uint2 pk;
    pk.x = pack_bf16(o[0], o[1]), pk.y = pack_bf16(o[2], o[3]);
    *reinterpret_cast<uint2*>(mem + orow * D + c) = pk;
  }
}

// dst[r, :] = src[r % period, :]  (bf16 broadcast of a learned table over the batch)
__global__ void bcast_rows_kernel(const bf16* __restrict__ src, bf16* __restrict__ dst, long rows, int period, int D) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * D / 8) return;
  const long r = idx / (D / 8);
  const int c = idx % (D / 8);
  reinterpret_cast<uint4*>(dst)[idx] = reinterpret_cast<const uint4*>(src)[(r % period) * (D / 8) + c];
}

// ---------------------------------------------------------------------------------- NavDP denoiser glue
// tgt[r, :] = input_embed(x_t[r]) + out_pos_embed[r % T]          (navdp.py L178, L190)
__global__ void embed_actions_kernel(const float* __restrict__ xt, const float* __restrict__ w /*[384,3]*/,
                                     const float* __restrict__ bias, const float* __restrict__ pos /*[T,384]*/,
                                     bf16* __restrict__ tgt, long rows, int T) {
  constexpr int D = 384;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * (D / 2)) return;
  const long r = idx / (D / 2);
  const int d = (idx % (D / 2)) * 2;
  const float x0 = xt[r * 3], x1 = xt[r * 3 + 1], x2 = xt[r * 3 + 2];
  const float* p = pos + (long)(r % T) * D + d;
  const float a = w[d * 3] * x0 + w[d * 3 + 1] * x1 + w[d * 3 + 2] * x2 + bias[d] + p[0];
  const float b2 = w[d * 3 + 3] * x0 + w[d * 3 + 4] * x1 + w[d * 3 + 5] * x2 + bias[d + 1] + p[1];
  *reinterpret_cast<uint32_t*>(tgt + r * D + d) = pack_bf16(a, b2);
}

// cond[b, 0] = sinusoidal(t_b) + cpe[0]; cond[b, 1] = goal[b] + cpe[1]; cond[b, 2+i] = rgbd[b, i] + cpe[2+i]
// (navdp.py L179-187; SinusoidalPosEmb navdp_backbone.py L14-21: half = 192, freq_j = exp(-j ln(1e4)/191),
// emb = [sin(t f), cos(t f)]).  first_slot/num_slots restrict the update to the time token inside the step loop.
__global__ void build_cond_kernel(const int* __restrict__ tsteps, int t_scalar, const bf16* __restrict__ goal,
                                  const bf16* __restrict__ rgbd, const float* __restrict__ cpe,
                                  bf16* __restrict__ cond, int B, int Mtok, int first_slot, int num_slots, int goal_slots) {
  constexpr int D = 384;
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)B * num_slots * D) return;
  const int d = idx % D;
  const int slot = first_slot + (idx / D) % num_slots;
  const int b = idx / ((long)D * num_slots);
  float v;
  if (slot == 0) {
    const float t = (float)(tsteps ? tsteps[b] : t_scalar);
    const int j = d % 192;
    const float f = expf((float)j * -(logf(10000.0f) / 191.0f));
    v = d < 192 ? sinf(t * f) : cosf(t * f);
  } else if (slot <= goal_slots) {
    v = goal ? __bfloat162float(goal[(long)b * D + d]) : 0.f;
  } else {
    const int first_mem = 1 + goal_slots;
    v = __bfloat162float(rgbd[((long)b * (Mtok - first_mem) + slot - first_mem) * D + d]);
  }
  cond[((long)b * Mtok + slot) * D + d] = __float2bfloat16(v + cpe[slot * D + d]);
}

// Final LayerNorm (eps 1e-5) + action_head 384 -> 3 (navdp.py L193-194), optionally fused with the DDPM ancestral
// update (diffusers 0.33.1 DDPMScheduler.step semantics restated in SURVEY.md App. B; navdp.py L250).
// One warp per row.  mode 0: eps_out[r] = eps_hat.  mode 1: x[r] <- x_{t-1}.
__global__ void __launch_bounds__(256) head_kernel(const bf16* __restrict__ h, const float* __restrict__ lw,
                                                   const float* __restrict__ lb, const float* __restrict__ hw /*[3,384]*/,
                                                   const float* __restrict__ hb, long rows, int mode,
                                                   float* __restrict__ x /*[rows,3] fp32 in/out*/,
                                                   const float* __restrict__ noise /*[rows,3] or null*/,
                                                   float* __restrict__ eps_out, DdpmCoef cf) {
  constexpr int D = 384;
  const long warp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const bf16* hr = h + warp * D;
  float v[12];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const uint2 q = *reinterpret_cast<const uint2*>(hr + (lane + i * 32) * 4);
    v[i * 4 + 0] = bf16_lo(q.x), v[i * 4 + 1] = bf16_hi(q.x), v[i * 4 + 2] = bf16_lo(q.y), v[i * 4 + 3] = bf16_hi(q.y);
    s += v[i * 4] + v[i * 4 + 1] + v[i * 4 + 2] + v[i * 4 + 3];
  }
  const float mean = warp_sum(s) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
  const float rstd = rsqrtf(warp_sum(sq) / D + 1e-5f);
  float e0 = 0.f, e1 = 0.f, e2 = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = (lane + i * 32) * 4 + j;
      const float y = (v[i * 4 + j] - mean) * rstd * lw[c] + lb[c];
      e0 += y * hw[c], e1 += y * hw[D + c], e2 += y * hw[2 * D + c];
    }
  }
  e0 = warp_sum(e0) + hb[0], e1 = warp_sum(e1) + hb[1], e2 = warp_sum(e2) + hb[2];
  if (lane < 3) {
    const float e = lane == 0 ? e0 : (lane == 1 ? e1 : e2);
    if (mode == 0) {
      eps_out[warp * 3 + lane] = e;
    } else {
      const float xt = x[warp * 3 + lane];
      float x0 = (xt - cf.sqrt_one_minus_acp * e) * cf.inv_sqrt_acp;
      x0 = fminf(fmaxf(x0, -1.f), 1.f);
      float xn = cf.c0 * x0 + cf.c1 * xt;
      if (noise) xn += cf.sigma * noise[warp * 3 + lane];
      x[warp * 3 + lane] = xn;
    }
  }
}

__global__ void f32_to_bf16_kernel(const float* __restrict__ s, bf16* __restrict__ d, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __float2bfloat16(s[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ s, float* __restrict__ d, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) d[i] = __bfloat162float(s[i]);
}

inline int nblk(long n, int t) { return (int)((n + t - 1) / t); }


// Critic head of the stand-alone NavDP policy (navdp_policy.py L183-186): LayerNorm (eps 1e-5) of every row, mean over the T
// rows of a sample, dot with critic_head.weight, + bias.  One warp per sample; D = 384 (12 values per lane).
__global__ void __launch_bounds__(256) critic_head_kernel(const bf16* __restrict__ h, const float* __restrict__ lw,
                                                          const float* __restrict__ lb, const float* __restrict__ cw,
                                                          const float* __restrict__ cb, long samples, int T,
                                                          float* __restrict__ out) {
  constexpr int D = 384;
  const long smp = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (smp >= samples) return;
  float acc[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) acc[i] = 0.f;
  for (int t = 0; t < T; ++t) {
    const bf16* xr = h + (smp * T + t) * D;
    float v[12];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      const uint2 q = *reinterpret_cast<const uint2*>(xr + (lane + i * 32) * 4);
      v[i * 4 + 0] = bf16_lo(q.x), v[i * 4 + 1] = bf16_hi(q.x), v[i * 4 + 2] = bf16_lo(q.y), v[i * 4 + 3] = bf16_hi(q.y);
      s += v[i * 4] + v[i * 4 + 1] + v[i * 4 + 2] + v[i * 4 + 3];
    }
    const float mean = warp_sum(s) / D;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(sq) / D + 1e-5f);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = (lane + i * 32) * 4 + j;
        acc[i * 4 + j] += (v[i * 4 + j] - mean) * rstd * lw[c] + lb[c];
      }
  }
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) dot += acc[i * 4 + j] * cw[(lane + i * 32) * 4 + j];
  dot = warp_sum(dot) / T;
  if (lane == 0) out[smp] = dot + cb[0];
}

__global__ void fill_int_kernel(int* p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace

void patchify_rgb(const float* img, bf16* out, int n_img, int ldk, cudaStream_t s, bool exact_norm) {
  const long total = (long)n_img * 256 * ldk;
  patchify_rgb_kernel<<<nblk(total, 256), 256, 0, s>>>(img, out, n_img, ldk, exact_norm ? 1 : 0);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void patchify_depth(const float* img, bf16* out, int n_img, int ldk, cudaStream_t s) {
  const long total = (long)n_img * 256 * ldk;
  patchify_depth_kernel<<<nblk(total, 256), 256, 0, s>>>(img, out, n_img, ldk);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void fill_cls(bf16* x, const float* cls_pos, int n_img, int tokens, int D, cudaStream_t s) {
  fill_cls_kernel<<<nblk((long)n_img * D, 256), 256, 0, s>>>(x, cls_pos, n_img, tokens, D);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void vit_out(const bf16* x, bf16* mem, const float* w, const float* b, const float* pe, int n_img, int frames,
             int slot_base, int slots, cudaStream_t s) {
  vit_out_kernel<<<nblk((long)n_img * 256 * 32, 256), 256, 0, s>>>(x, mem, w, b, pe, n_img, frames, slot_base, slots);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void bcast_rows(const bf16* src, bf16* dst, long rows, int period, int D, cudaStream_t s) {
  bcast_rows_kernel<<<nblk(rows * D / 8, 256), 256, 0, s>>>(src, dst, rows, period, D);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void embed_actions(const float* xt, const float* w, const float* bias, const float* pos, bf16* tgt, long rows, int T,
                   cudaStream_t s) {
  embed_actions_kernel<<<nblk(rows * 192, 256), 256, 0, s>>>(xt, w, bias, pos, tgt, rows, T);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void build_cond(const int* tsteps, int t_scalar, const bf16* goal, const bf16* rgbd, const float* cpe, bf16* cond, int B,
                int Mtok, int first_slot, int num_slots, cudaStream_t s, int goal_slots) {
  build_cond_kernel<<<nblk((long)B * num_slots * 384, 256), 256, 0, s>>>(tsteps, t_scalar, goal, rgbd, cpe, cond, B, Mtok,
                                                                        first_slot, num_slots, goal_slots);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void head_ddpm(const bf16* h, const float* lw, const float* lb, const float* hw, const float* hb, long rows, int mode,
               float* x, const float* noise, float* eps_out, const DdpmCoef& cf, cudaStream_t s) {
  head_kernel<<<nblk(rows * 32, 256), 256, 0, s>>>(h, lw, lb, hw, hb, rows, mode, x, noise, eps_out, cf);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void critic_head(const bf16* h, const float* lw, const float* lb, const float* cw, const float* cb, long samples, int T,
                 float* out, cudaStream_t s) {
  critic_head_kernel<<<nblk(samples * 32, 256), 256, 0, s>>>(h, lw, lb, cw, cb, samples, T, out);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void fill_int(int* p, int value, int n, cudaStream_t s) {
  fill_int_kernel<<<nblk(n, 256), 256, 0, s>>>(p, value, n);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void f32_to_bf16(const float* src, bf16* dst, long n, cudaStream_t s) {
  f32_to_bf16_kernel<<<nblk(n, 256), 256, 0, s>>>(src, dst, n);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void bf16_to_f32(const bf16* src, float* dst, long n, cudaStream_t s) {
  bf16_to_f32_kernel<<<nblk(n, 256), 256, 0, s>>>(src, dst, n);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace n1
