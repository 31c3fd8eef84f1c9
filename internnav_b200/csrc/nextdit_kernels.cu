// Row-wise kernels of the NextDiT System 1 (the released DualVLN trajectory head), everything between its GEMMs and
// attention calls.  Reference: internnav/model/basemodel/internvla_n1/nextdit_traj.py L125-178 (LuminaNextDiTBlock.forward:
// LuminaRMSNormZero modulation, tanh-gated RMSNorm residuals), L352-356 (LuminaLayerNormContinuous) and
// internvla_n1.py L399-427 (action encoder + positional code, classifier-free guidance, flow-matching Euler update).
//
// Rows are tokens of width D (384 for the DiT); the per-sample modulation vectors live once per GROUP of rows (all
// Ns * 32 tokens of the sampled trajectories of one environment half), so they are read through a group index instead
// of being materialised per row.  All kernels are HBM-bound: one warp per row, 16-byte loads, the row kept in registers
// between the statistics and the output pass (algorithmic traffic = 1 read + 1 write of the row, + 1 read of the residual).
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "n1_ops.h"

namespace n1 {
namespace {

constexpr int kMaxChunks = 4;  // D <= 1024: each lane holds up to 4 chunks of 8 elements

__device__ __forceinline__ void unpack8(const uint4& u, float (&f)[8]) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x, f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 u;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return u;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// mode 0: out = RMSNorm(x) * w * (1 + scale[g])            (LuminaRMSNormZero / ffn_norm1 with scale_mlp)
// mode 1: out = LayerNorm_noaffine(x) * (1 + scale[g])      (LuminaLayerNormContinuous)
// mode 2: out = res + tanh(gate[g]) * RMSNorm(x) * w        (the two gated residual updates of the block)
// w may be null (no elementwise weight); scale / gate may be null (factor 1).
__global__ void __launch_bounds__(256) mod_norm_kernel(const bf16* __restrict__ x, int ldx, const float* __restrict__ w,
                                                        const bf16* __restrict__ mod, int ld_mod, int rows_per_group,
                                                        const bf16* __restrict__ res, int ldr, bf16* __restrict__ out, int ldo,
                                                        long rows, int D, float eps, int mode) {
  const long row = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float v[kMaxChunks][8];
  float sum = 0.f, sq = 0.f;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < D) {
      unpack8(*reinterpret_cast<const uint4*>(x + row * ldx + col), v[c]);
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += v[c][i], sq += v[c][i] * v[c][i];
    }
  }
  sum = warp_sum(sum), sq = warp_sum(sq);
  const float mean = mode == 1 ? sum / D : 0.f;
  const float var = mode == 1 ? fmaxf(sq / D - mean * mean, 0.f) : sq / D;
  const float rstd = rsqrtf(var + eps);
  const bf16* m = mod ? mod + (row / rows_per_group) * ld_mod : nullptr;
#pragma unroll
  for (int c = 0; c < kMaxChunks; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < D) {
      float f[8], mm[8], rr[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = (v[c][i] - mean) * rstd;
      if (w) {
        const float4 a = *reinterpret_cast<const float4*>(w + col), b = *reinterpret_cast<const float4*>(w + col + 4);
        f[0] *= a.x, f[1] *= a.y, f[2] *= a.z, f[3] *= a.w, f[4] *= b.x, f[5] *= b.y, f[6] *= b.z, f[7] *= b.w;
      }
      if (m) {
        unpack8(*reinterpret_cast<const uint4*>(m + col), mm);
        if (mode == 2) {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] *= tanhf(mm[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] *= 1.f + mm[i];
        }
      }
      if (mode == 2) {
        unpack8(*reinterpret_cast<const uint4*>(res + row * ldr + col), rr);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] += rr[i];
      }
      *reinterpret_cast<uint4*>(out + row * ldo + col) = pack8(f);
    }
  }
}

__global__ void add_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out, long n8) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  float x[8], y[8];
  unpack8(reinterpret_cast<const uint4*>(a)[i], x);
  unpack8(reinterpret_cast<const uint4*>(b)[i], y);
#pragma unroll
  for (int k = 0; k < 8; ++k) x[k] += y[k];
  reinterpret_cast<uint4*>(out)[i] = pack8(x);
}

// feats[r, :] = bf16(W lat[r, :3] + b) + pos[r % T, :]   (action_encoder: nn.Linear(3, D); pos fp32 [T, D])
__global__ void action_embed_kernel(const float* __restrict__ lat, const float* __restrict__ w, const float* __restrict__ b,
                                    const float* __restrict__ pos, bf16* __restrict__ out, long rows, int T, int D) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = D / 8;
  if (i >= rows * per) return;
  const long r = i / per;
  const int col = (int)(i % per) * 8;
  // the latents are kept in the model dtype between steps (FlowMatchEulerDiscreteScheduler.step casts back): bf16 values
  const float x0 = lat[r * 3], x1 = lat[r * 3 + 1], x2 = lat[r * 3 + 2];
  float f[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float* wr = w + (col + k) * 3;
    const float lin = __bfloat162float(__float2bfloat16(wr[0] * x0 + wr[1] * x1 + wr[2] * x2 + b[col + k]));
    f[k] = lin + pos[(r % T) * D + col + k];
  }
  *reinterpret_cast<uint4*>(out + r * D + col) = pack8(f);
}

// pred [rows_total, ld] bf16 (first 3 columns): rows [0, n) unconditional, [n, 2n) conditional when cfg; the update
// lat = bf16(lat + dt * (u + s (c - u))) element-wise in the reference's operation order and dtypes
// (internvla_n1.py L422-427; scheduler step in fp32, result cast back to the model dtype).
__global__ void cfg_euler_kernel(const bf16* __restrict__ pred, int ld, long n, int cfg, float scale, float dt,
                                 float* __restrict__ lat) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const long r = i / 3;
  const int c = (int)(i % 3);
  float p;
  if (cfg) {
    const float u = __bfloat162float(pred[r * ld + c]), cd = __bfloat162float(pred[(n + r) * ld + c]);
    const float diff = __bfloat162float(__float2bfloat16(cd - u));
    const float sc = __bfloat162float(__float2bfloat16(scale * diff));
    p = __bfloat162float(__float2bfloat16(u + sc));
  } else {
    p = __bfloat162float(pred[r * ld + c]);
  }
  const float step = __bfloat162float(__float2bfloat16(dt * p));   // 0-dim fp32 sigma difference x bf16 tensor -> bf16
  lat[i] = __bfloat162float(__float2bfloat16(lat[i] + step));
}

inline unsigned blocks(long n, int per) { return (unsigned)((n + per - 1) / per); }

}  // namespace

void mod_norm(const bf16* x, int ldx, const float* w, const bf16* mod, int ld_mod, int rows_per_group, const bf16* res, int ldr,
              bf16* out, int ldo, long rows, int D, float eps, int mode, cudaStream_t s) {
  N1_CHECK(x && out && rows > 0 && D > 0 && D % 8 == 0 && D <= kMaxChunks * 256, "mod_norm: D must be a multiple of 8, <= 1024");
  N1_CHECK(mode >= 0 && mode <= 2 && (mode != 2 || res), "mod_norm: mode 0 / 1 / 2 (2 needs the residual)");
  N1_CHECK(ldx % 8 == 0 && ldo % 8 == 0 && (!mod || (ld_mod % 8 == 0 && rows_per_group > 0)) && (!res || ldr % 8 == 0),
           "mod_norm: leading dimensions must be multiples of 8");
  mod_norm_kernel<<<blocks(rows, 8), 256, 0, s>>>(x, ldx, w, mod, ld_mod, rows_per_group > 0 ? rows_per_group : 1, res, ldr, out,
                                                  ldo, rows, D, eps, mode);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

void add_bf16(const bf16* a, const bf16* b, bf16* out, long n, cudaStream_t s) {
  N1_CHECK(a && b && out && n > 0 && n % 8 == 0, "add_bf16: n must be a positive multiple of 8");
  add_kernel<<<blocks(n / 8, 256), 256, 0, s>>>(a, b, out, n / 8);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

void action_embed(const float* lat, const float* w, const float* b, const float* pos, bf16* out, long rows, int T, int D,
                  cudaStream_t s) {
  N1_CHECK(lat && w && b && pos && out && rows > 0 && T > 0 && D % 8 == 0, "action_embed: bad arguments");
  action_embed_kernel<<<blocks(rows * (D / 8), 256), 256, 0, s>>>(lat, w, b, pos, out, rows, T, D);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

void cfg_euler(const bf16* pred, int ld, long n, int cfg, float scale, float dt, float* lat, cudaStream_t s) {
  N1_CHECK(pred && lat && n > 0 && ld >= 3, "cfg_euler: bad arguments");
  cfg_euler_kernel<<<blocks(n * 3, 256), 256, 0, s>>>(pred, ld, n, cfg, scale, dt, lat);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace n1
