// LayerNorm / RMSNorm over the last dimension, bf16 in / bf16 out, fp32 statistics, one warp per row with
// 16-byte loads and warp-shuffle reductions.  Replaces nn.LayerNorm (navdp.py L57-66, L78; dinov2.py L98;
// navdp_backbone.py L148) and Qwen2RMSNorm (transformers modeling_qwen2_5_vl.py L57-74).
#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {
namespace {

// D <= 8 * 32 * MAXV elements; each lane holds MAXV 16-byte vectors (8 bf16) of the row in registers.
template <int MAXV>
__global__ void __launch_bounds__(256) norm_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy,
                                                   const float* __restrict__ w, const float* __restrict__ b, int rows,
                                                   int D, float eps, int rms) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const bf16* xr = x + (long)warp * ldx;
  bf16* yr = y + (long)warp * ldy;
  const int nvec = D >> 3;
  float v[MAXV][8];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      const uint4 q = *reinterpret_cast<const uint4*>(xr + vi * 8);
      v[i][0] = bf16_lo(q.x), v[i][1] = bf16_hi(q.x), v[i][2] = bf16_lo(q.y), v[i][3] = bf16_hi(q.y);
      v[i][4] = bf16_lo(q.z), v[i][5] = bf16_hi(q.z), v[i][6] = bf16_lo(q.w), v[i][7] = bf16_hi(q.w);
#pragma unroll
      for (int j = 0; j < 8; ++j) sum += v[i][j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[i][j] = 0.f;
    }
  }
  float mean = 0.f;
  if (!rms) mean = warp_sum(sum) / D;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[i][j] - mean;
        sq += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / D + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float o[8];
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(w + vi * 8));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(w + vi * 8 + 4));
      const float ww[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * ww[j];
      if (b) {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + vi * 8));
        const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + vi * 8 + 4));
        o[0] += b0.x, o[1] += b0.y, o[2] += b0.z, o[3] += b0.w;
        o[4] += b1.x, o[5] += b1.y, o[6] += b1.z, o[7] += b1.w;
      }
      uint4 pk;
      pk.x = pack_bf16(o[0], o[1]), pk.y = pack_bf16(o[2], o[3]);
      pk.z = pack_bf16(o[4], o[5]), pk.w = pack_bf16(o[6], o[7]);
      *reinterpret_cast<uint4*>(yr + vi * 8) = pk;
    }
  }
}

// D = 384 LayerNorm (the NavDP decoder / DINOv2 width): RPW rows per warp with all their loads issued up front (more
// bytes in flight than one row per warp), 8-byte vectors so the 96 vectors of a row split evenly over the 32 lanes,
// weights read once per warp.
template <int RPW>
__global__ void __launch_bounds__(256) ln384_kernel(const bf16* __restrict__ x, int ldx, bf16* __restrict__ y, int ldy,
                                                    const float* __restrict__ w, const float* __restrict__ b, int rows,
                                                    float eps) {
  constexpr int D = 384;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const long row0 = (long)warp * RPW;
  if (row0 >= rows) return;
  uint2 q[RPW][3];
#pragma unroll
  for (int r = 0; r < RPW; ++r)
#pragma unroll
    for (int i = 0; i < 3; ++i)
      q[r][i] = (row0 + r < rows) ? *reinterpret_cast<const uint2*>(x + (row0 + r) * ldx + (lane + i * 32) * 4)
                                  : make_uint2(0u, 0u);
  float ww[12], bb[12];
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(w + (lane + i * 32) * 4));
    const float4 c = __ldg(reinterpret_cast<const float4*>(b + (lane + i * 32) * 4));
    ww[i * 4] = a.x, ww[i * 4 + 1] = a.y, ww[i * 4 + 2] = a.z, ww[i * 4 + 3] = a.w;
    bb[i * 4] = c.x, bb[i * 4 + 1] = c.y, bb[i * 4 + 2] = c.z, bb[i * 4 + 3] = c.w;
  }
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    float v[12];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      v[i * 4] = bf16_lo(q[r][i].x), v[i * 4 + 1] = bf16_hi(q[r][i].x);
      v[i * 4 + 2] = bf16_lo(q[r][i].y), v[i * 4 + 3] = bf16_hi(q[r][i].y);
      s += v[i * 4] + v[i * 4 + 1] + v[i * 4 + 2] + v[i * 4 + 3];
    }
    const float mean = warp_sum(s) * (1.0f / D);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) sq += (v[i] - mean) * (v[i] - mean);
    const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    if (row0 + r < rows) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        uint2 o;
        o.x = pack_bf16((v[i * 4] - mean) * rstd * ww[i * 4] + bb[i * 4],
                        (v[i * 4 + 1] - mean) * rstd * ww[i * 4 + 1] + bb[i * 4 + 1]);
        o.y = pack_bf16((v[i * 4 + 2] - mean) * rstd * ww[i * 4 + 2] + bb[i * 4 + 2],
                        (v[i * 4 + 3] - mean) * rstd * ww[i * 4 + 3] + bb[i * 4 + 3]);
        *reinterpret_cast<uint2*>(y + (row0 + r) * ldy + (lane + i * 32) * 4) = o;
      }
    }
  }
}

}  // namespace

void layernorm(const bf16* x, int ldx, bf16* y, int ldy, const float* w, const float* b, int rows, int D, float eps,
               int rms, cudaStream_t stream) {
  if (rows <= 0) return;
  N1_CHECK(D % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0, "norm: D and leading dims must be multiples of 8");
  N1_CHECK(D <= 8 * 32 * 20, "norm: D too large");
  const int threads = 256;
  if (D == 384 && !rms && b != nullptr) {
    constexpr int RPW = 4;
    const int warps = (rows + RPW - 1) / RPW;
    ln384_kernel<RPW><<<(warps * 32 + threads - 1) / threads, threads, 0, stream>>>(x, ldx, y, ldy, w, b, rows, eps);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
    return;
  }
  const int blocks = (rows * 32 + threads - 1) / threads;
  if (D <= 8 * 32 * 2)
    norm_kernel<2><<<blocks, threads, 0, stream>>>(x, ldx, y, ldy, w, b, rows, D, eps, rms);
  else if (D <= 8 * 32 * 6)
    norm_kernel<6><<<blocks, threads, 0, stream>>>(x, ldx, y, ldy, w, b, rows, D, eps, rms);
  else if (D <= 8 * 32 * 14)
    norm_kernel<14><<<blocks, threads, 0, stream>>>(x, ldx, y, ldy, w, b, rows, D, eps, rms);
  else
    norm_kernel<20><<<blocks, threads, 0, stream>>>(x, ldx, y, ldy, w, b, rows, D, eps, rms);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace n1
