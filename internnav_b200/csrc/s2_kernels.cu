// Glue kernels of the System-2 path (Qwen2.5-VL ViT + decoder prefill): row gathers, rotary tables and application,
// embedding splice.  Reference arithmetic: transformers==4.51.0 modeling_qwen2_5_vl.py (un-vendored pin,
// requirements/internvla_n1.txt L7) as called from internvla_n1.py L320-347; restated in oracle/qwen_oracle.py.
#include "s2_kernels.h"

#include "n1_ptx.cuh"

namespace n1 {
namespace {

__global__ void gather_rows_kernel(const bf16* __restrict__ src, const int* __restrict__ idx, bf16* __restrict__ dst,
                                   long rows, int group, int vec_per_row) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * vec_per_row) return;
  const long r = i / vec_per_row;
  const int c = i % vec_per_row;
  const long sr = (long)idx[r / group] * group + r % group;
  reinterpret_cast<uint4*>(dst)[i] = __ldg(reinterpret_cast<const uint4*>(src) + sr * vec_per_row + c);
}

__global__ void vit_rope_table_kernel(const int* __restrict__ pos_hw, float2* __restrict__ cs, long tokens, int half,
                                      float theta) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * half) return;
  const long t = i / half;
  const int j = i % half;
  const int q = half / 2;  // frequencies per axis
  const int p = pos_hw[t * 2 + (j < q ? 0 : 1)];
  // inv_freq = 1 / theta^(arange(0, dim, 2) / dim) with dim = half (Qwen2_5_VisionRotaryEmbedding(head_dim // 2))
  const float inv = 1.0f / powf(theta, (float)(2 * (j % q)) / (float)half);
  float sn, cn;
  sincosf((float)p * inv, &sn, &cn);
  cs[i] = make_float2(cn, sn);
}

__global__ void mrope_table_kernel(const int* __restrict__ pos3, float2* __restrict__ cs, long tokens, int half,
                                   int sec_t, int sec_h, float theta) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * half) return;
  const long t = i / half;
  const int j = i % half;
  const int stream = j < sec_t ? 0 : (j < sec_t + sec_h ? 1 : 2);
  const int p = pos3[(long)stream * tokens + t];
  const float inv = 1.0f / powf(theta, (float)(2 * j) / (float)(2 * half));
  float sn, cn;
  sincosf((float)p * inv, &sn, &cn);
  cs[i] = make_float2(cn, sn);
}

// 8 rotation pairs (two 16-byte vectors) per thread; half % 8 == 0.
__global__ void apply_rope_kernel(bf16* __restrict__ x, int ld, const float2* __restrict__ cs, long tokens, int heads,
                                  int half) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int vec = half / 8;
  const long per_tok = (long)heads * vec;
  if (i >= tokens * per_tok) return;
  const long t = i / per_tok;
  const int h = (i % per_tok) / vec, j = (i % vec) * 8;
  bf16* p = x + t * ld + (long)h * 2 * half + j;
  uint4 a = *reinterpret_cast<const uint4*>(p), b = *reinterpret_cast<const uint4*>(p + half);
  const float4* c4 = reinterpret_cast<const float4*>(cs + t * half + j);  // (cos, sin) pairs
  uint32_t* au = reinterpret_cast<uint32_t*>(&a);
  uint32_t* bu = reinterpret_cast<uint32_t*>(&b);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float4 c = __ldg(c4 + k);  // c.x,c.y = cos,sin of pair 2k ; c.z,c.w of pair 2k+1
    const float a0 = bf16_lo(au[k]), a1 = bf16_hi(au[k]), b0 = bf16_lo(bu[k]), b1 = bf16_hi(bu[k]);
    au[k] = pack_bf16(a0 * c.x - b0 * c.y, a1 * c.z - b1 * c.w);
    bu[k] = pack_bf16(b0 * c.x + a0 * c.y, b1 * c.z + a1 * c.w);
  }
  *reinterpret_cast<uint4*>(p) = a;
  *reinterpret_cast<uint4*>(p + half) = b;
}

__global__ void build_embeds_kernel(const int* __restrict__ kind, const int* __restrict__ src,
                                    const bf16* __restrict__ emb, const bf16* __restrict__ img,
                                    const bf16* __restrict__ lat, bf16* __restrict__ out, long tokens, int vec_per_row) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tokens * vec_per_row) return;
  const long t = i / vec_per_row;
  const int c = i % vec_per_row;
  const int k = kind[t];
  const bf16* base = k == 0 ? emb : (k == 1 ? img : lat);
  reinterpret_cast<uint4*>(out)[i] = __ldg(reinterpret_cast<const uint4*>(base) + (long)src[t] * vec_per_row + c);
}

inline int nblk(long n) { return (int)((n + 255) / 256); }

}  // namespace

void gather_rows(const bf16* src, const int* idx, bf16* dst, long rows, int group, int cols, cudaStream_t s) {
  N1_CHECK(cols % 8 == 0, "gather_rows: cols % 8");
  gather_rows_kernel<<<nblk(rows * (cols / 8)), 256, 0, s>>>(src, idx, dst, rows, group, cols / 8);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void vit_rope_table(const int* pos_hw, float2* cs, long tokens, int half, float theta, cudaStream_t s) {
  vit_rope_table_kernel<<<nblk(tokens * half), 256, 0, s>>>(pos_hw, cs, tokens, half, theta);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void mrope_table(const int* pos3, float2* cs, long tokens, int half, int sec_t, int sec_h, float theta,
                 cudaStream_t s) {
  mrope_table_kernel<<<nblk(tokens * half), 256, 0, s>>>(pos3, cs, tokens, half, sec_t, sec_h, theta);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void apply_rope(bf16* x, int ld, const float2* cs, long tokens, int heads, int hd, cudaStream_t s) {
  N1_CHECK((hd / 2) % 8 == 0 && ld % 8 == 0, "apply_rope: head_dim / 2 and ld must be multiples of 8");
  apply_rope_kernel<<<nblk(tokens * heads * (hd / 16)), 256, 0, s>>>(x, ld, cs, tokens, heads, hd / 2);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}
void build_embeds(const int* kind, const int* src, const bf16* embed_tokens, const bf16* image_feats,
                  const bf16* latent_queries, bf16* out, long tokens, int H, cudaStream_t s) {
  N1_CHECK(H % 8 == 0, "build_embeds: H % 8");
  build_embeds_kernel<<<nblk(tokens * (H / 8)), 256, 0, s>>>(kind, src, embed_tokens, image_feats, latent_queries, out,
                                                            tokens, H / 8);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace n1
