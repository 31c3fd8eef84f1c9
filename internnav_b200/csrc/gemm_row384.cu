// Full-row GEMM for N = 384 with an optional fused LayerNorm epilogue:
//
//     x_new = [gamma *] (A[M,K] @ W[384,K]^T + bias) + residual          (bf16, the new residual stream)
//     ln    = LayerNorm(x_new) * ln_w + ln_b                              (bf16, the next block's input)   [optional]
//
// One CTA owns whole 128 x 384 output rows (TMEM columns 0..383, three tcgen05.mma N=128 per k-step), so
//   * the activation panel A is read ONCE instead of once per 128-wide N tile (the K <= 1536 GEMMs of the NavDP
//     decoder are L2->SMEM-bound, profiles/r1_ncu_small_v0_summary.txt), and
//   * the row statistics of the following LayerNorm are thread-local in the epilogue (one TMEM lane = one row),
//     which removes the stand-alone LayerNorm launch and its HBM round trip.
// Replaces `x = x + out_proj(...)` / `x = x + linear2(...)` followed by norm2 / norm3 / next-layer norm1 of
// nn.TransformerDecoderLayer(norm_first=True) (navdp.py L57-66), and the ls1/ls2-scaled residual adds of the DINOv2
// block (dinov2_layers/block.py L82-107).  Same warp-specialised skeleton as gemm_tcgen05.cu (TMA 128-B swizzle ring,
// 2-CTA W multicast), one accumulator stage.
#include <mutex>

#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {
namespace {

constexpr int N = 384, BM = 128, BK = 64;
constexpr int kStages = 3;
constexpr int kABytes = BM * BK * 2;   // 16 KB
constexpr int kBBytes = N * BK * 2;    // 48 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kStoreBytes = 8 * 2 * 2048;
constexpr int kThreads = 64 + 32 * 8;
constexpr int kSmem = kStages * kStageBytes + 256 + kStoreBytes + 1024;

struct RowArgs {
  int M, K, tiles_m;
  const float* bias;
  const float* gamma;
  const bf16* residual;
  int ldr;
  const float* ln_w;
  const float* ln_b;
  float ln_eps;
  int has_ln;
};

template <int CM>
__global__ void __launch_bounds__(kThreads, 1)
gemm_row384_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmLn,
                   const RowArgs args) {
  constexpr uint16_t kMask = (1u << CM) - 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tfull = bars + 2 * kStages;
  uint64_t* tempty = tfull + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 1);
  uint8_t* sStore = smem + kStages * kStageBytes + 256;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CM > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CM, num_clusters = gridDim.x / CM;
  const int super_m = (args.tiles_m + CM - 1) / CM;
  const int nkb = (args.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA), tma_prefetch_desc(&tmB), tma_prefetch_desc(&tmX);
    if (args.has_ln) tma_prefetch_desc(&tmLn);
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1), mbar_init(&empty[s], CM);
    mbar_init(tfull, 1), mbar_init(tempty, 8);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CM > 1) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = cluster_id; t < super_m; t += num_clusters) {
        const int tm = t * CM + rank;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], kStageBytes);
          tma_load_2d(sA + stage * kABytes, &tmA, &full[stage], kb * BK, tm * BM);
          uint8_t* dB = sB + stage * kBBytes;
          if (CM == 1) {  // box = 192 rows: two loads
            tma_load_2d(dB, &tmB, &full[stage], kb * BK, 0);
            tma_load_2d(dB + 192 * 128, &tmB, &full[stage], kb * BK, 192);
          } else {
            tma_load_2d_mc(dB + rank * 192 * 128, &tmB, &full[stage], kb * BK, rank * 192, kMask);
          }
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, 128);
      int stage = 0;
      uint32_t phase = 0, tphase = 0;
      for (int t = cluster_id; t < super_m; t += num_clusters) {
        mbar_wait(tempty, tphase ^ 1);
        tc_fence_after();
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t ad = umma_desc_sw128(smem_u32(sA + stage * kABytes));
          const uint64_t bd = umma_desc_sw128(smem_u32(sB + stage * kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k)
#pragma unroll
            for (int n = 0; n < 3; ++n)  // W rows n*128.. are 16 KB (= 1024 sixteen-byte units) further
              umma_f16(tmem + n * 128, ad + 2 * k, bd + n * 1024 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          if (CM == 1) umma_commit(&empty[stage]); else umma_commit_mc(&empty[stage], kMask);
          if (++stage == kStages) stage = 0, phase ^= 1;
        }
        umma_commit(tfull);
        tphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue: two warps share each 32-row quarter
    // (each owns 192 of the 384 output columns; both compute the full-row LayerNorm statistics for themselves)
    const int quarter = warp & 3, half = (warp - 2) >> 2;  // half h owns columns [192 h, 192 h + 192)
    const int r_in_tile = quarter * 32 + lane;
    uint8_t* my_store = sStore + (warp - 2) * 4096;
    int sbuf = 0;
    uint32_t tphase = 0;
    for (int t = cluster_id; t < super_m; t += num_clusters) {
      const int tm = t * CM + rank;
      const int row = tm * BM + r_in_tile;
      const bool row_ok = row < args.M;
      mbar_wait(tfull, tphase);
      tphase ^= 1;
      tc_fence_after();
      float mean = 0.f, rstd = 0.f;
      // pass 0: row statistics (only with LN); pass 1: produce and store the outputs
      for (int pass = args.has_ln ? 0 : 1; pass < 2; ++pass) {
        float sum = 0.f, sq = 0.f;
        // pass 0 walks the whole row (both column halves) so that no cross-warp exchange is needed; pass 1 only this
        // warp's 192 columns
        const int c_begin = pass == 0 ? 0 : half * 6, c_end = pass == 0 ? 12 : half * 6 + 6;
#pragma unroll 1
        for (int c = c_begin; c < c_end; ++c) {
          const int col0 = c * 32;
          uint32_t r[32];
          __syncwarp();
          tmem_ld32(tmem + (uint32_t(quarter * 32) << 16) + col0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
            v[j] = __uint_as_float(r[j]) + b.x, v[j + 1] = __uint_as_float(r[j + 1]) + b.y;
            v[j + 2] = __uint_as_float(r[j + 2]) + b.z, v[j + 3] = __uint_as_float(r[j + 3]) + b.w;
          }
          if (args.gamma) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(args.gamma + col0 + j));
              v[j] *= g.x, v[j + 1] *= g.y, v[j + 2] *= g.z, v[j + 3] *= g.w;
            }
          }
          if (row_ok && args.residual) {
            const bf16* rr = args.residual + (long)row * args.ldr + col0;
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              const uint4 q = *reinterpret_cast<const uint4*>(rr + j);
              v[j + 0] += bf16_lo(q.x), v[j + 1] += bf16_hi(q.x), v[j + 2] += bf16_lo(q.y), v[j + 3] += bf16_hi(q.y);
              v[j + 4] += bf16_lo(q.z), v[j + 5] += bf16_hi(q.z), v[j + 6] += bf16_lo(q.w), v[j + 7] += bf16_hi(q.w);
            }
          }
          // the residual stream is stored in bf16: the LayerNorm sees exactly those rounded values, like the reference
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __bfloat162float(__float2bfloat16(v[j]));
          if (pass == 0) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += v[j], sq += v[j] * v[j];
            continue;
          }
          // ---- x_new tile
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sb = my_store + sbuf * 2048;
#pragma unroll
          for (int j = 0; j < 32; j += 8)
            *reinterpret_cast<uint4*>(sb + lane * 64 + j * 2) =
                make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                           pack_bf16(v[j + 6], v[j + 7]));
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmX, sb, col0, tm * BM + quarter * 32);
            tma_store_commit();
          }
          sbuf ^= 1;
          if (args.has_ln) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 w = __ldg(reinterpret_cast<const float4*>(args.ln_w + col0 + j));
              const float4 b = __ldg(reinterpret_cast<const float4*>(args.ln_b + col0 + j));
              v[j] = (v[j] - mean) * rstd * w.x + b.x, v[j + 1] = (v[j + 1] - mean) * rstd * w.y + b.y;
              v[j + 2] = (v[j + 2] - mean) * rstd * w.z + b.z, v[j + 3] = (v[j + 3] - mean) * rstd * w.w + b.w;
            }
            if (lane == 0) tma_store_wait_read<1>();
            __syncwarp();
            uint8_t* sl = my_store + sbuf * 2048;
#pragma unroll
            for (int j = 0; j < 32; j += 8)
              *reinterpret_cast<uint4*>(sl + lane * 64 + j * 2) =
                  make_uint4(pack_bf16(v[j], v[j + 1]), pack_bf16(v[j + 2], v[j + 3]), pack_bf16(v[j + 4], v[j + 5]),
                             pack_bf16(v[j + 6], v[j + 7]));
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              tma_store_2d(&tmLn, sl, col0, tm * BM + quarter * 32);
              tma_store_commit();
            }
            sbuf ^= 1;
          }
        }
        if (pass == 0) {
          mean = sum * (1.0f / N);
          const float var = fmaxf(sq * (1.0f / N) - mean * mean, 0.f);
          rstd = rsqrtf(var + args.ln_eps);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty);
    }
    if (lane == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  if (CM > 1) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

template <int CM>
void launch_row(const bf16* A, int lda, const bf16* W, int ldw, const RowArgs& a, bf16* out, int ldo, bf16* ln_out,
                int ld_ln, cudaStream_t stream) {
  static std::once_flag once;
  static int max_clusters = 0;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_row384_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    max_clusters = device_sm_count() / CM;
    if (CM > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(device_sm_count() / CM * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
      cfg.attrs = &at, cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_row384_kernel<CM>, &cfg) == cudaSuccess && n > 0) max_clusters = n;
    }
  });
  CUtensorMap tmA = tma_map_2d(A, a.M, a.K, lda, BM, BK, true);
  CUtensorMap tmB = tma_map_2d(W, N, a.K, ldw, 192, BK, true);
  CUtensorMap tmX = tma_map_2d(out, a.M, N, ldo, 32, 32, false);
  CUtensorMap tmLn = ln_out ? tma_map_2d(ln_out, a.M, N, ld_ln, 32, 32, false) : tmX;
  const int super_m = (a.tiles_m + CM - 1) / CM;
  const int clusters = super_m < max_clusters ? super_m : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem, cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
  cfg.attrs = &at, cfg.numAttrs = 1;
  N1_CUDA(cudaLaunchKernelEx(&cfg, gemm_row384_kernel<CM>, tmA, tmB, tmX, tmLn, a));
  prof_count_gemm(2.0 * a.M * (double)N * a.K);
  N1_CUDA(cudaGetLastError());
}

}  // namespace

void gemm_row384(const bf16* A, int lda, const bf16* W, int ldw, int M, int K, const float* bias, const float* gamma,
                 const bf16* residual, int ldr, bf16* out, int ldo, const float* ln_w, const float* ln_b, float ln_eps,
                 bf16* ln_out, int ld_ln, cudaStream_t stream) {
  if (M <= 0) return;
  N1_CHECK(K > 0 && K % 8 == 0 && bias != nullptr, "gemm_row384: K % 8 and a bias are required");
  N1_CHECK((ln_out == nullptr) == (ln_w == nullptr), "gemm_row384: ln_out and ln_w go together");
  N1_CHECK(ln_out == nullptr || ln_b != nullptr, "gemm_row384: LayerNorm needs a bias vector");
  // with the fused LayerNorm every epilogue warp reads the WHOLE residual row for the statistics while its partner may
  // already be storing the new row: the output must not alias the residual
  N1_CHECK(ln_out == nullptr || static_cast<const void*>(out) != static_cast<const void*>(residual),
           "gemm_row384: in-place residual update is not allowed together with the fused LayerNorm");
  RowArgs a;
  a.M = M, a.K = K, a.tiles_m = (M + BM - 1) / BM;
  a.bias = bias, a.gamma = gamma, a.residual = residual, a.ldr = ldr;
  a.ln_w = ln_w, a.ln_b = ln_b, a.ln_eps = ln_eps, a.has_ln = ln_out ? 1 : 0;
  if (a.tiles_m >= 2)
    launch_row<2>(A, lda, W, ldw, a, out, ldo, ln_out, ld_ln, stream);
  else
    launch_row<1>(A, lda, W, ldw, a, out, ldo, ln_out, ld_ln, stream);
}

}  // namespace n1
