// Persistent warp-specialised bf16 GEMM for sm_100a:  out = epilogue(A[M,K] @ W[N,K]^T).
//
//   warp 0        TMA producer   (cp.async.bulk.tensor 2-D, 128-byte swizzle, STAGES-deep mbarrier ring)
//   warp 1        TMEM allocator + tcgen05.mma issuer (one elected lane; M=128, N=BN, K=16 per instruction)
//   warps 2..9    epilogue: tcgen05.ld (TMEM -> registers), bias / activation / layer-scale / residual,
//                 bf16 (or fp32) vector stores.  Two accumulator stages in TMEM so the epilogue of tile i
//                 overlaps the main loop of tile i+1.
//
// This one kernel carries every Linear / Conv-as-GEMM on the InternVLA-N1 hot path:
// the reference reaches cuBLAS through nn.Linear at navdp.py L57-66/L94-100, navdp_backbone.py L147-149,
// dinov2_layers/{attention.py L46-48, mlp.py L30-32, patch_embed.py L65} and the Qwen2.5-VL blocks.
#include <stdlib.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>

#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;

template <int BN>
struct Cfg {
  static constexpr int kABytes = BM * BK * 2;
  static constexpr int kBBytes = BN * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BN >= 256) ? 4 : (BN >= 128 ? 6 : 8);
  static constexpr int kTmemCols = (2 * BN < 32) ? 32 : 2 * BN;  // power of two for BN in {64,128,256}
  static constexpr int kBarBytes = 256;
  static constexpr int kStoreBytes = kEpiWarps * 2 * 2048;  // per-warp double-buffered 32x32 bf16 staging for TMA stores
  static constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + kStoreBytes + 1024;  // +1024: manual alignment
};

struct GemmArgs {
  int M, N, K;
  int tiles_m, tiles_n;
  void* out;
  int ldo;
  const float* bias;
  const float* gamma;
  const bf16* residual;
  int ldr;
  int act;
  int out_fp32;
  int rows_per_group, group_stride, group_offset;
  const float* row_add;
  int raster_g;   // M (super-)tiles per raster group (decode_tile)
  int tma_store;  // bf16 output tile goes registers -> smem staging -> cp.async.bulk.tensor store (full-sector writes)
};

// Grouped rasterisation: consecutive tile ids walk G M-tiles before moving to the next N-tile, so the CTAs resident at
// one time share few W panels, and the A panels of a group stay L2-resident while the group sweeps all N tiles.  DRAM
// traffic ~ W bytes x (tiles_m / G) + A bytes: G = 8 re-read W 19 times for the LLM gate/up GEMM (N = 37888: 5.3 GB per
// launch for 0.41 GB algorithmic, profiles/r1_ncu_full_gemm_llm_v0_summary.txt), so G now grows until the group's A
// panels fill about 28 MB (measured optimum: 64 MB groups are not L2-resident next to the streaming W -- the two L2
// partitions duplicate lines -- and 16 MB re-reads W more often; profiles/README.md).
__device__ __forceinline__ void decode_tile(int tile, int tiles_m, int tiles_n, int G, int& tm, int& tn) {
  const int per_group = G * tiles_n;
  const int g = tile / per_group;
  const int first_m = g * G;
  const int gsize = min(G, tiles_m - first_m);
  const int r = tile - g * per_group;
  tm = first_m + r % gsize;
  tn = r / gsize;
}

// CM = cluster size along M.  With CM > 1 the CM CTAs of a cluster work on CM vertically adjacent tiles of the same
// N panel: every CTA fetches 1/CM of the W tile and multicasts it to its peers, which divides the L2 -> SMEM traffic
// for W by CM (the short-K GEMMs of the path are L2-bandwidth-bound, profiles/r1_ncu_small_v0_summary.txt).
// UI ("uniform issue"): the MMA warp runs its loop with all 32 lanes and an elected lane issues, so the descriptors stay
// in uniform registers instead of being moved there (R2UR + ELECT + R2UR.BROADCAST) before every tcgen05.mma.
template <int BN, int CM, bool UI>
__global__ void __launch_bounds__(kThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const GemmArgs args) {
  using C = Cfg<BN>;
  constexpr uint16_t kMask = (1u << CM) - 1;
  const int rank = CM > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CM, num_clusters = gridDim.x / CM;
  const int super_m = (args.tiles_m + CM - 1) / CM;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + C::kStages;
  uint64_t* tfull = bars + 2 * C::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  uint8_t* sStore = smem + C::kStages * C::kStageBytes + C::kBarBytes;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_tiles = super_m * args.tiles_n;  // cluster-level (super) tiles: CM m-tiles x 1 n-tile
  const int nkb = (args.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (args.tma_store) tma_prefetch_desc(&tmC);
    for (int s = 0; s < C::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], CM);  // every CTA of the cluster releases the slot (its W slice lives in all of them)
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tfull[s], 1);
      mbar_init(&tempty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, C::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  if (CM > 1)
    cluster_sync_all();  // peers' barriers must be initialised before any multicast / remote arrive reaches them
  else
    __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        int tm, tn;
        decode_tile(tile, super_m, args.tiles_n, args.raster_g, tm, tn);
        tm = tm * CM + rank;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full[stage], C::kStageBytes);
          tma_load_2d(sA + stage * C::kABytes, &tmA, &full[stage], kb * BK, tm * BM);
          if (CM == 1) {
            tma_load_2d(sB + stage * C::kBBytes, &tmB, &full[stage], kb * BK, tn * BN);
          } else {
            constexpr int kSliceRows = BN / CM;
            tma_load_2d_mc(sB + stage * C::kBBytes + rank * kSliceRows * BK * 2, &tmB, &full[stage], kb * BK,
                           tn * BN + rank * kSliceRows, kMask);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (UI || lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
        mbar_wait(&tempty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = umma_desc_sw128(smem_u32(sA + stage * C::kABytes));
          const uint64_t bdesc = umma_desc_sw128(smem_u32(sB + stage * C::kBBytes));
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // advance 16 elements (32 bytes) along K inside the 128-byte swizzle row: +2 in 16-byte units
            if (UI) umma_f16_elect(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else umma_f16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if (CM == 1) {  // smem slot reusable once these MMAs have read it
            if (UI) umma_commit_elect(&empty[stage]); else umma_commit(&empty[stage]);
          } else {
            if (UI) umma_commit_mc_elect(&empty[stage], kMask); else umma_commit_mc(&empty[stage], kMask);
          }
          if (++stage == C::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (UI) umma_commit_elect(&tfull[acc]); else umma_commit(&tfull[acc]);  // accumulator complete
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue
    const int quarter = warp & 3;           // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;       // two warps per quarter split the column chunks
    int acc = 0;
    uint32_t acc_phase = 0;
    const bool swiglu = args.act == ACT_SWIGLU;
    uint8_t* my_store = sStore + (warp - 2) * 4096;
    int sbuf = 0;
    for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
      int tm, tn;
      decode_tile(tile, super_m, args.tiles_n, args.raster_g, tm, tn);
      tm = tm * CM + rank;
      const int row = tm * BM + quarter * 32 + lane;
      const bool row_ok = row < args.M;
      long out_row = row;
      int grp_row = 0;
      if (args.rows_per_group > 0) {
        grp_row = row % args.rows_per_group;
        out_row = (long)(row / args.rows_per_group) * args.group_stride + grp_row + args.group_offset;
      }
      // Residual tile: per-row 16-byte loads whose latency dominated the epilogue of the short-K GEMMs
      // (profiles/r1_ncu_gemm384_v3_summary.txt).  They do not depend on the accumulator, so the loads of chunk c + 2
      // are issued while chunk c is processed, and those of the first chunk before the accumulator wait.
      const bool use_res = args.residual != nullptr && row_ok && !swiglu;
      uint4 res_next[4];
      auto prefetch_res = [&](int chunk) {
        const int c0 = tn * BN + chunk * 32;
        const bf16* rr = args.residual + out_row * args.ldr + c0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          res_next[j] = (use_res && c0 + j * 8 < args.N) ? __ldg(reinterpret_cast<const uint4*>(rr + j * 8))
                                                         : make_uint4(0u, 0u, 0u, 0u);
      };
      prefetch_res(half);
      mbar_wait(&tfull[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int chunk = half; chunk < BN / 32; chunk += 2) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem_base + (uint32_t(quarter * 32) << 16) + acc * BN + chunk * 32, r);
        const uint4 res_cur[4] = {res_next[0], res_next[1], res_next[2], res_next[3]};
        if (chunk + 2 < BN / 32) prefetch_res(chunk + 2);
        tmem_ld_wait();
        const int col0 = tn * BN + chunk * 32;
        if (col0 >= args.N) continue;  // warp-uniform
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (args.bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < args.N) {
              const float4 b = __ldg(reinterpret_cast<const float4*>(args.bias + col0 + j));
              v[j] += b.x, v[j + 1] += b.y, v[j + 2] += b.z, v[j + 3] += b.w;
            }
          }
        }
        if (swiglu && args.tma_store) {
          // (gate, up) interleaved: 32 accumulator columns -> 16 outputs, staged as a 32 x 16 tile
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sb = my_store + sbuf * 2048;
#pragma unroll
          for (int j = 0; j < 16; j += 8) {
            uint4 pk;
            pk.x = pack_bf16(silu(v[2 * j + 0]) * v[2 * j + 1], silu(v[2 * j + 2]) * v[2 * j + 3]);
            pk.y = pack_bf16(silu(v[2 * j + 4]) * v[2 * j + 5], silu(v[2 * j + 6]) * v[2 * j + 7]);
            pk.z = pack_bf16(silu(v[2 * j + 8]) * v[2 * j + 9], silu(v[2 * j + 10]) * v[2 * j + 11]);
            pk.w = pack_bf16(silu(v[2 * j + 12]) * v[2 * j + 13], silu(v[2 * j + 14]) * v[2 * j + 15]);
            *reinterpret_cast<uint4*>(sb + lane * 32 + j * 2) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, sb, col0 >> 1, tm * BM + quarter * 32);
            tma_store_commit();
          }
          sbuf ^= 1;
          continue;
        }
        if (swiglu) {
          // (gate, up) interleaved: 32 accumulator columns -> 16 outputs
          if (row_ok) {
            const int ocol0 = col0 >> 1;
            bf16* orow = reinterpret_cast<bf16*>(args.out) + out_row * args.ldo + ocol0;
#pragma unroll
            for (int j = 0; j < 16; j += 8) {
              if (ocol0 + j < (args.N >> 1)) {
                uint4 pk;
                pk.x = pack_bf16(silu(v[2 * j + 0]) * v[2 * j + 1], silu(v[2 * j + 2]) * v[2 * j + 3]);
                pk.y = pack_bf16(silu(v[2 * j + 4]) * v[2 * j + 5], silu(v[2 * j + 6]) * v[2 * j + 7]);
                pk.z = pack_bf16(silu(v[2 * j + 8]) * v[2 * j + 9], silu(v[2 * j + 10]) * v[2 * j + 11]);
                pk.w = pack_bf16(silu(v[2 * j + 12]) * v[2 * j + 13], silu(v[2 * j + 14]) * v[2 * j + 15]);
                *reinterpret_cast<uint4*>(orow + j) = pk;
              }
            }
          }
          continue;
        }
        if (args.act == ACT_GELU) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float g[8] = {v[j], v[j + 1], v[j + 2], v[j + 3], v[j + 4], v[j + 5], v[j + 6], v[j + 7]};
            gelu_erf8(g);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[j + i] = g[i];
          }
        } else if (args.act == ACT_RELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.0f);
        } else if (args.act == ACT_GELU_TANH) {  // nn.GELU(approximate="tanh"): the NextDiT condition projections
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = 0.5f * v[j] * (1.0f + tanhf(0.7978845608028654f * (v[j] + 0.044715f * v[j] * v[j] * v[j])));
        } else if (args.act == ACT_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = silu(v[j]);
        }
        if (args.gamma) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < args.N) {
              const float4 g = __ldg(reinterpret_cast<const float4*>(args.gamma + col0 + j));
              v[j] *= g.x, v[j + 1] *= g.y, v[j + 2] *= g.z, v[j + 3] *= g.w;
            }
          }
        }
        if (row_ok) {
        if (args.row_add) {
          const float* ra = args.row_add + (long)grp_row * args.N + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col0 + j < args.N) {
              const float4 a = __ldg(reinterpret_cast<const float4*>(ra + j));
              v[j] += a.x, v[j + 1] += a.y, v[j + 2] += a.z, v[j + 3] += a.w;
            }
          }
        }
        if (args.residual) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < args.N) {
              const uint4 q = res_cur[j >> 3];
              v[j + 0] += bf16_lo(q.x), v[j + 1] += bf16_hi(q.x);
              v[j + 2] += bf16_lo(q.y), v[j + 3] += bf16_hi(q.y);
              v[j + 4] += bf16_lo(q.z), v[j + 5] += bf16_hi(q.z);
              v[j + 6] += bf16_lo(q.w), v[j + 7] += bf16_hi(q.w);
            }
          }
        }
        }  // row_ok (loads)
        if (args.tma_store) {
          if (lane == 0) tma_store_wait_read<1>();
          __syncwarp();
          uint8_t* sb = my_store + sbuf * 2048;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            uint4 pk;
            pk.x = pack_bf16(v[j + 0], v[j + 1]);
            pk.y = pack_bf16(v[j + 2], v[j + 3]);
            pk.z = pack_bf16(v[j + 4], v[j + 5]);
            pk.w = pack_bf16(v[j + 6], v[j + 7]);
            *reinterpret_cast<uint4*>(sb + lane * 64 + j * 2) = pk;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&tmC, sb, col0, tm * BM + quarter * 32);  // rows >= M / cols >= N are clipped by the tensor map
            tma_store_commit();
          }
          sbuf ^= 1;
          continue;
        }
        if (row_ok) {
        if (args.out_fp32) {
          float* orow = reinterpret_cast<float*>(args.out) + out_row * args.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            if (col0 + j < args.N)
              *reinterpret_cast<float4*>(orow + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        } else {
          bf16* orow = reinterpret_cast<bf16*>(args.out) + out_row * args.ldo + col0;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col0 + j < args.N) {
              uint4 pk;
              pk.x = pack_bf16(v[j + 0], v[j + 1]);
              pk.y = pack_bf16(v[j + 2], v[j + 3]);
              pk.z = pack_bf16(v[j + 4], v[j + 5]);
              pk.w = pack_bf16(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(orow + j) = pk;
            }
          }
        }
        }  // row_ok
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
    if (args.tma_store && lane == 0) tma_store_wait<0>();
    __syncwarp();
  }

  tc_fence_before();
  if (CM > 1)
    cluster_sync_all();  // no CTA may retire while a peer can still multicast into it or arrive on its barriers
  else
    __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) throw Error(-4, "cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU?)");
  return fn;
}

// 2-D bf16 tensor map over a row-major [rows, cols] matrix with leading dimension ld; box = [box_rows, box_cols];
// operands use box_cols = 64 with the 128-byte swizzle, output staging tiles are unswizzled.
CUtensorMap make_map(const bf16* ptr, long rows, long cols, long ld, int box_rows, int box_cols = BK, bool swizzle = true) {
  N1_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "GEMM operand not 16-byte aligned");
  N1_CHECK(ld % 8 == 0, "GEMM operand leading dimension must be a multiple of 8 elements");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstr, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-5, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return m;
}

template <int BN, int CM, bool UI>
void launch(const bf16* A, int lda, const bf16* W, int ldw, int M, int N, int K, GemmArgs& a, cudaStream_t stream) {
  using C = Cfg<BN>;
  static std::once_flag once;
  static int max_clusters = 0;
  std::call_once(once, [] {
    cudaFuncSetAttribute(gemm_kernel<BN, CM, UI>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    max_clusters = device_sm_count() / CM;
    if (CM > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(device_sm_count() / CM * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = C::kSmemBytes;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
      cfg.attrs = &at, cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_kernel<BN, CM, UI>, &cfg) == cudaSuccess && n > 0) max_clusters = n;
    }
  });
  a.tiles_m = (M + BM - 1) / BM;
  a.tiles_n = (N + BN - 1) / BN;
  {
    // raster group: as many (super-)tiles of A as fit ~28 MB of L2; only matters when W does not fit L2
    const long panel = (long)BM * CM * K * 2;
    const long w_bytes = (long)N * K * 2;
    int g = 8;
    static long budget_mb = -1;  // N1_GEMM_RASTER_MB: L2 budget for the A panels of a raster group
    if (budget_mb < 0) {
      const char* e = getenv("N1_GEMM_RASTER_MB");
      budget_mb = e ? atol(e) : 28;
    }
    // (L2 eviction-priority hints on the TMA loads -- W evict-first, A evict-last -- were measured and made it worse: the
    // concurrent clusters that share a W tile lose it before they read it; profiles/README.md, raster experiments)
    if (w_bytes > (64L << 20)) g = (int)std::max<long>(2, (budget_mb << 20) / panel);
    const int super_m = (a.tiles_m + CM - 1) / CM;
    a.raster_g = g < super_m ? g : super_m;
    if (a.raster_g < 1) a.raster_g = 1;
  }
  CUtensorMap tmA = make_map(A, M, K, lda, BM);
  CUtensorMap tmB = make_map(W, N, K, ldw, BN / CM);
  CUtensorMap tmC = tmA;  // placeholder when the direct-store epilogue is used
  if (a.tma_store) {
    const bool sw = a.act == ACT_SWIGLU;
    tmC = make_map(static_cast<const bf16*>(a.out), M, sw ? N / 2 : N, a.ldo, 32, sw ? 16 : 32, false);
  }
  const int super_tiles = ((a.tiles_m + CM - 1) / CM) * a.tiles_n;
  const int clusters = super_tiles < max_clusters ? super_tiles : max_clusters;
  if (CM == 1) {
    gemm_kernel<BN, CM, UI><<<clusters, kThreads, C::kSmemBytes, stream>>>(tmA, tmB, tmC, a);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(clusters * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = C::kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
    cfg.attrs = &at, cfg.numAttrs = 1;
    N1_CUDA(cudaLaunchKernelEx(&cfg, gemm_kernel<BN, CM, UI>, tmA, tmB, tmC, a));
  }
  N1_CUDA(cudaGetLastError());
}

std::atomic<int> g_cluster_m{2};
std::atomic<int> g_uniform_issue{-1};  // N1_GEMM_UI: 1 = warp-uniform MMA issue everywhere, 0 = nowhere, default (-1) = K >= 1024
std::atomic<int> g_tma_store{1};  // developer knob (N1_GEMM_TMA_STORE=0 selects the direct-store epilogue)  // developer knob (N1_GEMM_CLUSTER=1 disables the multicast path)

// ---- profiling state
std::atomic<long> g_total_launches{0}, g_gemm_launches{0};
std::atomic<bool> g_prof_on{false};
std::mutex g_prof_mu;
struct EvPair {
  cudaEvent_t a, b;
  double flops;
  int M, N, K;
};
std::vector<EvPair> g_events;
struct ShapeAcc {
  int M, N, K;
  long count;
  double ms;
};
std::vector<ShapeAcc> g_shapes;  // per-(M, N, K) sums of the event-timed launches since the last prof_read_shapes()

}  // namespace

CUtensorMap tma_map_2d(const bf16* ptr, long rows, long cols, long ld, int box_rows, int box_cols, bool swizzle) {
  return make_map(ptr, rows, cols, ld, box_rows, box_cols, swizzle);
}
// GEMM-class kernels in other files: counted like gemm_bf16 launches (event timing is only done for gemm_bf16)
void prof_count_gemm(double flops) {
  (void)flops;
  g_gemm_launches++;
  g_total_launches++;
}

void prof_enable(bool on) { g_prof_on = on; }
// Event bracket for GEMM-class kernels launched outside gemm_bf16 (the fused decoder blocks): returns a ticket (< 0: off)
int prof_begin(double flops, int M, int N, int K, cudaStream_t s) {
  if (!g_prof_on.load()) return -1;
  EvPair ev{};
  cudaEventCreate(&ev.a);
  cudaEventCreate(&ev.b);
  ev.flops = flops, ev.M = M, ev.N = N, ev.K = K;
  cudaEventRecord(ev.a, s);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  g_events.push_back(ev);
  return (int)g_events.size() - 1;
}
void prof_end(int ticket, cudaStream_t s) {
  if (ticket < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (ticket < (int)g_events.size()) cudaEventRecord(g_events[ticket].b, s);
}
void prof_count_launch(int n) { g_total_launches += n; }
ProfStats prof_read_and_reset() {
  ProfStats st;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  for (EvPair& e : g_events) {
    cudaEventSynchronize(e.b);
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, e.a, e.b) == cudaSuccess) {
      st.gemm_ms += ms, st.gemm_flops += e.flops;
      bool hit = false;
      for (ShapeAcc& sa : g_shapes)
        if (sa.M == e.M && sa.N == e.N && sa.K == e.K) {
          sa.count++, sa.ms += ms, hit = true;
          break;
        }
      if (!hit) g_shapes.push_back({e.M, e.N, e.K, 1, (double)ms});
    }
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  g_events.clear();
  st.gemm_launches = g_gemm_launches.exchange(0);
  st.total_launches = g_total_launches.exchange(0);
  return st;
}

int prof_read_shapes(int* mnk, long* count, double* ms, int cap) {
  std::lock_guard<std::mutex> lk(g_prof_mu);
  int n = 0;
  for (const ShapeAcc& sa : g_shapes) {
    if (n >= cap) break;
    mnk[3 * n] = sa.M, mnk[3 * n + 1] = sa.N, mnk[3 * n + 2] = sa.K;
    count[n] = sa.count, ms[n] = sa.ms;
    ++n;
  }
  g_shapes.clear();
  return n;
}

int device_sm_count() {
  static int sms = 0;
  if (!sms) {
    if (const char* e = getenv("N1_GEMM_CLUSTER")) g_cluster_m = atoi(e);
    if (const char* e = getenv("N1_GEMM_TMA_STORE")) g_tma_store = atoi(e);
    if (const char* e = getenv("N1_GEMM_UI")) g_uniform_issue = atoi(e);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) sms = 148;
  }
  return sms;
}

void gemm_bf16(const bf16* A, int lda, const bf16* W, int ldw, void* out, int ldo, int M, int N, int K,
               const GemmEpilogue& e, cudaStream_t stream) {
  if (M <= 0 || N <= 0) return;
  N1_CHECK(K > 0 && K % 8 == 0, "GEMM K must be a positive multiple of 8");
  N1_CHECK(N % 8 == 0, "GEMM N must be a multiple of 8 (pad the packed weight)");
  N1_CHECK(!(e.act == ACT_SWIGLU) || N % 16 == 0, "SwiGLU GEMM needs N % 16 == 0");
  GemmArgs a;
  a.M = M, a.N = N, a.K = K;
  a.out = out, a.ldo = ldo;
  a.bias = e.bias, a.gamma = e.gamma, a.residual = e.residual, a.ldr = e.ldr;
  a.act = e.act, a.out_fp32 = e.out_fp32;
  a.rows_per_group = e.rows_per_group, a.group_stride = e.group_stride, a.group_offset = e.group_offset;
  a.row_add = e.row_add;
  a.tma_store = (!e.out_fp32 && e.rows_per_group == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && ldo % 8 == 0 &&
                 g_tma_store.load() != 0)
                    ? 1
                    : 0;
  // Tile-width choice: fewest waves first, then the widest tile (fewer A re-reads, longer MMA bursts).
  const int sms = device_sm_count();
  const int tm = (M + BM - 1) / BM;
  auto cost = [&](int bn) {
    const long tiles = (long)tm * ((N + bn - 1) / bn);
    const long waves = (tiles + sms - 1) / sms;
    return waves * (bn + 48);  // per-tile time ~ BN plus a fixed prologue/epilogue share
  };
  EvPair ev{};
  const bool prof = g_prof_on.load();
  if (prof) {
    cudaEventCreate(&ev.a);
    cudaEventCreate(&ev.b);
    ev.flops = 2.0 * M * (double)N * K;
    ev.M = M, ev.N = N, ev.K = K;
    cudaEventRecord(ev.a, stream);
  }
  g_gemm_launches++;
  g_total_launches++;
  int bn = 256;
  long best = cost(256);
  if (cost(128) < best) best = cost(128), bn = 128;
  if (cost(64) < best) bn = 64;
  // clusters of 2 along M (W multicast) whenever there are at least two M tiles to pair up
  const bool pair = tm >= 2 && g_cluster_m.load() >= 2;
  // uniform issue pays where the K loop is long (decoder / vision-tower shapes: -1.5 .. -3 % per stage, A/B in
  // profiles/README.md); the 384-wide System-1 shapes were neutral to slightly slower, so they keep the one-lane loop
  const int ui_mode = g_uniform_issue.load();
  const bool ui = ui_mode == 1 || (ui_mode < 0 && K >= 1024);
#define N1_LAUNCH(BN_, CM_)                                                       \
  do {                                                                            \
    if (ui) launch<BN_, CM_, true>(A, lda, W, ldw, M, N, K, a, stream);           \
    else launch<BN_, CM_, false>(A, lda, W, ldw, M, N, K, a, stream);             \
  } while (0)
  if (bn == 256) {
    if (pair) N1_LAUNCH(256, 2); else N1_LAUNCH(256, 1);
  } else if (bn == 128) {
    if (pair) N1_LAUNCH(128, 2); else N1_LAUNCH(128, 1);
  } else {
    if (pair) N1_LAUNCH(64, 2); else N1_LAUNCH(64, 1);
  }
#undef N1_LAUNCH
  if (prof) {
    cudaEventRecord(ev.b, stream);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_events.push_back(ev);
  }
}

}  // namespace n1
