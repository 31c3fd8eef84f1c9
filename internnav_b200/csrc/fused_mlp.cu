// Fused transformer MLP for the NavDP decoder:  out = residual + W2 · GELU(W1 · x + b1) + b2      (D = 384, F = 1536)
//
// Replaces linear1 -> exact GELU -> linear2 -> residual add of nn.TransformerDecoderLayer (navdp.py L57-66; the FF
// block of `self.decoder`, called 16 x K times per System-1 step).  Unfused, the [rows, 1536] hidden makes a round trip
// through HBM and both GEMMs are L2->SMEM-bound (profiles/r1_ncu_small_v0_summary.txt); here it never leaves the SM:
//
//   per CTA: one 128-row tile.  x tile (A operand, 96 KB) is TMA-loaded once.  The hidden is produced 64 columns at a
//   time:  H_j = x · W1[64j:64j+64]^T  (tcgen05.mma M128 N64, accumulators double-buffered in TMEM columns 384..511)
//   -> 8 epilogue warps: tcgen05.ld, +b1, GELU, bf16, written to SMEM in the 128-byte-swizzled K-major operand layout
//   -> Y += H_j · W2[:, 64j:64j+64]^T  (2 x tcgen05.mma M128 N192 K16 per k-step into TMEM columns 0..383).
//   W1 / W2 slices stream through a 4-slot x 24 KB TMA ring; with CM = 2 the two CTAs of a cluster each fetch half of
//   every slice and multicast it.  Final epilogue: Y + b2 + residual -> bf16 -> SMEM staging -> TMA store.
#include <mutex>

#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {
namespace {

constexpr int D = 384, F = 1536, BM = 128, HC = 64;  // hidden chunk width
constexpr int NCH = F / HC;                          // 24 chunks
constexpr int kSlots = 4, kSlotBytes = 24576;
constexpr int kLag = 2;                              // GEMM2(j) is issued after GEMM1(j + kLag)
constexpr int kABytes = BM * D * 2;                  // 98304: 6 k-blocks of [128 x 64]
constexpr int kHBytes = BM * HC * 2;                 // 16384 per buffer
constexpr int kThreads = 64 + 32 * 8;
constexpr int kSmem = kABytes + 2 * kHBytes + kSlots * kSlotBytes + 512 + 1024;

struct MlpArgs {
  int M;
  int tiles_m;
  const float* b1;
  const float* b2;
  const bf16* residual;
  int ldr;
};

template <int CM>
__global__ void __launch_bounds__(kThreads, 1)
fused_mlp_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const MlpArgs args) {
  constexpr uint16_t kMask = (1u << CM) - 1;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sH = smem + kABytes;
  uint8_t* sW = sH + 2 * kHBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW + kSlots * kSlotBytes);
  uint64_t* w_full = bars;                 // [4]
  uint64_t* w_empty = bars + 4;            // [4]
  uint64_t* a_full = bars + 8;             // x tile landed
  uint64_t* a_empty = bars + 9;            // all GEMM1 of the tile issued and complete
  uint64_t* hacc_full = bars + 10;         // [2] GEMM1(j) complete -> TMEM H readable
  uint64_t* hacc_empty = bars + 12;        // [2] epilogue finished reading TMEM H
  uint64_t* hs_full = bars + 14;           // [2] SMEM H written
  uint64_t* hs_empty = bars + 16;          // [2] GEMM2 finished reading SMEM H
  uint64_t* y_full = bars + 18;
  uint64_t* y_empty = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CM > 1 ? (int)cluster_ctarank() : 0;
  const int cluster_id = blockIdx.x / CM, num_clusters = gridDim.x / CM;
  const int super_m = (args.tiles_m + CM - 1) / CM;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmX), tma_prefetch_desc(&tmW1), tma_prefetch_desc(&tmW2), tma_prefetch_desc(&tmOut);
    for (int s = 0; s < kSlots; ++s) mbar_init(&w_full[s], 1), mbar_init(&w_empty[s], CM);
    mbar_init(a_full, 1), mbar_init(a_empty, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&hacc_full[b], 1), mbar_init(&hacc_empty[b], 4);  // buffer b is served by epilogue group b (4 warps)
      mbar_init(&hs_full[b], 4), mbar_init(&hs_empty[b], 1);
    }
    mbar_init(y_full, 1), mbar_init(y_empty, 8);
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
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int slot = 0;
      uint32_t wphase = 0, tphase = 0;
      auto next = [&]() { if (++slot == kSlots) slot = 0, wphase ^= 1; };
      for (int t = cluster_id; t < super_m; t += num_clusters) {
        const int tm = t * CM + rank;
        mbar_wait(a_empty, tphase ^ 1);
        mbar_arrive_expect_tx(a_full, kABytes);
        for (int kb = 0; kb < D / 64; ++kb) tma_load_2d(sA + kb * 16384, &tmX, a_full, kb * 64, tm * BM);
        for (int j = 0; j < NCH + kLag; ++j) {
          if (j < NCH) {  // W1 rows [64j, 64j+64): two half-slices of 3 k-blocks
            for (int hf = 0; hf < 2; ++hf) {
              mbar_wait(&w_empty[slot], wphase ^ 1);
              mbar_arrive_expect_tx(&w_full[slot], kSlotBytes);
              uint8_t* dst = sW + slot * kSlotBytes;
              for (int kb = 0; kb < 3; ++kb) {
                if (CM == 1) {
                  tma_load_2d(dst + kb * 8192, &tmW1, &w_full[slot], (hf * 3 + kb) * 64, j * HC);
                } else {  // each CTA fetches 32 of the 64 rows and multicasts them
                  tma_load_2d_mc(dst + kb * 8192 + rank * 4096, &tmW1, &w_full[slot], (hf * 3 + kb) * 64,
                                 j * HC + rank * 32, kMask);
                }
              }
              next();
            }
          }
          if (j >= kLag) {  // W2[:, 64(j-kLag) : +64): two N-halves of 192 rows
            for (int nh = 0; nh < 2; ++nh) {
              mbar_wait(&w_empty[slot], wphase ^ 1);
              mbar_arrive_expect_tx(&w_full[slot], kSlotBytes);
              uint8_t* dst = sW + slot * kSlotBytes;
              if (CM == 1) {
                tma_load_2d(dst, &tmW2, &w_full[slot], (j - kLag) * HC, nh * 192);
              } else {
                tma_load_2d_mc(dst + rank * 12288, &tmW2, &w_full[slot], (j - kLag) * HC, nh * 192 + rank * 96, kMask);
              }
              next();
            }
          }
        }
        tphase ^= 1;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc1 = umma_idesc_bf16(BM, HC);
      constexpr uint32_t idesc2 = umma_idesc_bf16(BM, 192);
      int slot = 0;
      uint32_t wphase = 0, tphase = 0;
      uint32_t hacc_ph[2] = {0, 0}, hs_ph[2] = {0, 0};
      auto next = [&]() { if (++slot == kSlots) slot = 0, wphase ^= 1; };
      auto release = [&](uint64_t* bar) { if (CM == 1) umma_commit(bar); else umma_commit_mc(bar, kMask); };
      for (int t = cluster_id; t < super_m; t += num_clusters) {
        mbar_wait(a_full, tphase);
        tc_fence_after();
        for (int j = 0; j < NCH + kLag; ++j) {
          if (j < NCH) {
            const int b = j & 1;
            mbar_wait(&hacc_empty[b], hacc_ph[b] ^ 1);  // epilogue done with the previous use of TMEM H[b]
            tc_fence_after();
            for (int hf = 0; hf < 2; ++hf) {
              mbar_wait(&w_full[slot], wphase);
              tc_fence_after();
#pragma unroll
              for (int kb = 0; kb < 3; ++kb) {
                const uint64_t ad = umma_desc_sw128(smem_u32(sA + (hf * 3 + kb) * 16384));
                const uint64_t bd = umma_desc_sw128(smem_u32(sW + slot * kSlotBytes + kb * 8192));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem + 384 + b * HC, ad + 2 * k, bd + 2 * k, idesc1, (hf | kb | k) != 0 ? 1u : 0u);
              }
              release(&w_empty[slot]);
              next();
            }
            umma_commit(&hacc_full[b]);
            hacc_ph[b] ^= 1;
            if (j == NCH - 1) umma_commit(a_empty);  // x tile no longer needed once these MMAs retire
          }
          if (j >= kLag) {  // GEMM2 trails GEMM1 by kLag chunks so the tensor pipe has work while GELU(H) is produced
            const int jj = j - kLag, b = jj & 1;
            mbar_wait(&hs_full[b], hs_ph[b]);  // GELU(H_jj) is in shared memory
            hs_ph[b] ^= 1;
            if (jj == 0) mbar_wait(y_empty, tphase ^ 1);  // previous tile's Y has been read out
            tc_fence_after();
            const uint64_t ad = umma_desc_sw128(smem_u32(sH + b * kHBytes));
            for (int nh = 0; nh < 2; ++nh) {
              mbar_wait(&w_full[slot], wphase);
              tc_fence_after();
              const uint64_t bd = umma_desc_sw128(smem_u32(sW + slot * kSlotBytes));
#pragma unroll
              for (int k = 0; k < 4; ++k)
                umma_f16(tmem + nh * 192, ad + 2 * k, bd + 2 * k, idesc2, (jj | k) != 0 ? 1u : 0u);
              release(&w_empty[slot]);
              next();
            }
            umma_commit(&hs_empty[b]);
          }
        }
        umma_commit(y_full);
        tphase ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3, half = (warp - 2) >> 2;  // `half` doubles as the epilogue group id (= H buffer it serves)
    const int r_in_tile = quarter * 32 + lane;
    uint32_t hacc_phase = 0, hs_phase = 0, tphase = 0;
    // Output staging reuses the H buffers once the tile's GEMM2s are done.  Each warp stages inside the 4 KB slice of
    // H[half] that holds ITS OWN rows (32 rows x 128 B), so the next tile's GELU writes -- which only it performs, after
    // its own bulk-store reads have drained -- can never overwrite another warp's pending store.
    uint8_t* my_store = sH + half * kHBytes + quarter * 4096;
    for (int t = cluster_id; t < super_m; t += num_clusters) {
      const int tm = t * CM + rank;
      // group g (4 warps = 128 rows) owns the chunks j = g, g + 2, ...: two chunks are in flight, so the TMEM-load /
      // barrier latencies of one overlap the GELU arithmetic of the other
      for (int j = half; j < NCH; j += 2) {
        const int b = half;
        mbar_wait(&hacc_full[b], hacc_phase);
        hacc_phase ^= 1;
        tc_fence_after();
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem + (uint32_t(quarter * 32) << 16) + 384 + b * HC, r0);
        tmem_ld32(tmem + (uint32_t(quarter * 32) << 16) + 384 + b * HC + 32, r1);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hacc_empty[b]);
        const float* bias = args.b1 + j * HC;
        uint32_t pk[32];
#pragma unroll
        for (int c = 0; c < 64; c += 8) {
          const uint32_t* rr = c < 32 ? r0 + c : r1 + (c - 32);
          const float4 ba = __ldg(reinterpret_cast<const float4*>(bias + c));
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c + 4));
          float g[8] = {__uint_as_float(rr[0]) + ba.x, __uint_as_float(rr[1]) + ba.y, __uint_as_float(rr[2]) + ba.z,
                        __uint_as_float(rr[3]) + ba.w, __uint_as_float(rr[4]) + bb.x, __uint_as_float(rr[5]) + bb.y,
                        __uint_as_float(rr[6]) + bb.z, __uint_as_float(rr[7]) + bb.w};
          gelu_erf8(g);
          pk[c / 2] = pack_bf16(g[0], g[1]), pk[c / 2 + 1] = pack_bf16(g[2], g[3]);
          pk[c / 2 + 2] = pack_bf16(g[4], g[5]), pk[c / 2 + 3] = pack_bf16(g[6], g[7]);
        }
        mbar_wait(&hs_empty[b], hs_phase ^ 1);  // GEMM2(j-2) has finished reading this buffer
        hs_phase ^= 1;
        // K-major, 128-byte swizzle: 16-byte chunk c of row r lives at chunk (c ^ (r % 8)) of its 128-byte row
        uint8_t* row = sH + b * kHBytes + (r_in_tile >> 3) * 1024 + (r_in_tile & 7) * 128;
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<uint4*>(row + ((c ^ (r_in_tile & 7)) << 4)) =
              make_uint4(pk[c * 4], pk[c * 4 + 1], pk[c * 4 + 2], pk[c * 4 + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) mbar_arrive(&hs_full[b]);
      }
      // ---- final epilogue of the tile: Y + b2 + residual
      mbar_wait(y_full, tphase);
      tc_fence_after();
      const int row = tm * BM + r_in_tile;
      const bool row_ok = row < args.M;
      int sbuf = 0;
#pragma unroll 1
      for (int chunk = half; chunk < D / 32; chunk += 2) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld32(tmem + (uint32_t(quarter * 32) << 16) + chunk * 32, r);
        tmem_ld_wait();
        float v[32];
        const float* bias = args.b2 + chunk * 32;
#pragma unroll
        for (int c = 0; c < 32; c += 4) {
          const float4 bb = __ldg(reinterpret_cast<const float4*>(bias + c));
          v[c] = __uint_as_float(r[c]) + bb.x, v[c + 1] = __uint_as_float(r[c + 1]) + bb.y;
          v[c + 2] = __uint_as_float(r[c + 2]) + bb.z, v[c + 3] = __uint_as_float(r[c + 3]) + bb.w;
        }
        if (row_ok && args.residual) {
          const bf16* rr = args.residual + (long)row * args.ldr + chunk * 32;
#pragma unroll
          for (int c = 0; c < 32; c += 8) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(rr + c));
            v[c + 0] += bf16_lo(q.x), v[c + 1] += bf16_hi(q.x), v[c + 2] += bf16_lo(q.y), v[c + 3] += bf16_hi(q.y);
            v[c + 4] += bf16_lo(q.z), v[c + 5] += bf16_hi(q.z), v[c + 6] += bf16_lo(q.w), v[c + 7] += bf16_hi(q.w);
          }
        }
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
        uint8_t* sb = my_store + sbuf * 2048;
#pragma unroll
        for (int c = 0; c < 32; c += 8)
          *reinterpret_cast<uint4*>(sb + lane * 64 + c * 2) =
              make_uint4(pack_bf16(v[c], v[c + 1]), pack_bf16(v[c + 2], v[c + 3]), pack_bf16(v[c + 4], v[c + 5]),
                         pack_bf16(v[c + 6], v[c + 7]));
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&tmOut, sb, chunk * 32, tm * BM + quarter * 32);
          tma_store_commit();
        }
        sbuf ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(y_empty);
        tma_store_wait_read<0>();  // the staging area is the next tile's H buffer
      }
      __syncwarp();
      tphase ^= 1;
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

CUtensorMap map2d(const bf16* ptr, long rows, long cols, long ld, int box_rows, int box_cols, bool swizzle) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  if (!fn) throw Error(-4, "cuTensorMapEncodeTiled unavailable (no CUDA driver / GPU?)");
  N1_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && ld % 8 == 0, "fused_mlp: misaligned operand");
  CUtensorMap m;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(ptr), gdim, gstr, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(-5, "cuTensorMapEncodeTiled failed: " + std::to_string((int)r));
  return m;
}

template <int CM>
void launch_mlp(const bf16* x, int ldx, const bf16* w1, const bf16* w2, const MlpArgs& a, bf16* out, int ldo,
                cudaStream_t stream) {
  static std::once_flag once;
  static int max_clusters = 0;
  std::call_once(once, [] {
    cudaFuncSetAttribute(fused_mlp_kernel<CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    max_clusters = device_sm_count() / CM;
    if (CM > 1) {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(device_sm_count() / CM * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem;
      cudaLaunchAttribute at;
      at.id = cudaLaunchAttributeClusterDimension;
      at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
      cfg.attrs = &at, cfg.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, fused_mlp_kernel<CM>, &cfg) == cudaSuccess && n > 0) max_clusters = n;
    }
  });
  CUtensorMap tmX = map2d(x, a.M, D, ldx, BM, 64, true);
  CUtensorMap tmW1 = map2d(w1, F, D, D, HC / CM, 64, true);
  CUtensorMap tmW2 = map2d(w2, D, F, F, 192 / CM, 64, true);
  CUtensorMap tmOut = map2d(out, a.M, D, ldo, 32, 32, false);
  const int super_m = (a.tiles_m + CM - 1) / CM;
  const int clusters = super_m < max_clusters ? super_m : max_clusters;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CM), cfg.blockDim = dim3(kThreads), cfg.dynamicSmemBytes = kSmem, cfg.stream = stream;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = CM, at.val.clusterDim.y = 1, at.val.clusterDim.z = 1;
  cfg.attrs = &at, cfg.numAttrs = 1;
  N1_CUDA(cudaLaunchKernelEx(&cfg, fused_mlp_kernel<CM>, tmX, tmW1, tmW2, tmOut, a));
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace

// x [M, 384] bf16 (row stride ldx), w1 [1536, 384], w2 [384, 1536] bf16 contiguous, b1 [1536], b2 [384] fp32,
// residual [M, ldr] bf16 or null, out [M, ldo] bf16 (may alias residual).
void fused_mlp_384(const bf16* x, int ldx, const bf16* w1, const float* b1, const bf16* w2, const float* b2,
                   const bf16* residual, int ldr, bf16* out, int ldo, int M, int cluster, cudaStream_t stream) {
  if (M <= 0) return;
  MlpArgs a;
  a.M = M, a.tiles_m = (M + BM - 1) / BM;
  a.b1 = b1, a.b2 = b2, a.residual = residual, a.ldr = ldr;
  if (cluster >= 2 && a.tiles_m >= 2)
    launch_mlp<2>(x, ldx, w1, w2, a, out, ldo, stream);
  else
    launch_mlp<1>(x, ldx, w1, w2, a, out, ldo, stream);
}

}  // namespace n1
