// tcgen05 attention for head_dim 128: the causal GQA prefill of the Qwen2.5-VL decoder (28 query / 4 K-V heads;
// call sites of the reference: internvla_n1.py L206 / L338 -> Qwen2_5_VLAttention with attn_implementation =
// "flash_attention_2", internvla_n1_policy.py L33-38).  Replaces the mma.sync kernel (attention.cu, attn_kernel<128>)
// for sequences of up to 320 keys, which covers the benchmark prompts (S = 304); longer sequences keep the old kernel.
//
// Work item = (sequence, K/V head).  K and V of the item (<= 320 rows x 128) are TMA-loaded ONCE into shared memory and
// serve the 7 query heads of the GQA group x ceil(S / 128) query tiles:
//
//   warp 0   TMA producer: K, V per item (128-byte swizzle, [rows][64] tiles), Q per query tile
//   warp 1   MMA issuer:   S = Q K^T   (tcgen05.mma M128 N128/N64 K16, fp32 scores in TMEM columns 0..319; causal tiles
//                                       above the diagonal are never issued)
//                          O += P V    (P read from TENSOR MEMORY -- the softmax warps write the bf16 probabilities over the
//                                       score columns they have just consumed, so P never touches shared memory and no
//                                       block waits for another; V consumed in place as an MN-major B operand -- no
//                                       transpose; fp32 O in TMEM columns 384..511)
//   warps 2-17 softmax + epilogue, 4 warps per TMEM lane quarter (thread = query row x one 32-key slice of a 128-key block;
//              two warps per scheduler could not hide the MUFU / tcgen05.ld latencies: profiles/README.md):
//              pass 1 row max (tcgen05.ld), pass 2 p = exp2((s - max) * scale), bf16 pairs back into TMEM (tcgen05.st),
//              row sums; finally O / sum -> global.  Scores never leave the chip; with <= 320 keys the whole score row is
//              resident in TMEM, so no online rescaling of O is needed.
#include <math.h>

#include <mutex>

#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {
namespace {

constexpr int HD = 128, BQ = 128, KMAX = 320;
constexpr int kKBlock = KMAX * 128;             // bytes of one 64-column half of K (or V): 320 rows x 128 B = 40960
constexpr int kKVBytes = 2 * kKBlock;           // 81920
constexpr int kQBytes = BQ * HD * 2;            // 32768: two [128 x 64] k-blocks
constexpr int kSoftWarps = 16;
constexpr int kThreads = 64 + 32 * kSoftWarps;  // 576
constexpr int kStatBytes = 4 * 128 * 4 * 2;     // row max and row sum exchange between the four 32-key slices of a row
constexpr int kSmem = 2 * kKVBytes + kQBytes + kStatBytes + 128 + 1024;
static_assert(kSmem <= 232448, "attention_tc: shared memory budget");

struct TcArgs {
  const int* cu_q;      // [batch + 1] token offsets (query == key sequences: self-attention prefill)
  int batch, heads_q, heads_kv;
  int causal;
  float scale_log2;     // softmax scale * log2(e)
  bf16* o;
  int ldo;
  int num_items;        // batch * heads_kv
};

__device__ __forceinline__ uint64_t desc_k_sw128(uint32_t addr) { return umma_desc_sw128(addr); }
// MN-major B operand (V as stored: [keys][64 head-dim values] rows of 128 bytes, 128-byte swizzle): 8 keys form a
// 1024-byte atom (SBO), the second 64-wide half of the head dimension lies LBO bytes further
// (cute/atom/mma_traits_sm100.hpp, canonical Major-MN B128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t desc_v_mn_sw128(uint32_t addr, uint32_t lbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_bf16_bmn(int M, int N) { return umma_idesc_bf16(M, N) | (1u << 16); }

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void soft_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kSoftWarps * 32) : "memory"); }

__global__ void __launch_bounds__(kThreads, 1)
attn_tc128_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                  const __grid_constant__ CUtensorMap tmV, const TcArgs args) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = sK + kKVBytes;
  uint8_t* sQ = sV + kKVBytes;
  float* sStat = reinterpret_cast<float*>(sQ + kQBytes);   // [2 halves][128 rows][max, sum]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sQ + kQBytes + kStatBytes);
  uint64_t* kv_full = bars;       // K and V of the item landed
  uint64_t* kv_empty = bars + 1;  // all MMAs of the item complete
  uint64_t* q_full = bars + 2;
  uint64_t* q_empty = bars + 3;   // the tile's Q K^T MMAs complete
  uint64_t* s_full = bars + 4;    // scores in TMEM
  uint64_t* p_full = bars + 5;    // P of one 128-key block is in TMEM (8 warp arrivals)
  uint64_t* o_full = bars + 6;    // O complete in TMEM
  uint64_t* o_empty = bars + 7;   // softmax warps finished reading O (8 warp arrivals): next P V may overwrite it
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int group = args.heads_q / args.heads_kv;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ), tma_prefetch_desc(&tmK), tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1), mbar_init(kv_empty, 1), mbar_init(q_full, 1), mbar_init(q_empty, 1), mbar_init(s_full, 1);
    mbar_init(p_full, kSoftWarps), mbar_init(o_full, 1), mbar_init(o_empty, kSoftWarps);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t kv_ph = 0, q_ph = 0;
      for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
        const int b = item / args.heads_kv, kh = item % args.heads_kv;
        const int start = args.cu_q[b], len = args.cu_q[b + 1] - start;
        mbar_wait(kv_empty, kv_ph ^ 1);
        kv_ph ^= 1;
        mbar_arrive_expect_tx(kv_full, 2 * kKVBytes);
        for (int kb = 0; kb < 2; ++kb)
          for (int r = 0; r < 2; ++r) {  // 320 rows as two boxes of 160 (rows past the sequence are masked / weigh 0)
            tma_load_2d(sK + kb * kKBlock + r * 160 * 128, &tmK, kv_full, kh * HD + kb * 64, start + r * 160);
            tma_load_2d(sV + kb * kKBlock + r * 160 * 128, &tmV, kv_full, kh * HD + kb * 64, start + r * 160);
          }
        const int q_tiles = (len + BQ - 1) / BQ;
        for (int hq = 0; hq < group; ++hq)
          for (int qt = 0; qt < q_tiles; ++qt) {
            mbar_wait(q_empty, q_ph ^ 1);
            q_ph ^= 1;
            mbar_arrive_expect_tx(q_full, kQBytes);
            for (int kb = 0; kb < 2; ++kb)
              tma_load_2d(sQ + kb * 16384, &tmQ, q_full, (kh * group + hq) * HD + kb * 64, start + qt * BQ);
          }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t id_s128 = umma_idesc_bf16(BQ, 128), id_s64 = umma_idesc_bf16(BQ, 64);
      constexpr uint32_t id_o = idesc_bf16_bmn(BQ, HD);
      uint32_t kv_ph = 0, q_ph = 0, p_ph = 0, oe_ph = 0;
      for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
        const int b = item / args.heads_kv;
        const int len = args.cu_q[b + 1] - args.cu_q[b];
        const int q_tiles = (len + BQ - 1) / BQ;
        mbar_wait(kv_full, kv_ph);
        kv_ph ^= 1;
        for (int hq = 0; hq < group; ++hq)
          for (int qt = 0; qt < q_tiles; ++qt) {
            // keys this query tile can see: causal -> up to the end of its own tile
            const int k_hi = args.causal ? min(len, (qt + 1) * BQ) : len;
            const int k_tiles = (k_hi + 127) / 128;
            mbar_wait(q_full, q_ph);
            q_ph ^= 1;
            // The score / P columns are free: this thread issued the previous tile's last P V only after the softmax warps
            // had consumed every score (p_full of its last block), and the tensor pipe executes MMAs in issue order, so
            // this Q K^T cannot overtake those P V.  (The previous tile's O may still be in its epilogue: overlapped.)
            tc_fence_after();
            for (int n = 0; n < k_tiles; ++n) {
              const bool narrow = (n == 2);  // keys 256..319: a 64-wide tile
              for (int kb = 0; kb < 2; ++kb) {
                const uint64_t ad = desc_k_sw128(smem_u32(sQ + kb * 16384));
                const uint64_t bd = desc_k_sw128(smem_u32(sK + kb * kKBlock + n * 128 * 128));
#pragma unroll
                for (int k = 0; k < 4; ++k)
                  umma_f16(tmem + n * 128, ad + 2 * k, bd + 2 * k, narrow ? id_s64 : id_s128, (kb | k) != 0 ? 1u : 0u);
              }
            }
            umma_commit(q_empty);
            umma_commit(s_full);
            for (int j = 0; j < k_tiles; ++j) {
              mbar_wait(p_full, p_ph);
              p_ph ^= 1;
              if (j == 0) {
                mbar_wait(o_empty, oe_ph ^ 1);  // the previous tile's O has been read out
                oe_ph ^= 1;
              }
              tc_fence_after();
              const int keys = (j == 2) ? 64 : 128;
              for (int ks = 0; ks < keys / 16; ++ks) {
                // P of the 32-key slice ks / 2 sits in the first 16 columns of that slice's score columns, 8 columns per k-step
                const uint32_t a_tmem = tmem + j * 128 + (ks >> 1) * 32 + (ks & 1) * 8;
                const uint64_t bd = desc_v_mn_sw128(smem_u32(sV + (j * 128 + ks * 16) * 128), kKBlock);
                umma_f16_ts(tmem + 384, a_tmem, bd, id_o, (j | ks) != 0 ? 1u : 0u);
              }
            }
            umma_commit(o_full);
          }
        umma_commit(kv_empty);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / epilogue warps
    const int sw = warp - 2;                // 0..15
    const int quarter = warp & 3;           // TMEM lane quarter (hardware rule: warp id % 4)
    const int sub = sw >> 2;                // which 32-key slice of every 128-key block / which 32 columns of O
    const int r_in_tile = quarter * 32 + lane;
    const uint32_t lane_base = uint32_t(quarter * 32) << 16;
    uint32_t s_ph = 0, o_ph = 0;
    for (int item = blockIdx.x; item < args.num_items; item += gridDim.x) {
      const int b = item / args.heads_kv, kh = item % args.heads_kv;
      const int start = args.cu_q[b], len = args.cu_q[b + 1] - start;
      const int q_tiles = (len + BQ - 1) / BQ;
      for (int hq = 0; hq < group; ++hq)
        for (int qt = 0; qt < q_tiles; ++qt) {
          const int k_hi = args.causal ? min(len, (qt + 1) * BQ) : len;
          const int k_tiles = (k_hi + 127) / 128;
          const int q_pos = qt * BQ + r_in_tile;                       // position of this thread's query row
          const int vis = args.causal ? min(len, q_pos + 1) : len;     // keys [0, vis) are visible to it
          mbar_wait(s_full, s_ph);
          s_ph ^= 1;
          tc_fence_after();
          // ---- pass 1: row maximum over this warp's 32-key slices
          float mx = -INFINITY;
          for (int j = 0; j < k_tiles; ++j) {
            const int k0 = j * 128 + sub * 32;
            if (k0 >= 320) break;  // the third block has 64 keys only
            uint32_t r[32];
            tmem_ld32(tmem + lane_base + k0, r);
            tmem_ld_wait();
            if (k0 + 32 <= vis) {  // slice fully visible to this row (warp-divergent only on the diagonal block)
#pragma unroll
              for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
            } else {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (k0 + i < vis) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
          }
          sStat[(sub * 128 + r_in_tile) * 2] = mx;
          soft_barrier();
#pragma unroll
          for (int o = 1; o < 4; ++o) mx = fmaxf(mx, sStat[(((sub + o) & 3) * 128 + r_in_tile) * 2]);
          const float m_s = (mx == -INFINITY) ? 0.f : mx * args.scale_log2;
          // ---- pass 2: probabilities, written back over the scores as bf16 pairs (the A operand of P V)
          float sum = 0.f;
          for (int j = 0; j < k_tiles; ++j) {
            const int k0 = j * 128 + sub * 32;
            if (k0 < 320) {
              uint32_t r[32], pk[16];
              tmem_ld32(tmem + lane_base + k0, r);
              tmem_ld_wait();
              if (k0 + 32 <= vis) {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float p0 = fast_exp2(__uint_as_float(r[i]) * args.scale_log2 - m_s);
                  const float p1 = fast_exp2(__uint_as_float(r[i + 1]) * args.scale_log2 - m_s);
                  sum += p0 + p1;
                  pk[i / 2] = pack_bf16(p0, p1);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; i += 2) {
                  const float p0 = (k0 + i < vis) ? fast_exp2(__uint_as_float(r[i]) * args.scale_log2 - m_s) : 0.f;
                  const float p1 = (k0 + i + 1 < vis) ? fast_exp2(__uint_as_float(r[i + 1]) * args.scale_log2 - m_s) : 0.f;
                  sum += p0 + p1;
                  pk[i / 2] = pack_bf16(p0, p1);
                }
              }
              // 16 packed columns over the first half of this slice's own 32 score columns (already consumed)
              tmem_st16(tmem + lane_base + k0, pk);
              tmem_st_wait();
            }
            tc_fence_before();
            mbar_arrive_elect(p_full);
          }
          sStat[(sub * 128 + r_in_tile) * 2 + 1] = sum;
          soft_barrier();
#pragma unroll
          for (int o = 1; o < 4; ++o) sum += sStat[(((sub + o) & 3) * 128 + r_in_tile) * 2 + 1];
          const float inv = sum > 0.f ? 1.0f / sum : 0.f;
          // ---- epilogue: O / sum -> global (this warp: 32 of the 128 head-dim columns of its 32 rows)
          mbar_wait(o_full, o_ph);
          o_ph ^= 1;
          tc_fence_after();
          {
            uint32_t r[32];
            tmem_ld32(tmem + lane_base + 384 + sub * 32, r);
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive_elect(o_empty);
            if (q_pos < len) {
              bf16* orow = args.o + (long)(start + q_pos) * args.ldo + (kh * group + hq) * HD + sub * 32;
#pragma unroll
              for (int i = 0; i < 32; i += 8) {
                uint4 v;
                v.x = pack_bf16(__uint_as_float(r[i]) * inv, __uint_as_float(r[i + 1]) * inv);
                v.y = pack_bf16(__uint_as_float(r[i + 2]) * inv, __uint_as_float(r[i + 3]) * inv);
                v.z = pack_bf16(__uint_as_float(r[i + 4]) * inv, __uint_as_float(r[i + 5]) * inv);
                v.w = pack_bf16(__uint_as_float(r[i + 6]) * inv, __uint_as_float(r[i + 7]) * inv);
                *reinterpret_cast<uint4*>(orow + i) = v;
              }
            }
          }
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace

bool attention_tc_supported(const AttnParams& p) {
  return p.hd == HD && p.cu_q != nullptr && p.cu_k == p.cu_q && !p.k_len && p.kv_div == 1 && p.max_seq_q > 0 &&
         p.max_seq_q <= KMAX && p.total_rows > 0 && p.heads_q % p.heads_kv == 0 && p.ldq % 8 == 0 && p.ldk % 8 == 0 &&
         p.ldv % 8 == 0 && p.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(p.q) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.k) & 15) == 0 && (reinterpret_cast<uintptr_t>(p.v) & 15) == 0 &&
         (reinterpret_cast<uintptr_t>(p.o) & 15) == 0;
}

void attention_tc128(const AttnParams& p, cudaStream_t stream) {
  N1_CHECK(attention_tc_supported(p), "attention_tc128: unsupported arguments");
  static std::once_flag once;
  std::call_once(once, [] { cudaFuncSetAttribute(attn_tc128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem); });
  const long rows = p.total_rows;
  CUtensorMap tmQ = tma_map_2d(p.q, rows, (long)p.heads_q * HD, p.ldq, BQ, 64, true);
  CUtensorMap tmK = tma_map_2d(p.k, rows, (long)p.heads_kv * HD, p.ldk, 160, 64, true);
  CUtensorMap tmV = tma_map_2d(p.v, rows, (long)p.heads_kv * HD, p.ldv, 160, 64, true);
  TcArgs a;
  a.cu_q = p.cu_q, a.batch = p.batch, a.heads_q = p.heads_q, a.heads_kv = p.heads_kv, a.causal = p.causal;
  a.scale_log2 = p.scale * 1.4426950408889634f;
  a.o = p.o, a.ldo = p.ldo, a.num_items = p.batch * p.heads_kv;
  const int grid = a.num_items < device_sm_count() ? a.num_items : device_sm_count();
  attn_tc128_kernel<<<grid, kThreads, kSmem, stream>>>(tmQ, tmK, tmV, a);
  prof_count_launch();
  N1_CUDA(cudaGetLastError());
}

}  // namespace n1
