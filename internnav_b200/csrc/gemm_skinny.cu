// Weight-streaming GEMM for M <= 64 rows: out[M, N'] = epilogue(A[M, K] @ W[N, K]^T).
//
// The decode passes of the System-2 greedy loop (s2_model.cu::chunk_pass with one token per sequence, and the lm_head)
// multiply a handful of rows by every decoder weight: 14 GB of W per pass and < 1 TFLOP, i.e. HBM-bound by a factor of
// ~40.  The tcgen05 GEMM tiles 128 x BN and leaves 28-56 CTAs on the narrow layers (o_proj, down_proj), which is too
// few to saturate HBM (measured: 8.8 ms per pass = 24.5 % of the copy bandwidth, profiles/r1_generate_decode_timing.json).
// Here the grid is (N / 32) x split-K CTAs of 128 threads, sized to >= 2 CTAs per SM; every CTA streams its 32 weight
// rows over its K range exactly once through a 6-stage cp.async ring (the A rows ride along from L2), multiplies on
// mma.sync.m16n8k16 (bf16, fp32 accumulate; the tensor rate is irrelevant here) and stores an fp32 partial tile.  A
// second kernel sums the split-K partials and applies bias / activation / residual in the same order as the main GEMM
// epilogue (gemm_tcgen05.cu) before rounding to bf16.
//
// Replaces: the cuBLAS GEMV-like calls inside `model.generate` of the reference (internvla_n1_policy.py L169-176).
#include <algorithm>

#include "n1_ops.h"
#include "n1_ptx.cuh"

namespace n1 {
namespace {

constexpr int SBN = 32;              // weight rows (output columns) per CTA
constexpr int SKC = 64;              // K elements per pipeline stage
constexpr int SROW = SKC * 2 + 16;   // padded smem row: 144 B keeps the 8 rows of an ldmatrix in distinct banks
constexpr int SROWS = 64 + SBN;      // A rows then W rows
constexpr int SSTAGES = 6;
constexpr int SSTAGE_BYTES = SROWS * SROW;
constexpr int SSMEM = SSTAGES * SSTAGE_BYTES;  // 82,944 B -> 2 CTAs per SM

__device__ __forceinline__ void cp16(void* smem_dst, const void* gsrc, bool valid) {
  const int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

struct SkinnyArgs {
  const bf16* A;
  const bf16* W;
  float* part;  // [splits, 64, N] fp32
  int lda, ldw, M, N, K, k_per;  // k_per: K elements per split (multiple of SKC)
};

__global__ void __launch_bounds__(128) skinny_gemm_kernel(const SkinnyArgs p) {
  extern __shared__ __align__(16) uint8_t ssm[];
  const int n0 = blockIdx.x * SBN, split = blockIdx.y;
  const int k_lo = split * p.k_per, k_hi = min(p.K, k_lo + p.k_per);
  const int chunks = (k_hi - k_lo + SKC - 1) / SKC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // one stage: rows 0..63 = A[row, k0 : k0 + 64], rows 64..95 = W[n0 + r, k0 : k0 + 64]; 8 x 16 B per row
  auto load_stage = [&](int c) {
    uint8_t* st = ssm + (c % SSTAGES) * SSTAGE_BYTES;
    const int k0 = k_lo + c * SKC;
    for (int i = threadIdx.x; i < SROWS * 8; i += 128) {
      const int r = i >> 3, ch = i & 7;
      const int k = k0 + ch * 8;
      const bool kin = k < k_hi;  // K % 8 == 0 and k_per % 64 == 0: a 16-byte piece is wholly inside or outside
      const bf16* src;
      bool ok;
      if (r < 64) {
        ok = kin && r < p.M;
        src = p.A + (long)(ok ? r : 0) * p.lda + (ok ? k : 0);
      } else {
        const int n = n0 + r - 64;
        ok = kin && n < p.N;
        src = p.W + (long)(ok ? n : 0) * p.ldw + (ok ? k : 0);
      }
      cp16(st + r * SROW + ch * 16, src, ok);
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;

#pragma unroll 1
  for (int c = 0; c < SSTAGES - 1; ++c) {
    if (c < chunks) load_stage(c);
    cp_commit();
  }
  const int lm = lane >> 3, lr = lane & 7;
#pragma unroll 1
  for (int c = 0; c < chunks; ++c) {
    cp_wait<SSTAGES - 2>();   // chunk c has landed (one group per chunk, SSTAGES - 1 groups in flight)
    __syncthreads();          // ... for every thread, and everyone is done with the stage about to be refilled
    if (c + SSTAGES - 1 < chunks) load_stage(c + SSTAGES - 1);
    cp_commit();
    const uint8_t* st = ssm + (c % SSTAGES) * SSTAGE_BYTES;
    const uint8_t* sA = st;
    const uint8_t* sW = st + 64 * SROW;
#pragma unroll
    for (int ks = 0; ks < SKC / 16; ++ks) {
      uint32_t a[4];
      ldsm4(smem_u32(sA + (warp * 16 + lr + (lm & 1) * 8) * SROW + (ks * 16 + (lm >> 1) * 8) * 2), a[0], a[1], a[2], a[3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {  // pairs of 8-column n-tiles
        uint32_t b0, b1, b2, b3;
        ldsm4(smem_u32(sW + (np * 16 + (lm >> 1) * 8 + lr) * SROW + (ks * 16 + (lm & 1) * 8) * 2), b0, b1, b2, b3);
        mma16816(acc[2 * np], a, b0, b1);
        mma16816(acc[2 * np + 1], a, b2, b3);
      }
    }
  }
  cp_wait<0>();

  // fp32 partial tile: rows warp*16 + lane/4 (+8), columns n0 + tile*8 + (lane%4)*2 (+1)
  float* out = p.part + (long)split * 64 * p.N;
  const int r0 = warp * 16 + (lane >> 2);
#pragma unroll
  for (int t = 0; t < 4; ++t) {
    const int col = n0 + t * 8 + (lane & 3) * 2;
    if (col < p.N) {  // N % 8 == 0: the pair is inside together
      *reinterpret_cast<float2*>(out + (long)r0 * p.N + col) = make_float2(acc[t][0], acc[t][1]);
      *reinterpret_cast<float2*>(out + (long)(r0 + 8) * p.N + col) = make_float2(acc[t][2], acc[t][3]);
    }
  }
}

struct FinishArgs {
  const float* part;
  const float* bias;
  const bf16* residual;
  bf16* out;
  int splits, M, N, ldr, ldo, act;
};

// two adjacent accumulator columns per thread; same operation order as the tcgen05 epilogue
__global__ void skinny_finish_kernel(const FinishArgs p) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const int pairs = p.N >> 1;
  if (i >= (long)p.M * pairs) return;
  const int m = (int)(i / pairs), c = (int)(i % pairs) * 2;
  float v0 = 0.f, v1 = 0.f;
  for (int s = 0; s < p.splits; ++s) {
    const float2 t = *reinterpret_cast<const float2*>(p.part + ((long)s * 64 + m) * p.N + c);
    v0 += t.x, v1 += t.y;
  }
  if (p.bias) v0 += __ldg(p.bias + c), v1 += __ldg(p.bias + c + 1);
  if (p.act == ACT_SWIGLU) {  // (gate, up) interleaved -> one output column
    float r = silu(v0) * v1;
    const int oc = c >> 1;
    if (p.residual) r += __bfloat162float(p.residual[(long)m * p.ldr + oc]);
    p.out[(long)m * p.ldo + oc] = __float2bfloat16(r);
    return;
  }
  if (p.act == ACT_GELU) v0 = gelu_erf(v0), v1 = gelu_erf(v1);
  else if (p.act == ACT_RELU) v0 = fmaxf(v0, 0.f), v1 = fmaxf(v1, 0.f);
  if (p.residual) {
    const uint32_t r = *reinterpret_cast<const uint32_t*>(p.residual + (long)m * p.ldr + c);
    v0 += bf16_lo(r), v1 += bf16_hi(r);
  }
  *reinterpret_cast<uint32_t*>(p.out + (long)m * p.ldo + c) = pack_bf16(v0, v1);
}

}  // namespace

// Largest fp32 scratch gemm_skinny may need for an N-column product (the split count is capped so that
// splits * N <= kSkinnyCols).
constexpr long kSkinnyCols = 262144;
size_t gemm_skinny_workspace_bytes() { return (size_t)64 * kSkinnyCols * sizeof(float); }

bool gemm_skinny_supported(int M, int N, int K, const GemmEpilogue& e) {
  return M > 0 && M <= 64 && N % 8 == 0 && K % 8 == 0 && !e.gamma && !e.row_add && e.rows_per_group == 0 && !e.out_fp32 &&
         (e.act != ACT_SWIGLU || N % 16 == 0) && N <= kSkinnyCols;
}

void gemm_skinny(const bf16* A, int lda, const bf16* W, int ldw, bf16* out, int ldo, int M, int N, int K,
                 const GemmEpilogue& e, float* ws, cudaStream_t stream) {
  N1_CHECK(gemm_skinny_supported(M, N, K, e), "gemm_skinny: unsupported shape / epilogue");
  N1_CHECK(ws != nullptr, "gemm_skinny: null workspace");
  N1_CHECK((reinterpret_cast<uintptr_t>(A) & 15) == 0 && (reinterpret_cast<uintptr_t>(W) & 15) == 0 && lda % 8 == 0 &&
               ldw % 8 == 0 && ldo % 2 == 0 && (!e.residual || e.ldr % 2 == 0),
           "gemm_skinny: misaligned operands");
  static bool attr_set = false;
  if (!attr_set) {
    N1_CUDA(cudaFuncSetAttribute(skinny_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SSMEM));
    attr_set = true;
  }
  const int tiles = (N + SBN - 1) / SBN;
  const int chunks = (K + SKC - 1) / SKC;
  // enough CTAs for two per SM, at least 8 chunks (512 K elements) per split, partials bounded by the workspace
  int splits = (2 * device_sm_count() + tiles - 1) / tiles;
  splits = std::max(1, std::min(splits, std::max(1, chunks / 8)));
  while (splits > 1 && (long)splits * N > kSkinnyCols) --splits;
  const int per = (chunks + splits - 1) / splits;
  splits = (chunks + per - 1) / per;  // no empty split
  SkinnyArgs a{A, W, ws, lda, ldw, M, N, K, per * SKC};
  skinny_gemm_kernel<<<dim3(tiles, splits), 128, SSMEM, stream>>>(a);
  N1_CUDA(cudaGetLastError());
  FinishArgs f{ws, e.bias, e.residual, out, splits, M, N, e.ldr, ldo, e.act};
  const long n = (long)M * (N / 2);
  skinny_finish_kernel<<<(int)((n + 255) / 256), 256, 0, stream>>>(f);
  N1_CUDA(cudaGetLastError());
  prof_count_gemm(2.0 * M * (double)N * K);
  prof_count_launch();
}

}  // namespace n1
