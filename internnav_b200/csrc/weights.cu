// Weight intake and packing kernels (load-time only; not on the per-step path).
#include "weights.h"

#include "n1_ptx.cuh"

namespace n1 {

Arena::~Arena() {
  for (void* p : slabs_) cudaFree(p);
}

void* Arena::alloc(size_t bytes) {
  bytes = (bytes + 255) & ~size_t(255);
  if (bytes > left_) {
    const size_t slab = bytes > (size_t(64) << 20) ? bytes : (size_t(64) << 20);
    void* p = nullptr;
    N1_CUDA(cudaMalloc(&p, slab));
    slabs_.push_back(p);
    cur_ = static_cast<char*>(p);
    left_ = slab;
  }
  void* r = cur_;
  cur_ += bytes;
  left_ -= bytes;
  used_ += bytes;
  return r;
}

const SrcTensor& WeightSource::get(const std::string& name) const {
  auto it = map_->find(prefix_ + name);
  if (it == map_->end()) throw Error(-6, "missing weight tensor: " + prefix_ + name);
  return it->second;
}

namespace {

__device__ __forceinline__ float load_any(const void* p, int dtype, long i) {
  return dtype == 0 ? static_cast<const float*>(p)[i] : __bfloat162float(static_cast<const bf16*>(p)[i]);
}
__device__ __forceinline__ void store_any(void* p, int dtype, long i, float v) {
  if (dtype == 0)
    static_cast<float*>(p)[i] = v;
  else
    static_cast<bf16*>(p)[i] = __float2bfloat16(v);
}

__global__ void pack2d_kernel(const void* src, int sdt, long sld, long row0, long rows, long cols, void* dst, int ddt,
                              long dld) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * dld) return;
  const long r = i / dld, c = i % dld;
  store_any(dst, ddt, i, c < cols ? load_any(src, sdt, (row0 + r) * sld + c) : 0.f);
}
__global__ void fold_kernel(const void* src, int sdt, long rows, long groups, long cols, float* dst) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * cols) return;
  const long r = i / cols, c = i % cols;
  float s = 0.f;
  for (long g = 0; g < groups; ++g) s += load_any(src, sdt, r * groups * cols + g * cols + c);
  dst[i] = s;
}
__global__ void interleave_kernel(const void* a, const void* b, int sdt, long rows, long cols, void* dst, int ddt,
                                  long dld) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 2 * rows * dld) return;
  const long r = i / dld, c = i % dld;
  const void* src = (r & 1) ? b : a;
  store_any(dst, ddt, i, c < cols ? load_any(src, sdt, (r >> 1) * cols + c) : 0.f);
}

inline int nblk(long n) { return (int)((n + 255) / 256); }

}  // namespace

void pack2d(const void* src, int sdt, long sld, long row0, long rows, long cols, void* dst, int ddt, long dld,
            cudaStream_t s) {
  if (rows * dld == 0) return;
  pack2d_kernel<<<nblk(rows * dld), 256, 0, s>>>(src, sdt, sld, row0, rows, cols, dst, ddt, dld);
  N1_CUDA(cudaGetLastError());
}
void fold_groups(const void* src, int sdt, long rows, long groups, long cols, float* dst, cudaStream_t s) {
  fold_kernel<<<nblk(rows * cols), 256, 0, s>>>(src, sdt, rows, groups, cols, dst);
  N1_CUDA(cudaGetLastError());
}
void interleave_rows(const void* a, const void* b, int sdt, long rows, long cols, void* dst, int ddt, long dld,
                     cudaStream_t s) {
  interleave_kernel<<<nblk(2 * rows * dld), 256, 0, s>>>(a, b, sdt, rows, cols, dst, ddt, dld);
  N1_CUDA(cudaGetLastError());
}

float* WeightSource::f32(Arena& a, const std::string& name, cudaStream_t s) const {
  const SrcTensor& t = get(name);
  const long n = t.numel();
  float* d = a.alloc_n<float>(n);
  pack2d(t.data, t.dtype, n, 0, 1, n, d, 0, n, s);
  return d;
}
float* WeightSource::f32_rows(Arena& a, const std::string& name, long row0, long rows, long cols,
                              cudaStream_t s) const {
  const SrcTensor& t = get(name);
  N1_CHECK((row0 + rows) * cols <= t.numel(), "weight " + name + ": row slice out of range");
  float* d = a.alloc_n<float>(rows * cols);
  pack2d(t.data, t.dtype, cols, row0, rows, cols, d, 0, cols, s);
  return d;
}
bf16* WeightSource::mat(Arena& a, const std::string& name, long row0, long rows, long cols, int* ld_out,
                        cudaStream_t s) const {
  const SrcTensor& t = get(name);
  N1_CHECK((row0 + rows) * cols <= t.numel(), "weight " + name + ": row slice out of range");
  const long ld = (cols + 7) & ~7L;
  const long rows_pad = (rows + 7) & ~7L;
  bf16* d = a.alloc_n<bf16>(rows_pad * ld);
  N1_CUDA(cudaMemsetAsync(d, 0, rows_pad * ld * sizeof(bf16), s));
  pack2d(t.data, t.dtype, cols, row0, rows, cols, d, 1, ld, s);
  if (ld_out) *ld_out = (int)ld;
  return d;
}

}  // namespace n1
