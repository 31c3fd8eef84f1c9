// Weight intake: named tensors (HF state_dict names) -> packed device buffers owned by the handle.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "n1_ops.h"

namespace n1 {

struct SrcTensor {
  const void* data = nullptr;  // device pointer (caller-owned, must stay valid during n1_load_*)
  int dtype = 0;               // 0 = fp32, 1 = bf16
  std::vector<long> shape;
  long numel() const {
    long n = 1;
    for (long s : shape) n *= s;
    return n;
  }
};

// Bump allocator over cudaMalloc'd slabs; everything a model packs lives here until n1_destroy.
class Arena {
 public:
  ~Arena();
  void* alloc(size_t bytes);
  template <typename T>
  T* alloc_n(size_t n) {
    return static_cast<T*>(alloc(n * sizeof(T)));
  }
  size_t bytes_used() const { return used_; }

 private:
  std::vector<void*> slabs_;
  char* cur_ = nullptr;
  size_t left_ = 0, used_ = 0;
};

class WeightSource {
 public:
  explicit WeightSource(std::string prefix = "")
      : prefix_(std::move(prefix)), map_(std::make_shared<std::map<std::string, SrcTensor>>()) {}
  void add(const std::string& name, const SrcTensor& t) { (*map_)[name] = t; }
  bool has(const std::string& name) const { return map_->count(prefix_ + name) != 0; }
  const SrcTensor& get(const std::string& name) const;
  WeightSource sub(const std::string& more) const {
    WeightSource w(prefix_ + more);
    w.map_ = map_;
    return w;
  }
  // fp32 copy of a whole tensor (or rows [row0, row0+rows) of a 2-D view with `cols` columns)
  float* f32(Arena& a, const std::string& name, cudaStream_t s) const;
  float* f32_rows(Arena& a, const std::string& name, long row0, long rows, long cols, cudaStream_t s) const;
  // bf16 GEMM weight [rows, ld] from rows [row0, row0+rows) of the 2-D view [*, cols]; ld = cols rounded up to 8,
  // padding zero-filled.  `out_rows_pad` rounds the row count up (zero rows) so N % 8 == 0.
  bf16* mat(Arena& a, const std::string& name, long row0, long rows, long cols, int* ld_out, cudaStream_t s) const;

 private:
  std::string prefix_;
  std::shared_ptr<std::map<std::string, SrcTensor>> map_;
};

// dst[r, c] = src[(row0 + r), c] converted; c in [cols, dst_ld) zero.
void pack2d(const void* src, int src_dtype, long src_ld, long row0, long rows, long cols, void* dst, int dst_dtype,
            long dst_ld, cudaStream_t s);
// dst[r, c] = sum_k src[r, k * cols + c]  (fp32 out; folds the replicated depth channels into the conv weight)
void fold_groups(const void* src, int src_dtype, long rows, long groups, long cols, float* dst, cudaStream_t s);
// dst[2j] = a[j], dst[2j+1] = b[j]   (row interleave for the SwiGLU epilogue), rows of `cols` elements
void interleave_rows(const void* a, const void* b, int src_dtype, long rows, long cols, void* dst, int dst_dtype,
                     long dst_ld, cudaStream_t s);

}  // namespace n1
