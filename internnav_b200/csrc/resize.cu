// Pillow-exact batched bicubic resampling for the System-1 frames.  See resize.h.
#include "resize.h"

#include <math.h>

namespace n1 {

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS

double bicubic_filter(double x) {
  const double a = -0.5;
  if (x < 0.0) x = -x;
  if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
  if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
  return 0.0;
}

template <typename T>
T* to_device(const std::vector<T>& v, cudaStream_t s) {
  T* d = nullptr;
  N1_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
  N1_CUDA(cudaMemcpyAsync(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  return d;
}

// One thread per output pixel (all 3 channels).  in: [n, rows, in_len, 3] resampled along in_len when `horizontal`,
// else [n, in_len, cols, 3] resampled along in_len.  Generic strides keep one kernel for both passes.
__global__ void resample_u8_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out_u8,
                                   float* __restrict__ out_f32, const int32_t* __restrict__ bounds,
                                   const int32_t* __restrict__ kk, int ksize, long total, int out_len, int other,
                                   long in_img, long out_img, long in_step, long in_other, long out_step, long out_other,
                                   int fast_other) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // consecutive threads walk along image columns in both passes (coalesced rows)
  const int o = fast_other ? (int)((i / other) % out_len) : (int)(i % out_len);   // index along the resampled axis
  const int q = fast_other ? (int)(i % other) : (int)((i / out_len) % other);     // index along the untouched axis
  const long img = i / ((long)out_len * other);
  const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
  const int32_t* k = kk + (long)o * ksize;
  const uint8_t* p = in + img * in_img + (long)q * in_other + (long)lo * in_step;
  int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0;
  for (int x = 0; x < cnt; ++x) {
    const int w = k[x];
    s0 += (int)p[0] * w, s1 += (int)p[1] * w, s2 += (int)p[2] * w;
    p += in_step;
  }
  const int c0 = min(max(s0 >> kPrecisionBits, 0), 255), c1 = min(max(s1 >> kPrecisionBits, 0), 255),
            c2 = min(max(s2 >> kPrecisionBits, 0), 255);
  const long d = img * out_img + (long)q * out_other + (long)o * out_step;
  if (out_u8) out_u8[d] = (uint8_t)c0, out_u8[d + 1] = (uint8_t)c1, out_u8[d + 2] = (uint8_t)c2;
  if (out_f32) {
    out_f32[d] = __fdiv_rn((float)c0, 255.0f), out_f32[d + 1] = __fdiv_rn((float)c1, 255.0f);
    out_f32[d + 2] = __fdiv_rn((float)c2, 255.0f);
  }
}

__global__ void u8_to_unit_kernel(const uint8_t* __restrict__ in, float* __restrict__ out, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __fdiv_rn((float)in[i], 255.0f);
}

// float images: double accumulation in window order, no fused multiply-add (the C reference is compiled without it)
__global__ void resample_f32_kernel(const float* __restrict__ in, float* __restrict__ out,
                                    const int32_t* __restrict__ bounds, const double* __restrict__ kk, int ksize,
                                    long total, int out_len, int other, long in_img, long out_img, long in_step,
                                    long in_other, long out_step, long out_other, int finish, float mul, float clip_max,
                                    int fast_other) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // consecutive threads walk along image columns in both passes (coalesced rows)
  const int o = fast_other ? (int)((i / other) % out_len) : (int)(i % out_len);
  const int q = fast_other ? (int)(i % other) : (int)((i / out_len) % other);
  const long img = i / ((long)out_len * other);
  const int lo = bounds[2 * o], cnt = bounds[2 * o + 1];
  const double* k = kk + (long)o * ksize;
  const float* p = in + img * in_img + (long)q * in_other + (long)lo * in_step;
  double acc = 0.0;
  for (int x = 0; x < cnt; ++x) {
    acc = __dadd_rn(acc, __dmul_rn((double)p[0], k[x]));
    p += in_step;
  }
  float r = (float)acc;
  if (finish) {
    r = __fmul_rn(r, mul);
    if (r > clip_max) r = clip_max;  // `d[d > 5.0] = 5.0`: NaN stays NaN
  }
  out[img * out_img + (long)q * out_other + (long)o * out_step] = r;
}

__global__ void scale_clip_kernel(const float* __restrict__ in, float* __restrict__ out, long n, float mul,
                                  float clip_max) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    float r = __fmul_rn(in[i], mul);
    if (r > clip_max) r = clip_max;
    out[i] = r;
  }
}

inline int nblk(long n) { return (int)((n + 255) / 256); }

}  // namespace

void resize_coeffs(int in_size, int out_size, ResizeCoeffs& c) {
  N1_CHECK(in_size > 0 && out_size > 0, "resize: sizes must be positive");
  c.in_size = in_size, c.out_size = out_size;
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 2.0 * filterscale;
  c.ksize = (int)ceil(support) * 2 + 1;
  c.bounds.assign((size_t)out_size * 2, 0);
  c.weights.assign((size_t)out_size * c.ksize, 0.0);
  c.fixed.assign((size_t)out_size * c.ksize, 0);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int cnt = xmax - xmin;
    double* k = &c.weights[(size_t)xx * c.ksize];
    double ww = 0.0;
    for (int x = 0; x < cnt; ++x) {
      const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < cnt; ++x)
      if (ww != 0.0) k[x] /= ww;
    c.bounds[2 * xx] = xmin, c.bounds[2 * xx + 1] = cnt;
    for (int x = 0; x < c.ksize; ++x) {
      const double v = k[x];
      c.fixed[(size_t)xx * c.ksize + x] = v < 0 ? (int)(-0.5 + v * (double)(1 << kPrecisionBits))
                                                : (int)(0.5 + v * (double)(1 << kPrecisionBits));
    }
  }
}

ResizePlan::ResizePlan(int ih, int iw, int oh, int ow, cudaStream_t s) : in_h(ih), in_w(iw), out_h(oh), out_w(ow) {
  resize_coeffs(iw, ow, h);
  resize_coeffs(ih, oh, v);
  h_bounds = to_device(h.bounds, s), v_bounds = to_device(v.bounds, s);
  h_fixed = to_device(h.fixed, s), v_fixed = to_device(v.fixed, s);
  h_w = to_device(h.weights, s), v_w = to_device(v.weights, s);
  N1_CUDA(cudaStreamSynchronize(s));
}
ResizePlan::~ResizePlan() {
  cudaFree(h_bounds), cudaFree(v_bounds), cudaFree(h_fixed), cudaFree(v_fixed), cudaFree(h_w), cudaFree(v_w);
}
size_t ResizePlan::workspace_bytes(int n, int channels, bool is_float) const {
  return ((size_t)n * in_h * out_w * channels * (is_float ? sizeof(float) : 1) + 255) & ~size_t(255);
}

void resize_rgb_u8(const ResizePlan& p, const uint8_t* src, int n, float* dst_f32, uint8_t* dst_u8, void* ws,
                   cudaStream_t s) {
  N1_CHECK(src && (dst_f32 || dst_u8) && n > 0, "resize_rgb_u8: null arguments");
  const bool do_h = p.in_w != p.out_w, do_v = p.in_h != p.out_h;  // Resample.c skips a pass whose size is unchanged
  N1_CHECK(ws || !(do_h && do_v), "resize_rgb_u8: workspace required");
  const uint8_t* cur = src;
  if (do_h) {
    uint8_t* o8 = do_v ? static_cast<uint8_t*>(ws) : dst_u8;
    float* of = do_v ? nullptr : dst_f32;
    const long total = (long)n * p.in_h * p.out_w;
    resample_u8_kernel<<<nblk(total), 256, 0, s>>>(cur, o8, of, p.h_bounds, p.h_fixed, p.h.ksize, total, p.out_w, p.in_h,
                                                  (long)p.in_h * p.in_w * 3, (long)p.in_h * p.out_w * 3, 3,
                                                  (long)p.in_w * 3, 3, (long)p.out_w * 3, 0);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
    cur = o8;
  }
  if (do_v) {
    const long total = (long)n * p.out_h * p.out_w;
    // resampled axis = rows: out_len = out_h, other = out_w
    resample_u8_kernel<<<nblk(total), 256, 0, s>>>(cur, dst_u8, dst_f32, p.v_bounds, p.v_fixed, p.v.ksize, total, p.out_h,
                                                  p.out_w, (long)p.in_h * p.out_w * 3, (long)p.out_h * p.out_w * 3,
                                                  (long)p.out_w * 3, 3, (long)p.out_w * 3, 3, 1);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
  }
  if (!do_h && !do_v) {
    const long cnt = (long)n * p.in_h * p.in_w * 3;
    if (dst_u8) N1_CUDA(cudaMemcpyAsync(dst_u8, src, cnt, cudaMemcpyDeviceToDevice, s));
    if (dst_f32) {
      u8_to_unit_kernel<<<nblk(cnt), 256, 0, s>>>(src, dst_f32, cnt);
      prof_count_launch();
      N1_CUDA(cudaGetLastError());
    }
  }
}

void resize_f32(const ResizePlan& p, const float* src, int n, float mul, float clip_max, float* dst, void* ws,
                cudaStream_t s) {
  N1_CHECK(src && dst && n > 0, "resize_f32: null arguments");
  const bool do_h = p.in_w != p.out_w, do_v = p.in_h != p.out_h;
  N1_CHECK(ws || !(do_h && do_v), "resize_f32: workspace required");
  const float* cur = src;
  if (do_h) {
    float* o = do_v ? static_cast<float*>(ws) : dst;
    const long total = (long)n * p.in_h * p.out_w;
    resample_f32_kernel<<<nblk(total), 256, 0, s>>>(cur, o, p.h_bounds, p.h_w, p.h.ksize, total, p.out_w, p.in_h,
                                                   (long)p.in_h * p.in_w, (long)p.in_h * p.out_w, 1, p.in_w, 1, p.out_w,
                                                   do_v ? 0 : 1, mul, clip_max, 0);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
    cur = o;
  }
  if (do_v) {
    const long total = (long)n * p.out_h * p.out_w;
    resample_f32_kernel<<<nblk(total), 256, 0, s>>>(cur, dst, p.v_bounds, p.v_w, p.v.ksize, total, p.out_h, p.out_w,
                                                   (long)p.in_h * p.out_w, (long)p.out_h * p.out_w, p.out_w, 1, p.out_w,
                                                   1, 1, mul, clip_max, 1);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
  }
  if (!do_h && !do_v) {
    const long cnt = (long)n * p.in_h * p.in_w;
    scale_clip_kernel<<<nblk(cnt), 256, 0, s>>>(src, dst, cnt, mul, clip_max);
    prof_count_launch();
    N1_CUDA(cudaGetLastError());
  }
}

}  // namespace n1
