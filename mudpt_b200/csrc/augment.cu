// Image transform in front of the vision tower: crop -> bicubic resample (Pillow's
// antialiased separable filter) -> window -> horizontal flip -> /255 -> (x - mean) / std, from ragged 8-bit
// RGB images (HWC) to the fp32 [B, 3, out_h, out_w] batch MuDPT.parse_batch_train hands to the model
// (trainers/mudpt.py:263-268).
//
// The reference gets this tensor from Dassl's transform builder (un-vendored; the yaml names
// "random_resized_crop", "random_flip", "normalize" with bicubic interpolation,
// configs/trainers/MuDPT/vit_b16_bz4_ep10_nctx2_depth9.yaml:8-13), i.e. torchvision transforms on PIL
// images on the data-loader's CPU workers.  The arithmetic restated here is Pillow's
// libImaging/Resample.c (precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal_8bpc /
// Vertical_8bpc, bicubic_filter a = -0.5) and torchvision's to_tensor / normalize: results are
// BIT-IDENTICAL to that pipeline (tests/test_input_pipeline.py), which takes
//   * the filter weights in double precision with the C expression order and no fused multiply-add
//     (__dmul_rn / __dadd_rn / __ddiv_rn), converted to 22-bit fixed point exactly as Pillow does;
//   * an 8-bit rounded intermediate image between the horizontal and the vertical pass;
//   * IEEE fp32 division for /255 and /std (a 256-entry table per channel, built with __fdiv_rn: the
//     library is compiled with --use_fast_math).
// Random parameters (crop box, flip) are drawn on the host by mudpt_b200/input_pipeline.py.
//
// Two kernels per batch, both HBM/latency-bound byte work (no tensor cores):
//   resample_coeffs_kernel  one CTA per (image, axis): bounds + fixed-point weights of every output
//                           coordinate of the window, [tap][coordinate] layout
//   augment_smem_kernel     one CTA per (band of 16 output rows, image): the crop rows the band needs are
//                           staged in shared memory with 16-byte loads; horizontal pass from there
//                           (thread = output column, its filter weights in registers, loop over rows)
//                           into an 8-bit row buffer; vertical pass (weights in shared memory) + table
//                           lookup, coalesced fp32 stores to the three planes
//   augment_kernel          the same two passes reading the crop straight from global memory: used when
//                           the staged rows would not fit (very large / heavily down-scaled images)
#include "augment.h"

#include <cmath>

#include "common.cuh"
#include "launch_count.h"

namespace mudpt {

static constexpr int kPrecisionBits = 32 - 8 - 2;  // Resample.c PRECISION_BITS
static constexpr int kAugThreads = 256;

struct AxisGeom {
  int in_size;   // length of the cropped axis (resampling clamps at the crop, not at the image)
  int out_size;  // length it is resampled to
  int win0;      // first resampled coordinate of the output window
};

__device__ __forceinline__ AxisGeom axis_geom(const mudpt_image_desc& d, int axis) {
  AxisGeom g;
  g.in_size = axis == 0 ? d.box_w : d.box_h;
  g.out_size = axis == 0 ? d.rs_w : d.rs_h;
  g.win0 = axis == 0 ? d.win_x : d.win_y;
  return g;
}

// Resample.c:bicubic_filter, a = -0.5; every operation rounded separately (x86-64 C has no contraction)
__device__ __forceinline__ double bicubic_filter(double x) {
  if (x < 0.0) x = -x;
  if (x < 1.0) return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(1.5, x), -2.5), x), x), 1.0);
  if (x < 2.0) return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), -0.5);
  return 0.0;
}

// bounds[(img * 2 + axis) * out_len + t] = (first input coordinate, taps); kk[((img * 2 + axis) * kmax + tap) * out_len + t]
__global__ void __launch_bounds__(kAugThreads) resample_coeffs_kernel(const mudpt_image_desc* __restrict__ descs,
                                                                      int2* __restrict__ bounds, int* __restrict__ kk,
                                                                      int out_h, int out_w, int out_len, int kmax) {
  const int img = blockIdx.x >> 1, axis = blockIdx.x & 1;
  const mudpt_image_desc d = descs[img];
  const AxisGeom g = axis_geom(d, axis);
  const int n_out = axis == 0 ? out_w : out_h;
  // Resample.c:precompute_coeffs with in0 = 0, in1 = in_size (the crop is a new image)
  const double scale = __ddiv_rn(static_cast<double>(g.in_size), static_cast<double>(g.out_size));
  const double filterscale = scale < 1.0 ? 1.0 : scale;
  const double support = __dmul_rn(2.0, filterscale);
  const double ss = __ddiv_rn(1.0, filterscale);
  for (int t = threadIdx.x; t < n_out; t += blockDim.x) {
    const int xx = g.win0 + t;
    const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
    int xmin = static_cast<int>(__dadd_rn(__dadd_rn(center, -support), 0.5));  // C cast: toward zero
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > g.in_size) xmax = g.in_size;
    xmax -= xmin;
    if (xmax > kmax) xmax = kmax;  // cannot happen: the host sizes kmax from the same formula (ksize)
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
      ww = __dadd_rn(ww, bicubic_filter(__dmul_rn(__dadd_rn(__dadd_rn(static_cast<double>(x + xmin), -center), 0.5), ss)));
    int* k = kk + static_cast<size_t>(img * 2 + axis) * kmax * out_len + t;
    for (int x = 0; x < kmax; ++x) {
      int c = 0;
      if (x < xmax) {
        double w = bicubic_filter(__dmul_rn(__dadd_rn(__dadd_rn(static_cast<double>(x + xmin), -center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        // normalize_coeffs_8bpc
        c = w < 0.0 ? static_cast<int>(__dadd_rn(-0.5, __dmul_rn(w, static_cast<double>(1 << kPrecisionBits))))
                    : static_cast<int>(__dadd_rn(0.5, __dmul_rn(w, static_cast<double>(1 << kPrecisionBits))));
      }
      k[static_cast<size_t>(x) * out_len] = c;
    }
    bounds[static_cast<size_t>(img * 2 + axis) * out_len + t] = make_int2(xmin, xmax);
  }
}

__device__ __forceinline__ uint32_t clip8(int acc) {
  const int v = acc >> kPrecisionBits;  // arithmetic shift, as the C code's lookup index
  return static_cast<uint32_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// shared memory: [3][256] fp32 tables | tmp[rmax][out_w * 3] 8-bit horizontally resampled rows
template <int TR>
__global__ void __launch_bounds__(kAugThreads) augment_kernel(const mudpt_image_desc* __restrict__ descs,
                                                              const int2* __restrict__ bounds, const int* __restrict__ kk,
                                                              float* __restrict__ out, int out_h, int out_w, int out_len,
                                                              int kmax, int rmax, float m0, float m1, float m2, float s0,
                                                              float s1, float s2) {
  extern __shared__ __align__(16) uint8_t aug_smem[];
  float* lut = reinterpret_cast<float*>(aug_smem);
  uint8_t* tmp = aug_smem + 3 * 256 * sizeof(float);
  const int img = blockIdx.y, band = blockIdx.x;
  const mudpt_image_desc d = descs[img];
  const int y_first = band * TR;
  const int y_last = min(y_first + TR, out_h) - 1;
  // to_tensor (/255) then normalize ((x - mean) / std): one table entry per 8-bit value and channel
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int ch = i >> 8;
    const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2), sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), mean), sd);
  }
  const int2* bh = bounds + static_cast<size_t>(img * 2 + 0) * out_len;
  const int2* bv = bounds + static_cast<size_t>(img * 2 + 1) * out_len;
  const int* kh = kk + static_cast<size_t>(img * 2 + 0) * kmax * out_len;
  const int* kv = kk + static_cast<size_t>(img * 2 + 1) * kmax * out_len;
  // input rows (relative to the crop) this band reads: bounds are monotone in the output coordinate
  const int2 b_first = bv[y_first], b_last = bv[y_last];
  const int r0 = b_first.x;
  const int rows = b_last.x + b_last.y - r0;
  if (rows > rmax) __trap();  // host-side bound violated: fail loudly rather than corrupt shared memory
  const int row_bytes = out_w * 3;
  const uint8_t* src = d.src + static_cast<size_t>(d.box_y + r0) * d.pitch + static_cast<size_t>(d.box_x) * 3;
  // ---- horizontal pass: tmp[r][c] = clip8(2^21 + sum_k src[r][xmin_c + k] * kh[k][c])
  for (int idx = threadIdx.x; idx < rows * out_w; idx += blockDim.x) {
    const int r = idx / out_w, c = idx - r * out_w;
    const int2 b = bh[c];
    const uint8_t* p = src + static_cast<size_t>(r) * d.pitch + b.x * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int k = 0; k < b.y; ++k) {
      const int w = __ldg(kh + static_cast<size_t>(k) * out_len + c);
      a0 += static_cast<int>(__ldg(p + 3 * k)) * w;
      a1 += static_cast<int>(__ldg(p + 3 * k + 1)) * w;
      a2 += static_cast<int>(__ldg(p + 3 * k + 2)) * w;
    }
    uint8_t* q = tmp + r * row_bytes + c * 3;
    q[0] = static_cast<uint8_t>(clip8(a0));
    q[1] = static_cast<uint8_t>(clip8(a1));
    q[2] = static_cast<uint8_t>(clip8(a2));
  }
  __syncthreads();
  // ---- vertical pass + table lookup + (flipped) store
  const int band_rows = y_last - y_first + 1;
  for (int idx = threadIdx.x; idx < band_rows * out_w; idx += blockDim.x) {
    const int yl = idx / out_w, c = idx - yl * out_w;
    const int y = y_first + yl;
    const int2 b = bv[y];
    const uint8_t* p = tmp + (b.x - r0) * row_bytes + c * 3;
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
    for (int k = 0; k < b.y; ++k) {
      const int w = __ldg(kv + static_cast<size_t>(k) * out_len + y);
      a0 += static_cast<int>(p[0]) * w;
      a1 += static_cast<int>(p[1]) * w;
      a2 += static_cast<int>(p[2]) * w;
      p += row_bytes;
    }
    const int oc = d.flip ? out_w - 1 - c : c;
    const size_t plane = static_cast<size_t>(out_h) * out_w;
    float* o = out + static_cast<size_t>(img) * 3 * plane + static_cast<size_t>(y) * out_w + oc;
    o[0] = lut[clip8(a0)];
    o[plane] = lut[256 + clip8(a1)];
    o[2 * plane] = lut[512 + clip8(a2)];
  }
}

// Horizontal pass of one output column over all staged rows, filter weights (<= KR taps) in registers.
// Taps beyond the column's count have weight 0 and read bytes that lie inside the shared-memory allocation.
template <int KR>
__device__ __forceinline__ void hpass_column(const uint8_t* __restrict__ srcb, uint8_t* __restrict__ tmp, const int* __restrict__ kh,
                                             int2 b, int c, int rows, int src_stride, int row_bytes, int out_len,
                                             uintptr_t row_addr, int pitch) {
  int wk[KR];
#pragma unroll
  for (int k = 0; k < KR; ++k) wk[k] = k < b.y ? __ldg(kh + static_cast<size_t>(k) * out_len + c) : 0;
  const uint8_t* col = srcb + b.x * 3;
  uint8_t* q = tmp + c * 3;
#pragma unroll 2
  for (int r = 0; r < rows; ++r) {
    const uint8_t* p = col + r * src_stride + static_cast<int>(row_addr & 15);  // the row's 16-byte staging skew
    int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll
    for (int k = 0; k < KR; ++k) {
      a0 += static_cast<int>(p[3 * k]) * wk[k];
      a1 += static_cast<int>(p[3 * k + 1]) * wk[k];
      a2 += static_cast<int>(p[3 * k + 2]) * wk[k];
    }
    q[0] = static_cast<uint8_t>(clip8(a0));
    q[1] = static_cast<uint8_t>(clip8(a1));
    q[2] = static_cast<uint8_t>(clip8(a2));
    q += row_bytes;
    row_addr += pitch;
  }
}

// shared memory: [3][256] fp32 tables | kvs[TR][kmax] vertical weights | srcb[rmax][src_stride] staged crop rows
//                | tmp[rmax][out_w * 3] horizontally resampled rows (8-bit)
// srcb row r holds the 16-byte aligned span of global memory that covers crop row r0 + r: byte j of the row is
// at srcb[r * src_stride + (address of the row's first byte & 15) + j].
template <int TR>
__global__ void __launch_bounds__(kAugThreads) augment_smem_kernel(const mudpt_image_desc* __restrict__ descs,
                                                                   const int2* __restrict__ bounds, const int* __restrict__ kk,
                                                                   float* __restrict__ out, int out_h, int out_w, int out_len,
                                                                   int kmax, int rmax, int src_stride, float m0, float m1,
                                                                   float m2, float s0, float s1, float s2) {
  extern __shared__ __align__(16) uint8_t aug_smem[];
  float* lut = reinterpret_cast<float*>(aug_smem);
  int* kvs = reinterpret_cast<int*>(aug_smem + 3 * 256 * sizeof(float));
  uint8_t* srcb = reinterpret_cast<uint8_t*>(kvs + TR * kmax);
  uint8_t* tmp = srcb + static_cast<size_t>(rmax) * src_stride;
  const int img = blockIdx.y, band = blockIdx.x;
  const mudpt_image_desc d = descs[img];
  const int y_first = band * TR;
  const int y_last = min(y_first + TR, out_h) - 1;
  const int band_rows = y_last - y_first + 1;
  for (int i = threadIdx.x; i < 3 * 256; i += blockDim.x) {
    const int ch = i >> 8;
    const float mean = ch == 0 ? m0 : (ch == 1 ? m1 : m2), sd = ch == 0 ? s0 : (ch == 1 ? s1 : s2);
    lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(i & 255), 255.0f), mean), sd);
  }
  const int2* bh = bounds + static_cast<size_t>(img * 2 + 0) * out_len;
  const int2* bv = bounds + static_cast<size_t>(img * 2 + 1) * out_len;
  const int* kh = kk + static_cast<size_t>(img * 2 + 0) * kmax * out_len;
  const int* kv = kk + static_cast<size_t>(img * 2 + 1) * kmax * out_len;
  for (int i = threadIdx.x; i < band_rows * kmax; i += blockDim.x) {
    const int yl = i / kmax, k = i - yl * kmax;
    kvs[i] = __ldg(kv + static_cast<size_t>(k) * out_len + y_first + yl);
  }
  const int2 b_first = bv[y_first], b_last = bv[y_last];
  const int r0 = b_first.x;
  const int rows = b_last.x + b_last.y - r0;
  const int row_span = d.box_w * 3;
  if (rows > rmax || row_span + 15 > src_stride) __trap();  // host-side bounds violated: fail loudly
  const int row_bytes = out_w * 3;
  const uintptr_t first_row = reinterpret_cast<uintptr_t>(d.src) + static_cast<size_t>(d.box_y + r0) * d.pitch +
                              static_cast<size_t>(d.box_x) * 3;
  // ---- stage the crop rows: 16-byte aligned loads (an aligned 16-byte word that holds a valid byte never leaves
  // the allocation)
  const int cpr = src_stride >> 4;
  for (int i = threadIdx.x; i < rows * cpr; i += blockDim.x) {
    const int r = i / cpr, ch = i - r * cpr;
    const uintptr_t row = first_row + static_cast<size_t>(r) * d.pitch;
    const uintptr_t a0 = row & ~static_cast<uintptr_t>(15);
    if (a0 + 16 * ch < row + row_span)
      *reinterpret_cast<uint4*>(srcb + r * src_stride + 16 * ch) = __ldg(reinterpret_cast<const uint4*>(a0) + ch);
  }
  __syncthreads();
  // ---- horizontal pass: thread = output column
  double fs = __ddiv_rn(static_cast<double>(d.box_w), static_cast<double>(d.rs_w));
  if (fs < 1.0) fs = 1.0;
  const int ksize_h = static_cast<int>(ceil(__dmul_rn(2.0, fs))) * 2 + 1;  // upper bound of the taps of any column
  for (int c = threadIdx.x; c < out_w; c += blockDim.x) {
    const int2 b = bh[c];
    if (ksize_h <= 5) hpass_column<5>(srcb, tmp, kh, b, c, rows, src_stride, row_bytes, out_len, first_row, d.pitch);
    else if (ksize_h <= 7) hpass_column<7>(srcb, tmp, kh, b, c, rows, src_stride, row_bytes, out_len, first_row, d.pitch);
    else if (ksize_h <= 9) hpass_column<9>(srcb, tmp, kh, b, c, rows, src_stride, row_bytes, out_len, first_row, d.pitch);
    else if (ksize_h <= 11) hpass_column<11>(srcb, tmp, kh, b, c, rows, src_stride, row_bytes, out_len, first_row, d.pitch);
    else if (ksize_h <= 13) hpass_column<13>(srcb, tmp, kh, b, c, rows, src_stride, row_bytes, out_len, first_row, d.pitch);
    else {
      const uint8_t* col = srcb + b.x * 3;
      uintptr_t row_addr = first_row;
      for (int r = 0; r < rows; ++r, row_addr += d.pitch) {
        const uint8_t* p = col + r * src_stride + static_cast<int>(row_addr & 15);
        int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
        for (int k = 0; k < b.y; ++k) {
          const int w = __ldg(kh + static_cast<size_t>(k) * out_len + c);
          a0 += static_cast<int>(p[3 * k]) * w;
          a1 += static_cast<int>(p[3 * k + 1]) * w;
          a2 += static_cast<int>(p[3 * k + 2]) * w;
        }
        uint8_t* q = tmp + r * row_bytes + c * 3;
        q[0] = static_cast<uint8_t>(clip8(a0));
        q[1] = static_cast<uint8_t>(clip8(a1));
        q[2] = static_cast<uint8_t>(clip8(a2));
      }
    }
  }
  __syncthreads();
  // ---- vertical pass + table lookup + (flipped) store
  const size_t plane = static_cast<size_t>(out_h) * out_w;
  for (int yl = 0; yl < band_rows; ++yl) {
    const int y = y_first + yl;
    const int2 b = bv[y];
    const int* wv = kvs + yl * kmax;
    for (int c = threadIdx.x; c < out_w; c += blockDim.x) {
      const uint8_t* p = tmp + (b.x - r0) * row_bytes + c * 3;
      int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0;
#pragma unroll 4
      for (int k = 0; k < b.y; ++k) {
        const int w = wv[k];
        a0 += static_cast<int>(p[0]) * w;
        a1 += static_cast<int>(p[1]) * w;
        a2 += static_cast<int>(p[2]) * w;
        p += row_bytes;
      }
      const int oc = d.flip ? out_w - 1 - c : c;
      float* o = out + static_cast<size_t>(img) * 3 * plane + static_cast<size_t>(y) * out_w + oc;
      o[0] = lut[clip8(a0)];
      o[plane] = lut[256 + clip8(a1)];
      o[2 * plane] = lut[512 + clip8(a2)];
    }
  }
}

// ---------------------------------------------------------------------------------------- host side
static int ksize_of(int in_size, int out_size) {  // Resample.c: ksize = (int)ceil(support) * 2 + 1
  double fs = static_cast<double>(in_size) / static_cast<double>(out_size);
  if (fs < 1.0) fs = 1.0;
  return static_cast<int>(std::ceil(2.0 * fs)) * 2 + 1;
}

static const char* validate(const mudpt_image_desc* descs_host, int n, int out_h, int out_w, int* kmax_out, double* vscale_out,
                            int* max_box_w_out = nullptr) {
  if (!descs_host || n <= 0 || out_h <= 0 || out_w <= 0) return "augment: bad arguments";
  int kmax = 0, max_box_w = 0;
  double vscale = 1.0;
  for (int i = 0; i < n; ++i) {
    const mudpt_image_desc& d = descs_host[i];
    max_box_w = d.box_w > max_box_w ? d.box_w : max_box_w;
    if (!d.src || d.height <= 0 || d.width <= 0 || d.pitch < d.width * 3) return "augment: bad image descriptor";
    if (d.box_w <= 0 || d.box_h <= 0 || d.box_x < 0 || d.box_y < 0 || d.box_x + d.box_w > d.width || d.box_y + d.box_h > d.height)
      return "augment: crop box outside the image";
    if (d.rs_w <= 0 || d.rs_h <= 0 || d.win_x < 0 || d.win_y < 0 || d.win_x + out_w > d.rs_w || d.win_y + out_h > d.rs_h)
      return "augment: output window outside the resampled image";
    const int kh = ksize_of(d.box_w, d.rs_w), kv = ksize_of(d.box_h, d.rs_h);
    kmax = kh > kmax ? kh : kmax;
    kmax = kv > kmax ? kv : kmax;
    const double s = static_cast<double>(d.box_h) / static_cast<double>(d.rs_h);
    vscale = s > vscale ? s : vscale;
  }
  *kmax_out = kmax;
  *vscale_out = vscale;
  if (max_box_w_out) *max_box_w_out = max_box_w;
  return nullptr;
}

static size_t workspace_bytes(int n, int out_len, int kmax) {
  return static_cast<size_t>(n) * 2 * out_len * (sizeof(int2) + static_cast<size_t>(kmax) * sizeof(int));
}

long long augment_workspace_bytes(const mudpt_image_desc* descs_host, int n, int out_h, int out_w, const char** err) {
  int kmax = 0;
  double vs = 1.0;
  *err = validate(descs_host, n, out_h, out_w, &kmax, &vs);
  if (*err) return -1;
  return static_cast<long long>(workspace_bytes(n, out_h > out_w ? out_h : out_w, kmax));
}

template <int TR>
static const char* launch_augment(const mudpt_image_desc* descs, const int2* bounds, const int* kk, float* out, int n, int out_h,
                                  int out_w, int out_len, int kmax, int rmax, const float* mean, const float* sd, size_t smem,
                                  cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(augment_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return "augment: cudaFuncSetAttribute(max dynamic smem) failed";
    attr_done = true;
  }
  augment_kernel<TR><<<dim3((out_h + TR - 1) / TR, n), kAugThreads, smem, stream>>>(
      descs, bounds, kk, out, out_h, out_w, out_len, kmax, rmax, mean[0], mean[1], mean[2], sd[0], sd[1], sd[2]);
  count_launch();
  return launch_status("augment kernel launch failed");
}

template <int TR>
static const char* launch_staged(const mudpt_image_desc* descs, const int2* bounds, const int* kk, float* out, int n, int out_h,
                                 int out_w, int out_len, int kmax, int rmax, int src_stride, const float* mean, const float* sd,
                                 size_t smem, cudaStream_t stream) {
  static bool attr_done = false;
  if (!attr_done) {
    if (cudaFuncSetAttribute(augment_smem_kernel<TR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
      return "augment: cudaFuncSetAttribute(max dynamic smem) failed";
    attr_done = true;
  }
  augment_smem_kernel<TR><<<dim3((out_h + TR - 1) / TR, n), kAugThreads, smem, stream>>>(
      descs, bounds, kk, out, out_h, out_w, out_len, kmax, rmax, src_stride, mean[0], mean[1], mean[2], sd[0], sd[1], sd[2]);
  count_launch();
  return launch_status("augment kernel launch failed");
}

const char* augment_images(const mudpt_image_desc* descs, const mudpt_image_desc* descs_host, int n, int out_h, int out_w,
                           const float* mean_host, const float* std_host, void* workspace, long long ws_bytes, float* out,
                           cudaStream_t stream) {
  int kmax = 0, max_box_w = 0;
  double vscale = 1.0;
  const char* e = validate(descs_host, n, out_h, out_w, &kmax, &vscale, &max_box_w);
  if (e) return e;
  if (!descs || !mean_host || !std_host || !workspace || !out) return "augment: null argument";
  const int out_len = out_h > out_w ? out_h : out_w;
  if (ws_bytes < static_cast<long long>(workspace_bytes(n, out_len, kmax))) return "augment: workspace too small";
  int2* bounds = static_cast<int2*>(workspace);
  int* kk = reinterpret_cast<int*>(bounds + static_cast<size_t>(n) * 2 * out_len);
  resample_coeffs_kernel<<<n * 2, kAugThreads, 0, stream>>>(descs, bounds, kk, out_h, out_w, out_len, kmax);
  count_launch();
  if ((e = launch_status("resample_coeffs kernel launch failed"))) return e;
  // rows of the crop a band of TR output rows can touch: (TR - 1) * scale between the first and the last
  // centre, `support` = 2 * max(scale, 1) on either side, + 3 for the two roundings and the end-exclusive bound
  const double support = 2.0 * vscale;
  auto rows_for = [&](int tr) { return static_cast<int>((tr - 1) * vscale + 2.0 * support) + 3; };
  auto smem_for = [&](int tr) { return 3 * 256 * sizeof(float) + static_cast<size_t>(rows_for(tr)) * out_w * 3; };
  const size_t cap = 200 * 1024;
  // staged variant: the crop rows of a band in shared memory; 16-row bands (less overlap between the bands' row
  // ranges) when two or more CTAs still fit an SM, else 8-row bands
  const int src_stride = ((max_box_w * 3 + 30) / 16 + 1) * 16;  // >= any row's 16-byte aligned span
  auto staged_smem = [&](int tr) {
    return 3 * 256 * sizeof(float) + static_cast<size_t>(tr) * kmax * sizeof(int) +
           static_cast<size_t>(rows_for(tr)) * (src_stride + out_w * 3) + 64;  // + slack for zero-weight taps
  };
  if (staged_smem(16) <= 100 * 1024)
    return launch_staged<16>(descs, bounds, kk, out, n, out_h, out_w, out_len, kmax, rows_for(16), src_stride, mean_host, std_host, staged_smem(16), stream);
  if (staged_smem(8) <= 200 * 1024)
    return launch_staged<8>(descs, bounds, kk, out, n, out_h, out_w, out_len, kmax, rows_for(8), src_stride, mean_host, std_host, staged_smem(8), stream);
  if (smem_for(8) <= cap) return launch_augment<8>(descs, bounds, kk, out, n, out_h, out_w, out_len, kmax, rows_for(8), mean_host, std_host, smem_for(8), stream);
  if (smem_for(2) <= cap) return launch_augment<2>(descs, bounds, kk, out, n, out_h, out_w, out_len, kmax, rows_for(2), mean_host, std_host, smem_for(2), stream);
  if (smem_for(1) <= cap) return launch_augment<1>(descs, bounds, kk, out, n, out_h, out_w, out_len, kmax, rows_for(1), mean_host, std_host, smem_for(1), stream);
  return "augment: down-scaling factor too large for the shared-memory row buffer";
}

}  // namespace mudpt
