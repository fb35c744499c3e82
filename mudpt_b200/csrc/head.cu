// Tower heads and the logits / cross-entropy head.
//
//   feature head : f[s] = LN(x[s, row(s), :]) @ proj        row = 0 (CLS, clip/model.py:548-551)
//                                                            or the EOT index (trainers/mudpt.py:150-154;
//                                                            LN is row-wise, so LN of the gathered row
//                                                            equals the gathered row of LN(x))
//   logits head  : L2-normalise both feature sets, logits = exp(logit_scale) * i_hat t_hat^T
//                  (trainers/mudpt.py:178-182), mean cross-entropy (:250) and its gradients
//                  d f_img, d f_txt.
// These are the last ~0.6 GFLOP of a 14 TFLOP step and they decide the logits directly, so they stay
// in fp32 on the CUDA cores: a small strided SGEMM (64x64 tiles) plus row kernels (one warp per row).
#include "head.h"

#include <cstdlib>

#include "common.cuh"
#include "launch_count.h"

namespace mudpt {

// ------------------------------------------------------------------ small fp32 GEMM
// C[m, n] = alpha * sum_k A(m, k) * B(k, n),  A(m,k) = A[m*sam + k*sak],  B(k,n) = B[k*sbk + n*sbn],
// C row-major [M, N].  64 x 64 tile per block, 256 threads, 4 x 4 outputs per thread, k-step 16.
// Split-K: the head GEMMs are tiny (32 x 512 x 768 ...) and sit on the critical path between the towers'
// forward and backward, where nothing else can run; 8-16 CTAs walking K in 16-wide slices were pure latency
// (33-44 us each).  With `partial` != nullptr, blockIdx.z owns K-slice [z * kchunk, (z + 1) * kchunk) and writes
// its unscaled tile to partial[z][M][N]; splitk_reduce_kernel adds the slices in order.  The slice width is a
// constant (never a function of M), so a row's result does not depend on how many rows the call has: class
// shards stay bit-identical to the unsharded tower.
__global__ void __launch_bounds__(256) sgemm_strided_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            float* __restrict__ C, int M, int N, int K, long sam, long sak,
                                                            long sbk, long sbn, float alpha, float* __restrict__ partial,
                                                            int kchunk) {
  const int k_begin = blockIdx.z * kchunk;
  const int k_end = min(K, k_begin + kchunk);
  __shared__ float As[16][64 + 4];
  __shared__ float Bs[16][64 + 4];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  // the next k-slice is fetched into registers while the current one is multiplied (the plain
  // load -> sync -> multiply -> sync loop exposed one global-load latency per 16-wide slice: 62 us for the
  // 1000 x 512 x 512 text head)
  float ra[4], rb[4];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;  // 1024 elements of each tile
      // walk the unit-stride dimension with consecutive threads
      int am, ak, bk, bn;
      if (sak == 1) { ak = idx & 15; am = idx >> 4; } else { am = idx & 63; ak = idx >> 6; }
      if (sbn == 1) { bn = idx & 63; bk = idx >> 6; } else { bk = idx & 15; bn = idx >> 4; }
      const int gm = m0 + am, gk = k0 + ak;
      ra[i] = (gm < M && gk < k_end) ? A[gm * sam + gk * sak] : 0.f;
      const int gn = n0 + bn, gk2 = k0 + bk;
      rb[i] = (gn < N && gk2 < k_end) ? B[gk2 * sbk + gn * sbn] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = threadIdx.x + i * 256;
      int am, ak, bk, bn;
      if (sak == 1) { ak = idx & 15; am = idx >> 4; } else { am = idx & 63; ak = idx >> 6; }
      if (sbn == 1) { bn = idx & 63; bk = idx >> 6; } else { bk = idx & 15; bn = idx >> 4; }
      As[ak][am] = ra[i];
      Bs[bk][bn] = rb[i];
    }
  };
  fetch(k_begin);
  for (int k0 = k_begin; k0 < k_end; k0 += 16) {
    stash();
    __syncthreads();
    if (k0 + 16 < k_end) fetch(k0 + 16);
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      if (partial != nullptr) partial[(static_cast<size_t>(blockIdx.z) * M + gm) * N + gn] = acc[i][j];
      else C[static_cast<size_t>(gm) * N + gn] = alpha * acc[i][j];
    }
  }
}

// C[i] = alpha * (((p[0][i] + p[1][i]) + p[2][i]) + ...): fixed order, deterministic
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const float* __restrict__ partial, float* __restrict__ C, size_t mn,
                                                            int slices, float alpha) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= mn) return;
  float s = partial[i];
  for (int z = 1; z < slices; ++z) s += partial[static_cast<size_t>(z) * mn + i];
  C[i] = alpha * s;
}

static constexpr int kSplitChunk = 64;  // K-slice of one CTA (4 k-steps)
static int splitk_slices(int K) { return K >= 2 * kSplitChunk ? (K + kSplitChunk - 1) / kSplitChunk : 1; }
static size_t sgemm_scratch_floats(int M, int N, int K) {
  const int z = splitk_slices(K);
  return z > 1 ? static_cast<size_t>(z) * M * N : 0;
}

// scratch: sgemm_scratch_floats(M, N, K) floats (may be nullptr when that is 0)
static void sgemm(const float* A, const float* B, float* C, int M, int N, int K, long sam, long sak, long sbk, long sbn,
                  float alpha, float* scratch, cudaStream_t st) {
  const int z = scratch != nullptr ? splitk_slices(K) : 1;
  const dim3 grid((N + 63) / 64, (M + 63) / 64, z);
  if (z == 1) {
    sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, alpha, nullptr, K);
    count_launch(1);
    return;
  }
  sgemm_strided_kernel<<<grid, 256, 0, st>>>(A, B, C, M, N, K, sam, sak, sbk, sbn, alpha, scratch, kSplitChunk);
  const size_t mn = static_cast<size_t>(M) * N;
  splitk_reduce_kernel<<<static_cast<unsigned>((mn + 255) / 256), 256, 0, st>>>(scratch, C, mn, z, alpha);
  count_launch(2);
}

// ------------------------------------------------------------------ feature head
// y[s, :] = LN(x[s, row(s), :])   (one warp per sequence, fp32)
__global__ void __launch_bounds__(256) gather_ln_kernel(const float* __restrict__ x, const int* __restrict__ rows,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        float* __restrict__ y, int S, int L, int d, float eps) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  const float* xr = x + (static_cast<size_t>(s) * L + (rows ? rows[s] : 0)) * d;
  float sum = 0.f;
  for (int c = lane; c < d; c += 32) sum += xr[c];
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; sq += t * t; }
  const float rstd = rsqrtf(warp_sum(sq) / d + eps);
  for (int c = lane; c < d; c += 32) y[static_cast<size_t>(s) * d + c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
}

// dx[s, row(s), :] = LN_bwd(g[s, :]) with statistics recomputed from x (fp32 + bf16 copy); every other
// row of dx must have been zeroed by the caller.
__global__ void __launch_bounds__(256) scatter_ln_bwd_kernel(const float* __restrict__ g, const float* __restrict__ x,
                                                             const int* __restrict__ rows, const float* __restrict__ gamma,
                                                             float* __restrict__ dx, bf16* __restrict__ dx_bf16, int S, int L,
                                                             int d, float eps) {
  const int s = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= S) return;
  const size_t off = (static_cast<size_t>(s) * L + (rows ? rows[s] : 0)) * d;
  const float* xr = x + off;
  const float* gr = g + static_cast<size_t>(s) * d;
  float sum = 0.f;
  for (int c = lane; c < d; c += 32) sum += xr[c];
  const float mean = warp_sum(sum) / d;
  float sq = 0.f;
  for (int c = lane; c < d; c += 32) { const float t = xr[c] - mean; sq += t * t; }
  const float rstd = rsqrtf(warp_sum(sq) / d + eps);
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < d; c += 32) {
    const float gg = gr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
    s1 += gg;
    s2 += gg * xh;
  }
  s1 = warp_sum(s1) / d;
  s2 = warp_sum(s2) / d;
  for (int c = lane; c < d; c += 32) {
    const float gg = gr[c] * gamma[c], xh = (xr[c] - mean) * rstd;
    const float v = rstd * (gg - s1 - xh * s2);
    dx[off + c] = v;
    if (dx_bf16) dx_bf16[off + c] = __float2bfloat16(v);
  }
}

size_t feature_head_workspace_floats(int S, int d, int e) {
  const size_t a = sgemm_scratch_floats(S, e, d), b = sgemm_scratch_floats(S, d, e);
  return static_cast<size_t>(S) * d + (a > b ? a : b);
}

// proj is [d, e] row-major (x @ proj).  ws: feature_head_workspace_floats(S, d, e) floats.
const char* feature_head_fwd(const float* x, const int* rows, const float* gamma, const float* beta, const float* proj,
                             float* f, float* ws, int S, int L, int d, int e, float eps, cudaStream_t stream) {
  if (S <= 0) return nullptr;
  gather_ln_kernel<<<(S + 7) / 8, 256, 0, stream>>>(x, rows, gamma, beta, ws, S, L, d, eps);
  count_launch(1);
  sgemm(ws, proj, f, S, e, d, d, 1, e, 1, 1.f, ws + static_cast<size_t>(S) * d, stream);  // f = y @ proj
  return launch_status("feature head fwd launch failed");
}

const char* feature_head_bwd(const float* df, const float* x, const int* rows, const float* gamma, const float* proj,
                             float* dx, bf16* dx_bf16, float* ws, int S, int L, int d, int e, float eps, cudaStream_t stream) {
  if (S <= 0) return nullptr;
  sgemm(df, proj, ws, S, d, e, e, 1, 1, e, 1.f, ws + static_cast<size_t>(S) * d, stream);  // g = df @ proj^T : B(k = j, n = c) = proj[c*e + j]
  scatter_ln_bwd_kernel<<<(S + 7) / 8, 256, 0, stream>>>(ws, x, rows, gamma, dx, dx_bf16, S, L, d, eps);
  count_launch(1);
  return launch_status("feature head bwd launch failed");
}

// ------------------------------------------------------------------ logits / CE head
// fn = f / ||f||, inv = 1 / ||f||.  One warp per row.
__global__ void l2norm_kernel(const float* __restrict__ f, float* __restrict__ fn, float* __restrict__ inv, int rows, int e) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float sq = 0.f;
  for (int j = lane; j < e; j += 32) { const float v = f[static_cast<size_t>(r) * e + j]; sq += v * v; }
  const float rn = rsqrtf(warp_sum(sq));
  for (int j = lane; j < e; j += 32) fn[static_cast<size_t>(r) * e + j] = f[static_cast<size_t>(r) * e + j] * rn;
  if (lane == 0) inv[r] = rn;
}

// One warp per image: log-softmax CE of its logits row against the label, dlogits row
// (softmax - onehot) * grad_scale.
__global__ void __launch_bounds__(256) ce_rows_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                      float* __restrict__ loss_rows, float* __restrict__ dlogits, int Bn,
                                                      int C, float grad_scale) {
  const int b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (b >= Bn) return;
  const float* lg = logits + static_cast<size_t>(b) * C;
  float mx = -INFINITY;
  for (int c = lane; c < C; c += 32) mx = fmaxf(mx, lg[c]);
  mx = warp_max(mx);
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(lg[c] - mx);
  se = warp_sum(se);
  const float lse = mx + logf(se);
  // a label outside [0, C) never indexes the row: its loss is NaN, which the trainer's finiteness check reports
  const long long yl = labels[b];
  const bool y_ok = yl >= 0 && yl < C;
  const int y = y_ok ? static_cast<int>(yl) : -1;
  if (lane == 0) loss_rows[b] = y_ok ? lse - lg[y] : __int_as_float(0x7fc00000);
  for (int c = lane; c < C; c += 32)
    dlogits[static_cast<size_t>(b) * C + c] = (expf(lg[c] - lse) - (c == y ? 1.f : 0.f)) * grad_scale;
}

// normalize_bwd: d f[r] = (g[r] - f_hat[r] (f_hat[r] . g[r])) * inv_norm[r], in place on g.  One warp per row.
__global__ void __launch_bounds__(256) normalize_bwd_kernel(float* __restrict__ g, const float* __restrict__ f_hat,
                                                            const float* __restrict__ inv, int rows, int e) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float dot = 0.f;
  for (int j = lane; j < e; j += 32) dot += g[static_cast<size_t>(r) * e + j] * f_hat[static_cast<size_t>(r) * e + j];
  dot = warp_sum(dot);
  const float iv = inv[r];
  for (int j = lane; j < e; j += 32) {
    const size_t o = static_cast<size_t>(r) * e + j;
    g[o] = (g[o] - f_hat[o] * dot) * iv;
  }
}

__global__ void sum_scale_kernel(const float* __restrict__ v, float* __restrict__ out, int n, float scale) {
  // single warp: deterministic order
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) s += v[i];
  s = warp_sum(s);
  if (threadIdx.x == 0) out[0] = s * scale;
}

static bool logits_head_fused(const float* f_img, const float* f_txt, const long long* labels, float scale, int B, int C, int e,
                              float inv_global_batch, float* ws, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                              cudaStream_t stream);

static size_t logits_head_fixed_floats(int B, int C, int e) {
  return static_cast<size_t>(B) * e + static_cast<size_t>(C) * e + B + C + B + static_cast<size_t>(B) * C;
}
size_t logits_head_workspace_floats(int B, int C, int e) {
  size_t sc = sgemm_scratch_floats(B, C, e);
  const size_t s2 = sgemm_scratch_floats(B, e, C), s3 = sgemm_scratch_floats(C, e, B);
  sc = s2 > sc ? s2 : sc;
  sc = s3 > sc ? s3 : sc;
  return logits_head_fixed_floats(B, C, e) + sc;
}

const char* logits_head(const float* f_img, const float* f_txt, const long long* labels, float scale, int B, int C, int e,
                        float inv_global_batch, float* ws, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                        cudaStream_t stream) {
  if (B <= 0 || C <= 0) return nullptr;
  if (logits_head_fused(f_img, f_txt, labels, scale, B, C, e, inv_global_batch, ws, logits, loss, d_f_img, d_f_txt, stream))
    return launch_status("fused logits head launch failed");
  // workspace layout (floats): in_hat[B*e] tn_hat[C*e] inv_i[B] inv_t[C] loss_rows[B] dlogits[B*C]
  float* in_hat = ws;
  float* tn_hat = in_hat + static_cast<size_t>(B) * e;
  float* inv_i = tn_hat + static_cast<size_t>(C) * e;
  float* inv_t = inv_i + B;
  float* loss_rows = inv_t + C;
  float* dlogits = loss_rows + B;
  float* scratch = ws + logits_head_fixed_floats(B, C, e);  // split-K partials of the three small GEMMs
  l2norm_kernel<<<(B + 7) / 8, 256, 0, stream>>>(f_img, in_hat, inv_i, B, e);
  l2norm_kernel<<<(C + 7) / 8, 256, 0, stream>>>(f_txt, tn_hat, inv_t, C, e);
  count_launch(2);
  sgemm(in_hat, tn_hat, logits, B, C, e, e, 1, 1, e, scale, scratch, stream);  // logits = scale * i_hat @ t_hat^T
  if (labels != nullptr) {
    ce_rows_kernel<<<(B + 7) / 8, 256, 0, stream>>>(logits, labels, loss_rows, dlogits, B, C, inv_global_batch);
    sum_scale_kernel<<<1, 32, 0, stream>>>(loss_rows, loss, B, inv_global_batch);
    count_launch(2);
    if (d_f_img) {
      sgemm(dlogits, tn_hat, d_f_img, B, e, C, C, 1, e, 1, scale, scratch, stream);  // scale * dl @ t_hat
      normalize_bwd_kernel<<<(B + 7) / 8, 256, 0, stream>>>(d_f_img, in_hat, inv_i, B, e);
      count_launch(1);
    }
    if (d_f_txt) {
      sgemm(dlogits, in_hat, d_f_txt, C, e, B, 1, C, e, 1, scale, scratch, stream);  // scale * dl^T @ i_hat : A(m = c, k = b) = dl[b*C + c]
      normalize_bwd_kernel<<<(C + 7) / 8, 256, 0, stream>>>(d_f_txt, tn_hat, inv_t, C, e);
      count_launch(1);
    }
  }
  return launch_status("logits head launch failed");
}

// ------------------------------------------------------------------ fused logits / CE head (training)
// L2-normalise, scaled cosine logits, cross-entropy and both feature gradients in TWO launches
// (trainers/mudpt.py:178-182, :250; the serial section between the towers' forward and backward):
//   head_rows_kernel   one thread-block CLUSTER of 8 CTAs per image row b; CTA r owns the classes [r C/8, (r+1) C/8).
//                      One pass over its class features gives the dots with i_hat_b AND their norms (t_hat is never
//                      materialised), the row's softmax statistics are combined through distributed shared memory,
//                      a second pass accumulates this slice's part of d i_hat_b, which rank 0 sums over the cluster
//                      (DSMEM again) and pushes through the normalisation backward.  Outputs: logits, dlogits,
//                      loss per row, 1 / ||f_txt||, d f_img.
//   head_cols_kernel   one warp per class: d t_hat_c = scale sum_b dlogits[b, c] i_hat_b, normalisation backward;
//                      block 0 also adds the per-row losses in a fixed order.
// Everything stays fp32 and every reduction has a fixed order: results are deterministic.
static constexpr int HEAD_CL = 8;       // cluster size (portable maximum)
static constexpr int HEAD_THREADS = 256;
static constexpr int HEAD_MAX_E = 1024;

struct HeadShared {
  float ihat[HEAD_MAX_E];      // normalised image feature of the row
  float part[HEAD_MAX_E];      // this CTA's part of d i_hat
  float red[32];
  float stat[2];               // (max, sum exp) of this CTA's logits slice
  float lse;
};

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (HEAD_THREADS >> 5); ++w) t += red[w];
  return t;
}

__global__ void __cluster_dims__(HEAD_CL, 1, 1) __launch_bounds__(HEAD_THREADS)
head_rows_kernel(const float* __restrict__ f_img, const float* __restrict__ f_txt, const long long* __restrict__ labels,
                 float scale, int C, int e, float grad_scale, float* __restrict__ logits, float* __restrict__ dlogits,
                 float* __restrict__ loss_rows, float* __restrict__ inv_t, float* __restrict__ d_f_img) {
  extern __shared__ float head_dyn[];  // [per] logits (then dlogits) of this CTA's class slice, [per] 1 / ||f_txt||
  __shared__ HeadShared sh;
  const int b = blockIdx.x / HEAD_CL;
  uint32_t rank;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int per = (C + HEAD_CL - 1) / HEAD_CL;
  const int c_lo = static_cast<int>(rank) * per, c_hi = min(C, c_lo + per);
  float* inv_s = head_dyn + per;
  // 1. i_hat_b
  float sq = 0.f;
  for (int j = threadIdx.x; j < e; j += HEAD_THREADS) {
    const float v = f_img[static_cast<size_t>(b) * e + j];
    sh.ihat[j] = v;
    sq += v * v;
  }
  const float inv_i = rsqrtf(block_sum(sq, sh.red));
  for (int j = threadIdx.x; j < e; j += HEAD_THREADS) sh.ihat[j] *= inv_i;
  __syncthreads();
  // 2. logits of the slice: one warp per class, dot and squared norm in the same pass
  for (int c = c_lo + warp; c < c_hi; c += (HEAD_THREADS >> 5)) {
    const float* tr = f_txt + static_cast<size_t>(c) * e;
    float dot = 0.f, nn = 0.f;
    for (int j = lane * 4; j < e; j += 128) {
      const float4 t4 = *reinterpret_cast<const float4*>(tr + j);
      const float4 i4 = *reinterpret_cast<const float4*>(sh.ihat + j);
      dot += t4.x * i4.x + t4.y * i4.y + t4.z * i4.z + t4.w * i4.w;
      nn += t4.x * t4.x + t4.y * t4.y + t4.z * t4.z + t4.w * t4.w;
    }
    dot = warp_sum(dot);
    nn = warp_sum(nn);
    if (lane == 0) {
      const float it = rsqrtf(nn);
      const float lg = scale * dot * it;
      head_dyn[c - c_lo] = lg;
      inv_s[c - c_lo] = it;
      logits[static_cast<size_t>(b) * C + c] = lg;
      if (b == 0 && inv_t != nullptr) inv_t[c] = it;
    }
  }
  __syncthreads();
  if (labels == nullptr) return;  // logits only (every CTA of the cluster takes this branch together)
  // 3. softmax statistics of the row: slice max / sum, combined over the cluster through DSMEM
  float mx = -INFINITY;
  for (int c = c_lo + threadIdx.x; c < c_hi; c += HEAD_THREADS) mx = fmaxf(mx, head_dyn[c - c_lo]);
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) sh.red[warp] = mx;
  __syncthreads();
  mx = -INFINITY;
  for (int w = 0; w < (HEAD_THREADS >> 5); ++w) mx = fmaxf(mx, sh.red[w]);
  float se = 0.f;
  for (int c = c_lo + threadIdx.x; c < c_hi; c += HEAD_THREADS) se += expf(head_dyn[c - c_lo] - mx);
  se = block_sum(se, sh.red);
  if (threadIdx.x == 0) { sh.stat[0] = mx; sh.stat[1] = se; }
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (threadIdx.x == 0) {
    float gm = -INFINITY, st[HEAD_CL][2];
    for (uint32_t r = 0; r < HEAD_CL; ++r) {
      uint32_t a;
      asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(sh.stat)), "r"(r));
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(st[r][0]) : "r"(a) : "memory");
      asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(st[r][1]) : "r"(a + 4) : "memory");
      gm = fmaxf(gm, st[r][0]);
    }
    float gs = 0.f;
    for (uint32_t r = 0; r < HEAD_CL; ++r) gs += st[r][1] > 0.f ? st[r][1] * expf(st[r][0] - gm) : 0.f;
    sh.lse = gm + logf(gs);
  }
  __syncthreads();
  const float lse = sh.lse;
  // 4. loss of the row (the CTA owning the label's class; a label outside [0, C) gives NaN -- see ce_rows_kernel)
  const long long yl = labels[b];
  const bool y_ok = yl >= 0 && yl < C;
  const int y = y_ok ? static_cast<int>(yl) : -1;
  if (threadIdx.x == 0) {
    if (y_ok && y >= c_lo && y < c_hi) loss_rows[b] = lse - head_dyn[y - c_lo];
    else if (!y_ok && rank == 0) loss_rows[b] = __int_as_float(0x7fc00000);
  }
  __syncthreads();  // (the label's logit is read before step 5 turns the slice into dlogits)
  // 5. dlogits of the slice (kept in shared memory as the weights of the second pass)
  for (int c = c_lo + threadIdx.x; c < c_hi; c += HEAD_THREADS) {
    const float dl = (expf(head_dyn[c - c_lo] - lse) - (c == y ? 1.f : 0.f)) * grad_scale;
    dlogits[static_cast<size_t>(b) * C + c] = dl;
    head_dyn[c - c_lo] = dl;
  }
  __syncthreads();
  if (d_f_img == nullptr) {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    return;
  }
  // 6. this slice's part of d i_hat_b = scale sum_c dl_c t_hat_c  (thread j owns columns j, j + 256, ...)
  float acc[HEAD_MAX_E / HEAD_THREADS];
#pragma unroll
  for (int k = 0; k < HEAD_MAX_E / HEAD_THREADS; ++k) acc[k] = 0.f;
  for (int c = c_lo; c < c_hi; ++c) {
    const float* tr = f_txt + static_cast<size_t>(c) * e;
    const float w = head_dyn[c - c_lo] * inv_s[c - c_lo];  // dl_c / ||t_c||
#pragma unroll
    for (int k = 0; k < HEAD_MAX_E / HEAD_THREADS; ++k) {
      const int j = threadIdx.x + k * HEAD_THREADS;
      if (j < e) acc[k] += w * tr[j];
    }
  }
#pragma unroll
  for (int k = 0; k < HEAD_MAX_E / HEAD_THREADS; ++k) {
    const int j = threadIdx.x + k * HEAD_THREADS;
    if (j < e) sh.part[j] = acc[k] * scale;
  }
  // 7. rank 0 adds the eight parts (fixed order) and applies the normalisation backward of the image feature
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (rank == 0) {
    float g[HEAD_MAX_E / HEAD_THREADS], dot = 0.f;
#pragma unroll
    for (int k = 0; k < HEAD_MAX_E / HEAD_THREADS; ++k) {
      const int j = threadIdx.x + k * HEAD_THREADS;
      g[k] = 0.f;
      if (j < e) {
        for (uint32_t r = 0; r < HEAD_CL; ++r) {
          uint32_t a;
          float v;
          asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(sh.part + j)), "r"(r));
          asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory");
          g[k] += v;
        }
        dot += g[k] * sh.ihat[j];
      }
    }
    dot = block_sum(dot, sh.red);
#pragma unroll
    for (int k = 0; k < HEAD_MAX_E / HEAD_THREADS; ++k) {
      const int j = threadIdx.x + k * HEAD_THREADS;
      if (j < e) d_f_img[static_cast<size_t>(b) * e + j] = (g[k] - sh.ihat[j] * dot) * inv_i;
    }
  }
  // nobody leaves while rank 0 still reads its shared memory
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// One warp per class c: d f_txt[c] = normalisation backward of scale * sum_b dlogits[b, c] i_hat_b.
// Shared memory: i_hat [B, e] (B <= 64 images at e = 512 fit 128 KB; larger batches read f_img rows from L2).
__global__ void __launch_bounds__(HEAD_THREADS) head_cols_kernel(const float* __restrict__ f_img, const float* __restrict__ f_txt,
                                                               const float* __restrict__ dlogits, const float* __restrict__ inv_t,
                                                               const float* __restrict__ loss_rows, float scale, int Bn, int C,
                                                               int e, int b_smem, float loss_scale, float* __restrict__ loss,
                                                               float* __restrict__ d_f_txt) {
  extern __shared__ float cols_dyn[];  // [b_smem][e] normalised image features + [Bn] 1 / ||f_img||
  float* inv_i = cols_dyn + static_cast<size_t>(b_smem) * e;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int b = warp; b < Bn; b += (HEAD_THREADS >> 5)) {
    float sq = 0.f;
    for (int j = lane; j < e; j += 32) { const float v = f_img[static_cast<size_t>(b) * e + j]; sq += v * v; }
    sq = warp_sum(sq);
    if (lane == 0) inv_i[b] = rsqrtf(sq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < b_smem * e; i += HEAD_THREADS) cols_dyn[i] = f_img[i] * inv_i[i / e];
  __syncthreads();
  if (blockIdx.x == 0 && warp == 0 && loss != nullptr) {  // mean loss: fixed-order sum of the rows
    float sacc = 0.f;
    for (int b = lane; b < Bn; b += 32) sacc += loss_rows[b];
    sacc = warp_sum(sacc);
    if (lane == 0) loss[0] = sacc * loss_scale;
  }
  const int c = blockIdx.x * (HEAD_THREADS >> 5) + warp;
  if (c >= C) return;
  const float it = inv_t[c];
  const float* tr = f_txt + static_cast<size_t>(c) * e;
  float dot = 0.f;
  for (int j0 = lane * 4; j0 < e; j0 += 128) {
    float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int b = 0; b < Bn; ++b) {
      const float dl = dlogits[static_cast<size_t>(b) * C + c];
      float4 i4;
      if (b < b_smem) i4 = *reinterpret_cast<const float4*>(cols_dyn + static_cast<size_t>(b) * e + j0);
      else {
        i4 = *reinterpret_cast<const float4*>(f_img + static_cast<size_t>(b) * e + j0);
        const float s = inv_i[b];
        i4.x *= s; i4.y *= s; i4.z *= s; i4.w *= s;
      }
      g.x += dl * i4.x; g.y += dl * i4.y; g.z += dl * i4.z; g.w += dl * i4.w;
    }
    g.x *= scale; g.y *= scale; g.z *= scale; g.w *= scale;
    const float4 t4 = *reinterpret_cast<const float4*>(tr + j0);
    dot += (g.x * t4.x + g.y * t4.y + g.z * t4.z + g.w * t4.w) * it;  // g . t_hat
    *reinterpret_cast<float4*>(d_f_txt + static_cast<size_t>(c) * e + j0) = g;  // (finished below)
  }
  dot = warp_sum(dot);
  for (int j0 = lane * 4; j0 < e; j0 += 128) {
    float4 g = *reinterpret_cast<const float4*>(d_f_txt + static_cast<size_t>(c) * e + j0);
    const float4 t4 = *reinterpret_cast<const float4*>(tr + j0);
    g.x = (g.x - t4.x * it * dot) * it; g.y = (g.y - t4.y * it * dot) * it;
    g.z = (g.z - t4.z * it * dot) * it; g.w = (g.w - t4.w * it * dot) * it;
    *reinterpret_cast<float4*>(d_f_txt + static_cast<size_t>(c) * e + j0) = g;
  }
}

static bool fused_head_enabled() {
  static int en = -1;
  if (en < 0) {
    const char* v = getenv("MUDPT_FUSED_HEAD");
    en = v ? (atoi(v) != 0) : 1;
  }
  return en != 0;
}

// Fused path of logits_head (same arguments); returns false when the shape is outside what the kernels take.
static bool logits_head_fused(const float* f_img, const float* f_txt, const long long* labels, float scale, int B, int C, int e,
                              float inv_global_batch, float* ws, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                              cudaStream_t stream) {
  if (!fused_head_enabled() || e % 4 != 0 || e > HEAD_MAX_E) return false;
  const int per = (C + HEAD_CL - 1) / HEAD_CL;
  const size_t dyn_rows = 2 * static_cast<size_t>(per) * sizeof(float);
  if (dyn_rows > 64 * 1024) return false;
  // workspace (floats): inv_t[C] loss_rows[B] dlogits[B*C]   (laid out inside the fixed part of logits_head's workspace)
  float* inv_t = ws;
  float* loss_rows = inv_t + C;
  float* dlogits = loss_rows + B;
  static bool attr_done = false;
  if (!attr_done) {
    cudaFuncSetAttribute(head_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(head_cols_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    attr_done = true;
  }
  head_rows_kernel<<<B * HEAD_CL, HEAD_THREADS, dyn_rows, stream>>>(f_img, f_txt, labels, scale, C, e, inv_global_batch, logits, dlogits,
                                                                  loss_rows, inv_t, labels ? d_f_img : nullptr);
  count_launch(1);
  if (labels != nullptr) {
    int b_smem = B;
    while (b_smem > 0 && (static_cast<size_t>(b_smem) * e + B) * sizeof(float) > 150 * 1024) --b_smem;
    const size_t dyn_cols = (static_cast<size_t>(b_smem) * e + B) * sizeof(float);
    if (d_f_txt != nullptr) {
      head_cols_kernel<<<(C + 7) / 8, HEAD_THREADS, dyn_cols, stream>>>(f_img, f_txt, dlogits, inv_t, loss_rows, scale, B, C, e, b_smem,
                                                                       inv_global_batch, loss, d_f_txt);
      count_launch(1);
    } else {
      sum_scale_kernel<<<1, 32, 0, stream>>>(loss_rows, loss, B, inv_global_batch);
      count_launch(1);
    }
  }
  return true;
}

// Backward of the logits alone for the module-level autograd path (CustomCLIP.forward returns
// logits; the caller's loss supplies dlogits [B, C]).
const char* logits_head_bwd(const float* f_img, const float* f_txt, const float* dlogits, float scale, int B, int C, int e,
                            float* ws, float* d_f_img, float* d_f_txt, cudaStream_t stream) {
  if (B <= 0 || C <= 0) return nullptr;
  float* in_hat = ws;
  float* tn_hat = in_hat + static_cast<size_t>(B) * e;
  float* inv_i = tn_hat + static_cast<size_t>(C) * e;
  float* inv_t = inv_i + B;
  float* scratch = ws + logits_head_fixed_floats(B, C, e);
  l2norm_kernel<<<(B + 7) / 8, 256, 0, stream>>>(f_img, in_hat, inv_i, B, e);
  l2norm_kernel<<<(C + 7) / 8, 256, 0, stream>>>(f_txt, tn_hat, inv_t, C, e);
  count_launch(2);
  sgemm(dlogits, tn_hat, d_f_img, B, e, C, C, 1, e, 1, scale, scratch, stream);
  normalize_bwd_kernel<<<(B + 7) / 8, 256, 0, stream>>>(d_f_img, in_hat, inv_i, B, e);
  sgemm(dlogits, in_hat, d_f_txt, C, e, B, 1, C, e, 1, scale, scratch, stream);
  normalize_bwd_kernel<<<(C + 7) / 8, 256, 0, stream>>>(d_f_txt, tn_hat, inv_t, C, e);
  count_launch(2);
  return launch_status("logits head bwd launch failed");
}


// ------------------------------------------------------------------ fused SGD step
struct SgdTable {
  float* p[SGD_MAX_TENSORS];
  const float* g[SGD_MAX_TENSORS];
  float* b[SGD_MAX_TENSORS];
  long long n[SGD_MAX_TENSORS];
  int count;
};

__global__ void __launch_bounds__(256) sgd_step_kernel(const SgdTable t, float lr, float momentum, float dampening, float wd,
                                                       int nesterov, int first_step) {
  for (int k = blockIdx.y; k < t.count; k += gridDim.y) {
    float* __restrict__ p = t.p[k];
    const float* __restrict__ g = t.g[k];
    float* __restrict__ b = t.b[k];
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < t.n[k];
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const float pv = p[i];
      float d = g[i];
      if (wd != 0.f) d = d + wd * pv;  // torch: d_p.add(p, alpha=wd)
      if (momentum != 0.f) {
        float bv;
        if (first_step) bv = d;  // buf = clone(d_p)
        else bv = b[i] * momentum + (1.f - dampening) * d;  // buf.mul_(momentum).add_(d_p, alpha=1-dampening)
        b[i] = bv;
        d = nesterov ? d + momentum * bv : bv;
      }
      p[i] = pv + (-lr) * d;  // p.add_(d_p, alpha=-lr)
    }
  }
}

const char* sgd_step(void* const* params, const void* const* grads, void* const* bufs, const long long* numel, int n, float lr,
                     float momentum, float dampening, float weight_decay, bool nesterov, bool first_step, cudaStream_t stream) {
  if (n <= 0) return nullptr;
  if (n > SGD_MAX_TENSORS) return "sgd_step: too many tensors";
  SgdTable t;
  long long mx = 0;
  for (int i = 0; i < n; ++i) {
    if (!params[i] || !grads[i] || (momentum != 0.f && !bufs[i])) return "sgd_step: null tensor";
    t.p[i] = static_cast<float*>(params[i]);
    t.g[i] = static_cast<const float*>(grads[i]);
    t.b[i] = static_cast<float*>(bufs ? bufs[i] : nullptr);
    t.n[i] = numel[i];
    if (numel[i] > mx) mx = numel[i];
  }
  t.count = n;
  long long bx = (mx + 255) / 256;
  if (bx > 148 * 4) bx = 148 * 4;
  if (bx < 1) bx = 1;
  sgd_step_kernel<<<dim3(static_cast<unsigned>(bx), n), 256, 0, stream>>>(t, lr, momentum, dampening, weight_decay, nesterov ? 1 : 0,
                                                                         first_step ? 1 : 0);
  count_launch(1);
  return launch_status("sgd step launch failed");
}

}  // namespace mudpt
