// Tower heads and the logits / cross-entropy head.
//
//   feature head : f[s] = LN(x[s, row(s), :]) @ proj        row = 0 (CLS, clip/model.py:548-551)
//                                                            or the EOT index (trainers/mudpt.py:150-154;
//                                                            LN is row-wise, so LN of the gathered row
//                                                            equals the gathered row of LN(x))
//   logits head  : L2-normalise both feature sets, logits = exp(logit_scale) * i_hat t_hat^T
//                  (trainers/mudpt.py:178-182), mean cross-entropy (:250) and its gradients
//                  d f_img, d f_txt.  Small problem (B x C x e = 32 x 1000 x 512): fp32 CUDA cores.
#include "head.h"

#include "common.cuh"
#include "launch_count.h"

namespace mudpt {

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < nw; ++i) t += red[i];
  return t;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = -INFINITY;
  for (int i = 0; i < nw; ++i) t = fmaxf(t, red[i]);
  return t;
}

// ------------------------------------------------------------------ feature head forward
// One block per sequence. proj is [d, e] row-major (x @ proj).
__global__ void __launch_bounds__(256) feature_head_fwd_kernel(const float* __restrict__ x, const int* __restrict__ rows,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const float* __restrict__ proj, float* __restrict__ f, int L,
                                                               int d, int e, float eps) {
  extern __shared__ float sh[];  // d floats (normalized row) + 32 reduce
  float* y = sh;
  float* red = sh + d;
  const int s = blockIdx.x;
  const int row = rows ? rows[s] : 0;
  const float* xr = x + (static_cast<size_t>(s) * L + row) * d;
  float sum = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) sum += xr[c];
  const float mean = block_sum(sum, red) / d;
  float sq = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { const float t = xr[c] - mean; sq += t * t; }
  const float rstd = rsqrtf(block_sum(sq, red) / d + eps);
  for (int c = threadIdx.x; c < d; c += blockDim.x) y[c] = (xr[c] - mean) * rstd * gamma[c] + beta[c];
  __syncthreads();
  for (int j = threadIdx.x; j < e; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < d; ++k) acc = fmaf(y[k], proj[static_cast<size_t>(k) * e + j], acc);
    f[static_cast<size_t>(s) * e + j] = acc;
  }
}

const char* feature_head_fwd(const float* x, const int* rows, const float* gamma, const float* beta, const float* proj,
                             float* f, int S, int L, int d, int e, float eps, cudaStream_t stream) {
  if (S <= 0) return nullptr;
  feature_head_fwd_kernel<<<S, 256, (d + 32) * sizeof(float), stream>>>(x, rows, gamma, beta, proj, f, L, d, e, eps);
  count_launch(1);
  return launch_status("feature head fwd launch failed");
}

// ------------------------------------------------------------------ feature head backward
// dx[s, row(s), :] = LN_bwd(df[s] @ proj^T) (fp32 + bf16 copy); all other rows of dx must have
// been zeroed by the caller.
__global__ void __launch_bounds__(256) feature_head_bwd_kernel(const float* __restrict__ df, const float* __restrict__ x,
                                                               const int* __restrict__ rows, const float* __restrict__ gamma,
                                                               const float* __restrict__ proj, float* __restrict__ dx,
                                                               bf16* __restrict__ dx_bf16, int L, int d, int e, float eps) {
  extern __shared__ float sh[];  // e (df) + d (g) + d (xhat) + 32
  float* sdf = sh;
  float* g = sh + e;
  float* xh = g + d;
  float* red = xh + d;
  const int s = blockIdx.x;
  const int row = rows ? rows[s] : 0;
  const size_t off = (static_cast<size_t>(s) * L + row) * d;
  for (int j = threadIdx.x; j < e; j += blockDim.x) sdf[j] = df[static_cast<size_t>(s) * e + j];
  float sum = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) sum += x[off + c];
  const float mean = block_sum(sum, red) / d;
  float sq = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { const float t = x[off + c] - mean; sq += t * t; }
  const float rstd = rsqrtf(block_sum(sq, red) / d + eps);
  // dy = df @ proj^T, one warp per output k (coalesced over e)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int k = warp; k < d; k += nw) {
    float acc = 0.f;
    for (int j = lane; j < e; j += 32) acc = fmaf(sdf[j], proj[static_cast<size_t>(k) * e + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      g[k] = acc * gamma[k];
      xh[k] = (x[off + k] - mean) * rstd;
    }
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int c = threadIdx.x; c < d; c += blockDim.x) { s1 += g[c]; s2 += g[c] * xh[c]; }
  s1 = block_sum(s1, red) / d;
  s2 = block_sum(s2, red) / d;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float v = rstd * (g[c] - s1 - xh[c] * s2);
    dx[off + c] = v;
    if (dx_bf16) dx_bf16[off + c] = __float2bfloat16(v);
  }
}

const char* feature_head_bwd(const float* df, const float* x, const int* rows, const float* gamma, const float* proj,
                             float* dx, bf16* dx_bf16, int S, int L, int d, int e, float eps, cudaStream_t stream) {
  if (S <= 0) return nullptr;
  feature_head_bwd_kernel<<<S, 256, (e + 2 * d + 32) * sizeof(float), stream>>>(df, x, rows, gamma, proj, dx, dx_bf16, L,
                                                                                d, e, eps);
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

// One block per image: logits row, log-softmax CE against the label, dlogits row
// (softmax - onehot) * grad_scale.  labels == nullptr -> logits only.
__global__ void __launch_bounds__(256) logits_ce_kernel(const float* __restrict__ in_hat, const float* __restrict__ tn_hat,
                                                        const long long* __restrict__ labels, float scale,
                                                        float* __restrict__ logits, float* __restrict__ loss_rows,
                                                        float* __restrict__ dlogits, int C, int e, float grad_scale) {
  extern __shared__ float sh[];  // e (image feature) + C (logits) + 32
  float* fi = sh;
  float* lg = sh + e;
  float* red = lg + C;
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < e; j += blockDim.x) fi[j] = in_hat[static_cast<size_t>(b) * e + j];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int c = warp; c < C; c += nw) {
    float acc = 0.f;
    for (int j = lane; j < e; j += 32) acc = fmaf(fi[j], tn_hat[static_cast<size_t>(c) * e + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      lg[c] = acc * scale;
      logits[static_cast<size_t>(b) * C + c] = acc * scale;
    }
  }
  __syncthreads();
  if (labels == nullptr) return;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, lg[c]);
  mx = block_max(mx, red);
  float se = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) se += __expf(lg[c] - mx);
  se = block_sum(se, red);
  const float lse = mx + __logf(se);
  const int y = static_cast<int>(labels[b]);
  if (threadIdx.x == 0) loss_rows[b] = lse - lg[y];
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float p = __expf(lg[c] - lse);
    dlogits[static_cast<size_t>(b) * C + c] = (p - (c == y ? 1.f : 0.f)) * grad_scale;
  }
}

// d f_a[r] = normalize_bwd( scale * sum_k dl(r,k) * b_hat[k] ), with dl indexed [r,k] (TRANS=0)
// or [k,r] (TRANS=1).  normalize_bwd(g) = (g - a_hat (a_hat . g)) * inv_norm.  One block per row r.
template <bool TRANS>
__global__ void __launch_bounds__(256) dfeat_kernel(const float* __restrict__ dl, const float* __restrict__ b_hat,
                                                    const float* __restrict__ a_hat, const float* __restrict__ a_inv,
                                                    float scale, float* __restrict__ da, int R, int Kn, int e) {
  extern __shared__ float sh[];  // Kn (dl row) + 32
  float* w = sh;
  float* red = sh + Kn;
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < Kn; k += blockDim.x)
    w[k] = TRANS ? dl[static_cast<size_t>(k) * R + r] : dl[static_cast<size_t>(r) * Kn + k];
  __syncthreads();
  // blockDim.x >= e is not required: loop over feature columns
  float dot_part = 0.f;
  for (int j = threadIdx.x; j < e; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < Kn; ++k) acc = fmaf(w[k], b_hat[static_cast<size_t>(k) * e + j], acc);
    acc *= scale;
    da[static_cast<size_t>(r) * e + j] = acc;  // staged, fixed up below
    dot_part += acc * a_hat[static_cast<size_t>(r) * e + j];
  }
  const float dot = block_sum(dot_part, red);
  const float inv = a_inv[r];
  for (int j = threadIdx.x; j < e; j += blockDim.x) {
    const size_t o = static_cast<size_t>(r) * e + j;
    da[o] = (da[o] - a_hat[o] * dot) * inv;
  }
}

__global__ void sum_scale_kernel(const float* __restrict__ v, float* __restrict__ out, int n, float scale) {
  __shared__ float red[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += v[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s * scale;
}

const char* logits_head(const float* f_img, const float* f_txt, const long long* labels, float scale, int B, int C, int e,
                        float inv_global_batch, float* ws, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                        cudaStream_t stream) {
  if (B <= 0 || C <= 0) return nullptr;
  // workspace layout (floats): in_hat[B*e] tn_hat[C*e] inv_i[B] inv_t[C] loss_rows[B] dlogits[B*C]
  float* in_hat = ws;
  float* tn_hat = in_hat + static_cast<size_t>(B) * e;
  float* inv_i = tn_hat + static_cast<size_t>(C) * e;
  float* inv_t = inv_i + B;
  float* loss_rows = inv_t + C;
  float* dlogits = loss_rows + B;
  l2norm_kernel<<<(B + 7) / 8, 256, 0, stream>>>(f_img, in_hat, inv_i, B, e);
  l2norm_kernel<<<(C + 7) / 8, 256, 0, stream>>>(f_txt, tn_hat, inv_t, C, e);
  const size_t sm = (static_cast<size_t>(e) + C + 32) * sizeof(float);
  if (sm > 200 * 1024) return "logits head: too many classes for the shared-memory row";
  if (sm > 48 * 1024) cudaFuncSetAttribute(logits_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sm));
  logits_ce_kernel<<<B, 256, sm, stream>>>(in_hat, tn_hat, labels, scale, logits, loss_rows, dlogits, C, e, inv_global_batch);
  if (labels != nullptr) {
    sum_scale_kernel<<<1, 256, 0, stream>>>(loss_rows, loss, B, inv_global_batch);
    if (d_f_img) {
      const size_t s1 = (static_cast<size_t>(C) + 32) * sizeof(float);
      if (s1 > 48 * 1024) cudaFuncSetAttribute(dfeat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s1));
      dfeat_kernel<false><<<B, 256, s1, stream>>>(dlogits, tn_hat, in_hat, inv_i, scale, d_f_img, B, C, e);
    }
    if (d_f_txt) {
      const size_t s2 = (static_cast<size_t>(B) + 32) * sizeof(float);
      if (s2 > 48 * 1024) return "logits head: batch too large";
      dfeat_kernel<true><<<C, 256, s2, stream>>>(dlogits, in_hat, tn_hat, inv_t, scale, d_f_txt, C, B, e);
    }
  }
  count_launch(labels != nullptr ? 4 + (d_f_img ? 1 : 0) + (d_f_txt ? 1 : 0) : 3);
  return launch_status("logits head launch failed");
}

size_t logits_head_workspace_floats(int B, int C, int e) {
  return static_cast<size_t>(B) * e + static_cast<size_t>(C) * e + B + C + B + static_cast<size_t>(B) * C;
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
  l2norm_kernel<<<(B + 7) / 8, 256, 0, stream>>>(f_img, in_hat, inv_i, B, e);
  l2norm_kernel<<<(C + 7) / 8, 256, 0, stream>>>(f_txt, tn_hat, inv_t, C, e);
  const size_t s1 = (static_cast<size_t>(C) + 32) * sizeof(float);
  if (s1 > 200 * 1024) return "logits head: too many classes";
  if (s1 > 48 * 1024) cudaFuncSetAttribute(dfeat_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(s1));
  dfeat_kernel<false><<<B, 256, s1, stream>>>(dlogits, tn_hat, in_hat, inv_i, scale, d_f_img, B, C, e);
  const size_t s2 = (static_cast<size_t>(B) + 32) * sizeof(float);
  if (s2 > 48 * 1024) return "logits head: batch too large";
  dfeat_kernel<true><<<C, 256, s2, stream>>>(dlogits, in_hat, tn_hat, inv_t, scale, d_f_txt, C, B, e);
  count_launch(4);
  return launch_status("logits head bwd launch failed");
}

}  // namespace mudpt
