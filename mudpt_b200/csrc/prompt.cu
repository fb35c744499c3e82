// The trainable prompt algebra of MuDPT as four launches (two forward, two backward) instead of ~48 framework ops:
//   visual_prompts = deep_projections(deep_prompts)            trainers/mudpt.py:127   (text -> vision, deep)
//   shared_ctx     = embed_projection(ctx)                     trainers/mudpt.py:128   (text -> vision, shallow)
//   P_v[0] = ln_pre(visual_ctx + shared_ctx)                   clip/model.py:534-541
//   P_v[i] = visual_prompts[i-1] + visual_ctx_deep_prompts[i-1]                 :537
//   v2t    = visual_ctx_deep_projections(visual_ctx_deep_prompts)               :539   (vision -> text)
//   P_t[0] = ctx + positional_embedding[1:1+n]                 trainers/mudpt.py:143
//   P_t[i] = deep_prompts[i-1] + v2t[i-1]                                       :175
// and its backward into the 10 trainable tensors.  Everything is fp32 on the CUDA cores: a few dozen prompt rows
// against three [768 x 512]-sized weights (~40 MFLOP per direction), latency-bound, at the very start and end of the
// step.  All reductions have a fixed order (deterministic).
#include "prompt.h"

#include "common.cuh"
#include "launch_count.h"

namespace mudpt {

static constexpr int PR_ROWS = 16;      // prompt rows per pass (accumulators per thread)
static constexpr int PR_THREADS = 256;

// ------------------------------------------------------------------ forward: three Linear layers
// One warp per output feature i of one of the three layers; the input rows sit in shared memory.
//   job 0: Y = deep W_d^T + b_d     -> P_v[n + r] = Y + vdeep[r]
//   job 1: sh = ctx W_e^T + b_e     -> ln_in[r] = vctx[r] + sh          (LayerNorm'd by prompt_assemble_kernel)
//   job 2: v2t = vdeep W_v^T + b_v  -> P_t[n + r] = deep[r] + v2t
__global__ void __launch_bounds__(PR_THREADS) prompt_linear_fwd_kernel(PromptArgs a) {
  extern __shared__ float xs[];  // [PR_ROWS][K]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb0 = (a.dv + 7) / 8, nb1 = nb0;
  int job, blk = blockIdx.x;
  if (blk < nb0) job = 0; else if (blk < nb0 + nb1) { job = 1; blk -= nb0; } else { job = 2; blk -= nb0 + nb1; }
  const int R = job == 1 ? a.n : (a.depth - 1) * a.n;
  const int K = job == 2 ? a.dv : a.dt, N = job == 2 ? a.dt : a.dv;
  const float* X = job == 0 ? a.deep : job == 1 ? a.ctx : a.vdeep;
  const float* W = job == 0 ? a.Wd : job == 1 ? a.We : a.Wv;
  const float* bias = job == 0 ? a.bd : job == 1 ? a.be : a.bv;
  const int i = blk * 8 + warp;
  for (int r0 = 0; r0 < R; r0 += PR_ROWS) {
    const int nr = min(PR_ROWS, R - r0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < PR_ROWS * K; idx += PR_THREADS) xs[idx] = idx < nr * K ? X[static_cast<size_t>(r0) * K + idx] : 0.f;
    __syncthreads();
    if (i >= N) continue;
    float acc[PR_ROWS];
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r) acc[r] = 0.f;
    for (int k = lane; k < K; k += 32) {
      const float w = W[static_cast<size_t>(i) * K + k];
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) acc[r] = fmaf(w, xs[r * K + k], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < PR_ROWS; ++r) acc[r] = warp_sum(acc[r]);
    if (lane < nr) {
      float v = 0.f;
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) v = lane == r ? acc[r] : v;
      v += bias[i];
      const int row = r0 + lane;
      if (job == 0) a.P_v[static_cast<size_t>(a.n + row) * a.dv + i] = v + a.vdeep[static_cast<size_t>(row) * a.dv + i];
      else if (job == 1) a.ln_in[static_cast<size_t>(row) * a.dv + i] = v + a.vctx[static_cast<size_t>(row) * a.dv + i];
      else a.P_t[static_cast<size_t>(a.n + row) * a.dt + i] = v + a.deep[static_cast<size_t>(row) * a.dt + i];
    }
  }
}

// P_v[r] = ln_pre(ln_in[r]),  P_t[r] = ctx[r] + pos[r]   (one block per shallow prompt row)
__global__ void __launch_bounds__(PR_THREADS) prompt_assemble_kernel(PromptArgs a) {
  __shared__ float red[32];
  const int r = blockIdx.x;
  const float* x = a.ln_in + static_cast<size_t>(r) * a.dv;
  float s = 0.f;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) s += x[j];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  float mean = 0.f;
  for (int w = 0; w < PR_THREADS / 32; ++w) mean += red[w];
  mean /= a.dv;
  __syncthreads();
  float q = 0.f;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) { const float t = x[j] - mean; q += t * t; }
  q = warp_sum(q);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = q;
  __syncthreads();
  float var = 0.f;
  for (int w = 0; w < PR_THREADS / 32; ++w) var += red[w];
  const float rstd = rsqrtf(var / a.dv + a.eps);
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) a.P_v[static_cast<size_t>(r) * a.dv + j] = (x[j] - mean) * rstd * a.ln_g[j] + a.ln_b[j];
  for (int j = threadIdx.x; j < a.dt; j += PR_THREADS) a.P_t[static_cast<size_t>(r) * a.dt + j] = a.ctx[static_cast<size_t>(r) * a.dt + j] + a.pos[static_cast<size_t>(r) * a.dt + j];
}

// ------------------------------------------------------------------ backward
// u[r] = LayerNorm backward of dP_v[r] at ln_in[r]  (= d visual_ctx[r] = d shared_ctx[r]); one block per row
__global__ void __launch_bounds__(PR_THREADS) prompt_ln_bwd_kernel(PromptArgs a) {
  __shared__ float red[3][32];
  const int r = blockIdx.x;
  const float* x = a.ln_in + static_cast<size_t>(r) * a.dv;
  const float* dy = a.dP_v + static_cast<size_t>(r) * a.dv;
  auto bsum = [&](float v, int slot) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) red[slot][threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < PR_THREADS / 32; ++w) t += red[slot][w];
    __syncthreads();
    return t;
  };
  float s = 0.f;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) s += x[j];
  const float mean = bsum(s, 0) / a.dv;
  float q = 0.f;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) { const float t = x[j] - mean; q += t * t; }
  const float rstd = rsqrtf(bsum(q, 0) / a.dv + a.eps);
  float s1 = 0.f, s2 = 0.f;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) {
    const float g = dy[j] * a.ln_g[j], xh = (x[j] - mean) * rstd;
    s1 += g;
    s2 += g * xh;
  }
  s1 = bsum(s1, 1) / a.dv;
  s2 = bsum(s2, 2) / a.dv;
  for (int j = threadIdx.x; j < a.dv; j += PR_THREADS) {
    const float g = dy[j] * a.ln_g[j], xh = (x[j] - mean) * rstd;
    const float u = rstd * (g - s1 - xh * s2);
    a.u[static_cast<size_t>(r) * a.dv + j] = u;
    a.d_vctx[static_cast<size_t>(r) * a.dv + j] = u;
  }
}

// Everything else of the backward in one launch; block ranges select the job:
//   A  input gradients: out[r, j] = base[r, j] + sum_i G[r, i] W[i, j]   (32 columns j per block, 8 warps split i)
//        A0 d deep  = dP_t[n+r] + dY W_d      A1 d ctx = dP_t[r] + u W_e      A2 d vdeep = dP_v[n+r] + dv2t W_v
//   B  weight gradients: dW[i, j] = sum_r G[r, i] X[r, j]               (16 rows i per block, thread per column j)
//        B0 dW_d = dY^T deep                  B1 dW_e = u^T ctx               B2 dW_v = dv2t^T vdeep
//   C  bias gradients: db[i] = sum_r G[r, i]                            (one block per layer)
// with dY = dP_v[n:], dv2t = dP_t[n:].
__global__ void __launch_bounds__(PR_THREADS) prompt_bwd_kernel(PromptArgs a) {
  extern __shared__ float sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Rd = (a.depth - 1) * a.n;
  const float* dY = a.dP_v + static_cast<size_t>(a.n) * a.dv;
  const float* dv2t = a.dP_t + static_cast<size_t>(a.n) * a.dt;
  const int nA0 = (a.dt + 31) / 32, nA1 = nA0, nA2 = (a.dv + 31) / 32;
  const int nB0 = (a.dv + 15) / 16, nB1 = nB0, nB2 = (a.dt + 15) / 16;
  int blk = blockIdx.x;
  if (blk < nA0 + nA1 + nA2) {
    int job;
    if (blk < nA0) job = 0; else if (blk < nA0 + nA1) { job = 1; blk -= nA0; } else { job = 2; blk -= nA0 + nA1; }
    const int R = job == 1 ? a.n : Rd;
    const int I = job == 2 ? a.dt : a.dv, J = job == 2 ? a.dv : a.dt;  // contraction length, output width
    const float* G = job == 0 ? dY : job == 1 ? a.u : dv2t;            // [R, I]
    const float* W = job == 0 ? a.Wd : job == 1 ? a.We : a.Wv;         // [I, J]
    const float* base = job == 0 ? dv2t : job == 1 ? a.dP_t : dY;      // [R, J]
    float* out = job == 0 ? a.d_deep : job == 1 ? a.d_ctx : a.d_vdeep;
    float* gs = sm;                              // [PR_ROWS][I]
    float* part = sm + PR_ROWS * I;              // [8][PR_ROWS][32]
    const int j = blk * 32 + lane;
    const int i_per = (I + 7) / 8, i_lo = warp * i_per, i_hi = min(I, i_lo + i_per);
    for (int r0 = 0; r0 < R; r0 += PR_ROWS) {
      const int nr = min(PR_ROWS, R - r0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < PR_ROWS * I; idx += PR_THREADS) gs[idx] = idx < nr * I ? G[static_cast<size_t>(r0) * I + idx] : 0.f;
      __syncthreads();
      float acc[PR_ROWS];
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) acc[r] = 0.f;
      if (j < J) {
        for (int i = i_lo; i < i_hi; ++i) {
          const float w = W[static_cast<size_t>(i) * J + j];
#pragma unroll
          for (int r = 0; r < PR_ROWS; ++r) acc[r] = fmaf(gs[r * I + i], w, acc[r]);
        }
      }
#pragma unroll
      for (int r = 0; r < PR_ROWS; ++r) part[(warp * PR_ROWS + r) * 32 + lane] = acc[r];
      __syncthreads();
      // warp w finishes rows 2w, 2w + 1: the eight i-slices are added in order
      for (int rr = 0; rr < 2; ++rr) {
        const int r = warp * 2 + rr;
        if (r < nr && j < J) {
          float v = base[static_cast<size_t>(r0 + r) * J + j];
          for (int w8 = 0; w8 < 8; ++w8) v += part[(w8 * PR_ROWS + r) * 32 + lane];
          out[static_cast<size_t>(r0 + r) * J + j] = v;
        }
      }
    }
    return;
  }
  blk -= nA0 + nA1 + nA2;
  if (blk < nB0 + nB1 + nB2) {
    int job;
    if (blk < nB0) job = 0; else if (blk < nB0 + nB1) { job = 1; blk -= nB0; } else { job = 2; blk -= nB0 + nB1; }
    const int R = job == 1 ? a.n : Rd;
    const int I = job == 2 ? a.dt : a.dv, J = job == 2 ? a.dv : a.dt;  // dW is [I, J]
    const float* G = job == 0 ? dY : job == 1 ? a.u : dv2t;            // [R, I]
    const float* X = job == 0 ? a.deep : job == 1 ? a.ctx : a.vdeep;   // [R, J]
    float* dW = job == 0 ? a.d_Wd : job == 1 ? a.d_We : a.d_Wv;
    float* gs = sm;  // [R][16]: the 16 rows i of this block
    const int i0 = blk * 16;
    for (int idx = threadIdx.x; idx < R * 16; idx += PR_THREADS) {
      const int r = idx >> 4, ii = idx & 15;
      gs[idx] = i0 + ii < I ? G[static_cast<size_t>(r) * I + i0 + ii] : 0.f;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < J; j += PR_THREADS) {
      float acc[16];
#pragma unroll
      for (int ii = 0; ii < 16; ++ii) acc[ii] = 0.f;
      for (int r = 0; r < R; ++r) {
        const float x = X[static_cast<size_t>(r) * J + j];
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) acc[ii] = fmaf(gs[r * 16 + ii], x, acc[ii]);
      }
#pragma unroll
      for (int ii = 0; ii < 16; ++ii)
        if (i0 + ii < I) dW[static_cast<size_t>(i0 + ii) * J + j] = acc[ii];
    }
    return;
  }
  blk -= nB0 + nB1 + nB2;
  {  // bias gradients
    const int job = blk;
    const int R = job == 1 ? a.n : Rd;
    const int I = job == 2 ? a.dt : a.dv;
    const float* G = job == 0 ? dY : job == 1 ? a.u : dv2t;
    float* db = job == 0 ? a.d_bd : job == 1 ? a.d_be : a.d_bv;
    for (int i = threadIdx.x; i < I; i += PR_THREADS) {
      float v = 0.f;
      for (int r = 0; r < R; ++r) v += G[static_cast<size_t>(r) * I + i];
      db[i] = v;
    }
  }
}

static const char* check(const PromptArgs& a) {
  if (a.n <= 0 || a.depth < 1 || a.dt <= 0 || a.dv <= 0) return "prompt algebra: bad geometry";
  if (a.dt > 1024 || a.dv > 1024) return "prompt algebra: widths above 1024 are not supported";
  return nullptr;
}

const char* prompt_forward(const PromptArgs& a, cudaStream_t stream) {
  if (const char* e = check(a)) return e;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(prompt_linear_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PR_ROWS * 1024 * 4);
    cudaFuncSetAttribute(prompt_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (PR_ROWS * 1024 + 8 * PR_ROWS * 32) * 4);
    attr = true;
  }
  const int K = a.dt > a.dv ? a.dt : a.dv;
  const int blocks = 2 * ((a.dv + 7) / 8) + (a.depth > 1 ? (a.dt + 7) / 8 : 0);
  prompt_linear_fwd_kernel<<<blocks, PR_THREADS, static_cast<size_t>(PR_ROWS) * K * 4, stream>>>(a);
  prompt_assemble_kernel<<<a.n, PR_THREADS, 0, stream>>>(a);
  count_launch(2);
  return launch_status("prompt forward launch failed");
}

const char* prompt_backward(const PromptArgs& a, cudaStream_t stream) {
  if (const char* e = check(a)) return e;
  prompt_ln_bwd_kernel<<<a.n, PR_THREADS, 0, stream>>>(a);
  const int K = a.dt > a.dv ? a.dt : a.dv;
  const int nA = 2 * ((a.dt + 31) / 32) + (a.dv + 31) / 32;
  const int nB = 2 * ((a.dv + 15) / 16) + (a.dt + 15) / 16;
  const size_t smem = (static_cast<size_t>(PR_ROWS) * K + 8 * PR_ROWS * 32) * 4;
  const size_t smem_b = static_cast<size_t>((a.depth - 1) * a.n > a.n ? (a.depth - 1) * a.n : a.n) * 16 * 4;
  prompt_bwd_kernel<<<nA + nB + 3, PR_THREADS, smem > smem_b ? smem : smem_b, stream>>>(a);
  count_launch(2);
  return launch_status("prompt backward launch failed");
}

}  // namespace mudpt
