// Flash-style multi-head self-attention, forward and dgrad, for the short fixed sequences of
// the CLIP towers: vision L = 197 + n_ctx (no mask), text L <= 77 (causal mask,
// clip/model.py:810-816), head width 64.  Replaces the scaled_dot_product_attention call
// inside nn.MultiheadAttention (clip/model.py:271-273) and its autograd backward.
//
// Whole K/V (resp. Q/dO) of one (sequence, head) fits in shared memory, so each CTA owns a
// block of 16*nwarps rows and every warp is the sole owner of 16 of them: no atomics and no
// cross-warp reductions anywhere.
//   forward : S = QK^T, online softmax, O = PV, LSE (log2 domain) saved
//   dQ pass : rows = queries.  P recomputed from LSE, dP = dO V^T, dS = P*(dP - D), dQ = dS K
//   dKV pass: rows = keys.     P^T, dV = P^T dO, dP^T = V dO^T, dS^T, dK = dS^T Q
// With L <= ~260 and head width 64 the kernel is bound by its Q/K/V/O traffic, not by MMA
// rate (~100 FLOP/B vision, ~38 text), so the products run on warp-level
// mma.sync (HMMA) tiles fed by ldmatrix from XOR-swizzled shared memory.  The backward passes
// work on 32-column chunks and re-read their A fragments from shared memory instead of pinning
// them in registers: that keeps them under ~128 registers so 4 CTAs fit per SM.
#include "attention.h"

#include <cstdlib>

#include "common.cuh"
#include "launch_count.h"

namespace mudpt {

static constexpr int DH = 64;         // head width (all CLIP ViT towers)
static constexpr int ROW_BYTES = 128; // 64 bf16

__device__ __forceinline__ uint32_t swz(int row, int chunk) {
  return static_cast<uint32_t>(row * ROW_BYTES + ((chunk ^ (row & 7)) << 4));
}

// Load `nrows` rows x 64 bf16 columns into a swizzled smem tile; global row = row_begin + r,
// rows >= row_limit are zero-filled.
__device__ __forceinline__ void load_tile(uint32_t tile, const bf16* g, int ld, int row_begin, int row_limit, int nrows) {
  const int c = threadIdx.x & 7, r0 = threadIdx.x >> 3, rstep = blockDim.x >> 3;
  if ((rstep & 7) == 0) {
    // CTAs of >= 64 threads: a thread's rows are 8k apart, so its swizzled 16 B slot is loop-invariant
    // and both addresses advance by constants (the load loop was ~15 % of all issued instructions)
    uint32_t dst = tile + static_cast<uint32_t>(r0 * ROW_BYTES + ((c ^ (r0 & 7)) << 4));
    const bf16* src = g + static_cast<size_t>(row_begin + r0) * ld + c * 8;
    const size_t sstep = static_cast<size_t>(rstep) * ld;
    for (int r = r0; r < nrows; r += rstep) {
      const bool ok = row_begin + r < row_limit;
      cp_async16(dst, ok ? src : g, ok);
      dst += rstep * ROW_BYTES;
      src += sstep;
    }
  } else {
    for (int r = r0; r < nrows; r += rstep) {
      const int gr = row_begin + r;
      const bool ok = gr < row_limit;
      cp_async16(tile + swz(r, c), g + static_cast<size_t>(ok ? gr : 0) * ld + c * 8, ok);
    }
  }
}

__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], uint32_t tile, int row0, int ks, int lane) {
  ldsm_x4(tile + swz(row0 + (lane & 15), ks * 2 + (lane >> 4)), a[0], a[1], a[2], a[3]);
}

// acc(16 x 16*NG) = A(16 rows of tileA at row0, 64 dh) * T[c0 .. c0+16*NG)^T -- contraction over the
// head dimension.  Only 16-column groups g in [g_lo, g_hi) are computed, the rest stay 0.
// k-step outermost: consecutive MMAs hit different accumulators (no back-to-back dependency).
template <int NG>
__device__ __forceinline__ void mma_rows_x_cols(float (&acc)[2 * NG][4], uint32_t tileA, int row0, uint32_t tile, int c0,
                                                int g_lo, int g_hi, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    load_a_frag(a, tileA, row0, ks, lane);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
      if (g >= g_lo && g < g_hi) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4(tile + swz(c0 + g * 16 + (lane & 7) + ((lane >> 4) << 3), ks * 2 + ((lane >> 3) & 1)), b0, b1, b2, b3);
        mma16816(acc[2 * g], a, b0, b1);
        mma16816(acc[2 * g + 1], a, b2, b3);
      }
    }
  }
}

// out(16 x 64dh) += P(16 x 16*NG) * T[c0 .. c0+16*NG)  -- contraction over the tile rows; P comes
// straight from accumulator registers (converted to bf16 A fragments).
template <int NG>
__device__ __forceinline__ void mma_p_x_tile(float (&out)[8][4], const float (&p)[2 * NG][4], uint32_t tile, int c0,
                                             int g_lo, int g_hi, int lane) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
    if (g >= g_lo && g < g_hi) {
      uint32_t a[4];
      a[0] = pack_bf16(p[2 * g][0], p[2 * g][1]);
      a[1] = pack_bf16(p[2 * g][2], p[2 * g][3]);
      a[2] = pack_bf16(p[2 * g + 1][0], p[2 * g + 1][1]);
      a[3] = pack_bf16(p[2 * g + 1][2], p[2 * g + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t b0, b1, b2, b3;
        ldsm_x4_t(tile + swz(c0 + g * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), dp * 2 + (lane >> 4)), b0, b1, b2, b3);
        mma16816(out[2 * dp], a, b0, b1);
        mma16816(out[2 * dp + 1], a, b2, b3);
      }
    }
  }
}

// Write a warp's 16 x 64 fp32 accumulator (scaled) as bf16 to global rows, staged through the
// warp's own (no longer needed) 16 smem rows so the global stores are 16 B and coalesced.
__device__ __forceinline__ void store_rows_bf16(const float (&acc)[8][4], float scale0, float scale1, uint8_t* smem_gen,
                                                uint32_t tile_off, int row0, bf16* g, int ld, int grow0, int row_limit,
                                                int lane) {
  __syncwarp();
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    const int r_lo = row0 + (lane >> 2), r_hi = r_lo + 8;
    const int within = (lane & 3) * 4;  // byte offset of this thread's column pair inside the 16 B chunk nt
    *reinterpret_cast<uint32_t*>(smem_gen + tile_off + swz(r_lo, nt) + within) = pack_bf16(acc[nt][0] * scale0, acc[nt][1] * scale0);
    *reinterpret_cast<uint32_t*>(smem_gen + tile_off + swz(r_hi, nt) + within) = pack_bf16(acc[nt][2] * scale1, acc[nt][3] * scale1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int r = idx >> 3, c = idx & 7;
    const int gr = grow0 + r;
    if (gr < row_limit) {
      uint4 v = *reinterpret_cast<const uint4*>(smem_gen + tile_off + swz(row0 + r, c));
      *reinterpret_cast<uint4*>(g + static_cast<size_t>(gr) * ld + c * 8) = v;
    }
  }
}

template <int N>
__device__ __forceinline__ void zero_acc(float (&a)[N][4]) {
#pragma unroll
  for (int i = 0; i < N; ++i) a[i][0] = a[i][1] = a[i][2] = a[i][3] = 0.f;
}

// ---------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(128) attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ o,
                                                       float* __restrict__ lse2, int L, int H, int d, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nwarps = blockDim.x >> 5, BQ = nwarps * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sh = blockIdx.y, s = sh / H, h = sh - s * H;
  const int q0 = blockIdx.x * BQ;
  const int kv_len = CAUSAL ? min(L, q0 + BQ) : L;
  const int kv_rows = (kv_len + 15) & ~15;
  const int Lp = (L + 15) & ~15;
  const uint32_t sQ = smem_u32(smem), sK = sQ + BQ * ROW_BYTES, sV = sK + Lp * ROW_BYTES;
  const int ld = 3 * d;
  const bf16* base = qkv + static_cast<size_t>(s) * L * ld + h * DH;
  load_tile(sQ, base, ld, q0, L, BQ);
  load_tile(sK, base + d, ld, 0, L, kv_rows);
  load_tile(sV, base + 2 * d, ld, 0, L, kv_rows);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  const int row0 = warp * 16;
  if (q0 + row0 >= L) return;  // whole warp out of range (no further block-wide syncs below)
  float oacc[8][4];
  zero_acc(oacc);
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};
  const int qrow[2] = {q0 + row0 + (lane >> 2), q0 + row0 + (lane >> 2) + 8};
  const int kv_warp = CAUSAL ? min(kv_len, q0 + row0 + 16) : kv_len;  // keys this warp can see

  for (int c0 = 0; c0 < kv_warp; c0 += 64) {
    const int g_hi = min(4, (kv_warp - c0 + 15) >> 4);
    float sacc[8][4];
    zero_acc(sacc);
    mma_rows_x_cols<4>(sacc, sQ, row0, sK, c0, 0, g_hi, lane);
    float mx[2] = {-INFINITY, -INFINITY};
    // the kernel is instruction-bound (ncu: 3 % of the issued instructions are MMAs), so the softmax
    // touches only the 8-column tiles that were computed (nt < 2*g_hi, warp-uniform) and skips the
    // mask arithmetic for chunks that lie entirely inside the valid / causal region
    const bool need_mask = (c0 + 64 > L) || (CAUSAL && c0 + 63 > q0 + row0);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * g_hi) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          float v = sacc[nt][e] * scale_log2e;
          if (need_mask) {
            const int col = c0 + nt * 8 + (lane & 3) * 2 + (e & 1);
            const bool ok = col < L && (!CAUSAL || col <= qrow[r]);
            v = ok ? v : -INFINITY;
          }
          sacc[nt][e] = v;
          mx[r] = fmaxf(mx[r], v);
        }
      }
    }
    float alpha[2], m_use[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_run[r], mx[r]);
      m_use[r] = (m_new == -INFINITY) ? 0.f : m_new;
      alpha[r] = exp2f(m_run[r] - m_use[r]);  // m_run = -inf -> 0
      m_run[r] = m_new;
      l_run[r] *= alpha[r];
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      if (nt < 2 * g_hi) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          const float p = exp2f(sacc[nt][e] - m_use[r]);
          sacc[nt][e] = p;
          l_run[r] += p;
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) oacc[nt][e] *= alpha[e >> 1];
    }
    mma_p_x_tile<4>(oacc, sacc, sV, c0, 0, g_hi, lane);
  }
  float inv[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
    inv[r] = l_run[r] > 0.f ? 1.f / l_run[r] : 0.f;
    if ((lane & 3) == 0 && qrow[r] < L)
      lse2[(static_cast<size_t>(s) * H + h) * L + qrow[r]] = m_run[r] + log2f(l_run[r]);
  }
  store_rows_bf16(oacc, inv[0], inv[1], smem, 0, row0, o + static_cast<size_t>(s) * L * d + h * DH, d, q0 + row0, L, lane);
}

// ---------------------------------------------------------------------------------------
// backward, dQ pass (rows = queries).  Also produces D = rowsum(dO * O) for the dKV pass.
// ---------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(128, 4) attn_bwd_dq_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ o,
                                                             const bf16* __restrict__ d_o, const float* __restrict__ lse2,
                                                             float* __restrict__ dsum, bf16* __restrict__ dqkv, int L,
                                                             int H, int d, float scale, float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nwarps = blockDim.x >> 5, BQ = nwarps * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sh = blockIdx.y, s = sh / H, h = sh - s * H;
  const int q0 = blockIdx.x * BQ;
  const int kv_len = CAUSAL ? min(L, q0 + BQ) : L;
  const int kv_rows = (kv_len + 15) & ~15;
  const int Lp = (L + 15) & ~15;
  const uint32_t sQ = smem_u32(smem), sdO = sQ + BQ * ROW_BYTES, sO = sdO + BQ * ROW_BYTES;
  const uint32_t sK = sO + BQ * ROW_BYTES, sV = sK + Lp * ROW_BYTES;
  const int ld = 3 * d;
  const size_t seq_row = static_cast<size_t>(s) * L;
  const bf16* base = qkv + seq_row * ld + h * DH;
  load_tile(sQ, base, ld, q0, L, BQ);
  load_tile(sdO, d_o + seq_row * d + h * DH, d, q0, L, BQ);
  load_tile(sO, o + seq_row * d + h * DH, d, q0, L, BQ);
  load_tile(sK, base + d, ld, 0, L, kv_rows);
  load_tile(sV, base + 2 * d, ld, 0, L, kv_rows);
  cp_async_commit();
  cp_async_wait_all();
  __syncthreads();

  const int row0 = warp * 16;
  if (q0 + row0 >= L) return;
  // D_i = sum_j dO_ij O_ij for the warp's 16 rows: lane -> (row = lane/2, half = lane%2)
  float dpart = 0.f;
  {
    const int r = row0 + (lane >> 1);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int chunk = (lane & 1) * 4 + c;
      uint4 a = *reinterpret_cast<const uint4*>(smem + (sdO - sQ) + swz(r, chunk));
      uint4 b = *reinterpret_cast<const uint4*>(smem + (sO - sQ) + swz(r, chunk));
      const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* pb = reinterpret_cast<const uint32_t*>(&b);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float2 x = unpack_bf16(pa[i]), y = unpack_bf16(pb[i]);
        dpart += x.x * y.x + x.y * y.y;
      }
    }
    dpart += __shfl_xor_sync(0xffffffffu, dpart, 1);
  }
  const int qrow[2] = {q0 + row0 + (lane >> 2), q0 + row0 + (lane >> 2) + 8};
  float Dr[2], lse[2];
  Dr[0] = __shfl_sync(0xffffffffu, dpart, (lane >> 2) * 2);
  Dr[1] = __shfl_sync(0xffffffffu, dpart, ((lane >> 2) + 8) * 2);
  const size_t stat_base = (static_cast<size_t>(s) * H + h) * L;
  if ((lane & 1) == 0) {
    const int r = q0 + row0 + (lane >> 1);
    if (r < L) dsum[stat_base + r] = dpart;
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) lse[r] = qrow[r] < L ? lse2[stat_base + qrow[r]] : 0.f;

  float dq[8][4];
  zero_acc(dq);
  const int kv_warp = CAUSAL ? min(kv_len, q0 + row0 + 16) : kv_len;
  for (int c0 = 0; c0 < kv_warp; c0 += 32) {
    const int g_hi = min(2, (kv_warp - c0 + 15) >> 4);
    float sacc[4][4], dp[4][4];
    zero_acc(sacc);
    zero_acc(dp);
    mma_rows_x_cols<2>(sacc, sQ, row0, sK, c0, 0, g_hi, lane);
    mma_rows_x_cols<2>(dp, sdO, row0, sV, c0, 0, g_hi, lane);
    // rows >= L of the last tile carry lse = 0 and finite garbage: they are never stored, so only the
    // column mask matters; chunks entirely inside the valid / causal region skip it
    const bool need_mask = (c0 + 32 > L) || (CAUSAL && c0 + 31 > q0 + row0);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt < 2 * g_hi) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int r = e >> 1;
          float p = exp2f(sacc[nt][e] * scale_log2e - lse[r]);
          if (need_mask) {
            const int col = c0 + nt * 8 + (lane & 3) * 2 + (e & 1);
            const bool ok = col < L && (!CAUSAL || col <= qrow[r]);
            p = ok ? p : 0.f;
          }
          sacc[nt][e] = p * (dp[nt][e] - Dr[r]);  // dS (unscaled)
        }
      }
    }
    mma_p_x_tile<2>(dq, sacc, sK, c0, 0, g_hi, lane);
  }
  store_rows_bf16(dq, scale, scale, smem, 0, row0, dqkv + seq_row * ld + h * DH, ld, q0 + row0, L, lane);
}

// ---------------------------------------------------------------------------------------
// backward, dK/dV pass (rows = keys)
// ---------------------------------------------------------------------------------------
template <bool CAUSAL>
__global__ void __launch_bounds__(128, 3) attn_bwd_dkv_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ d_o,
                                                              const float* __restrict__ lse2, const float* __restrict__ dsum,
                                                              bf16* __restrict__ dqkv, int L, int H, int d, float scale,
                                                              float scale_log2e) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int nwarps = blockDim.x >> 5, BKV = nwarps * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sh = blockIdx.y, s = sh / H, h = sh - s * H;
  const int k0 = blockIdx.x * BKV;
  const int Lp = (L + 15) & ~15;
  const uint32_t sK = smem_u32(smem), sV = sK + BKV * ROW_BYTES, sQ = sV + BKV * ROW_BYTES, sdO = sQ + Lp * ROW_BYTES;
  float* sLse = reinterpret_cast<float*>(smem + (2 * BKV + 2 * Lp) * ROW_BYTES);
  float* sD = sLse + Lp;
  const int ld = 3 * d;
  const size_t seq_row = static_cast<size_t>(s) * L;
  const bf16* base = qkv + seq_row * ld + h * DH;
  const int q_blk = CAUSAL ? (k0 & ~15) : 0;  // queries below the block's first key never see it
  load_tile(sK, base + d, ld, k0, L, BKV);
  load_tile(sV, base + 2 * d, ld, k0, L, BKV);
  load_tile(sQ + q_blk * ROW_BYTES, base, ld, q_blk, L, Lp - q_blk);
  load_tile(sdO + q_blk * ROW_BYTES, d_o + seq_row * d + h * DH, d, q_blk, L, Lp - q_blk);
  cp_async_commit();
  const size_t stat_base = (static_cast<size_t>(s) * H + h) * L;
  for (int i = threadIdx.x; i < Lp; i += blockDim.x) {
    sLse[i] = i < L ? lse2[stat_base + i] : 0.f;
    sD[i] = i < L ? dsum[stat_base + i] : 0.f;
  }
  cp_async_wait_all();
  __syncthreads();

  const int row0 = warp * 16;
  if (k0 + row0 >= L) return;
  float dk[8][4], dv[8][4];
  zero_acc(dk);
  zero_acc(dv);
  const int krow[2] = {k0 + row0 + (lane >> 2), k0 + row0 + (lane >> 2) + 8};
  const int q_first = CAUSAL ? (k0 + row0) : 0;  // first query that can see any of this warp's keys
  for (int c0 = q_first & ~31; c0 < L; c0 += 32) {
    const int g_lo = max(0, (q_first - c0) >> 4);
    const int g_hi = min(2, (L - c0 + 15) >> 4);
    float st[4][4], dpt[4][4];
    zero_acc(st);
    zero_acc(dpt);
    mma_rows_x_cols<2>(st, sK, row0, sQ, c0, g_lo, g_hi, lane);
    mma_rows_x_cols<2>(dpt, sV, row0, sdO, c0, g_lo, g_hi, lane);
    // key rows >= L only pollute their own (never stored) dK/dV rows; the query (column) mask is
    // needed only where the chunk crosses L or the causal diagonal of this warp's 16 keys
    const bool need_mask = (c0 + 32 > L) || (CAUSAL && c0 < k0 + row0 + 15);
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      if (nt >= 2 * g_lo && nt < 2 * g_hi) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = c0 + nt * 8 + (lane & 3) * 2 + (e & 1);  // query index (column), < Lp here
          const int r = e >> 1;
          float p = exp2f(st[nt][e] * scale_log2e - sLse[qi]);
          if (need_mask) {
            const bool ok = qi < L && (!CAUSAL || krow[r] <= qi);
            p = ok ? p : 0.f;
          }
          st[nt][e] = p;                               // P^T
          dpt[nt][e] = p * (dpt[nt][e] - sD[qi]);      // dS^T (unscaled), in place
        }
      }
    }
    mma_p_x_tile<2>(dv, st, sdO, c0, g_lo, g_hi, lane);
    mma_p_x_tile<2>(dk, dpt, sQ, c0, g_lo, g_hi, lane);
  }
  bf16* out = dqkv + seq_row * ld + h * DH;
  store_rows_bf16(dk, scale, scale, smem, 0, row0, out + d, ld, k0 + row0, L, lane);
  store_rows_bf16(dv, 1.f, 1.f, smem, BKV * ROW_BYTES, row0, out + 2 * d, ld, k0 + row0, L, lane);
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
// Warps (16-row tiles) per CTA.  The problems are tiny, so what matters is how many independent
// load->compute->store chains an SM has in flight: short sequences (text) use 2-warp CTAs, which
// doubles the resident CTAs per SM; long ones (vision) are limited by the K/V tile in shared memory
// anyway and keep 4 warps per CTA to amortise it.  MUDPT_ATTN_WARPS overrides (tuning aid).
static int pick_warps(int L) {
  const int tiles = (L + 15) / 16;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("MUDPT_ATTN_WARPS");
    forced = e ? atoi(e) : 0;
  }
  int nw = forced > 0 ? forced : (L <= 128 ? 2 : 4);
  if (nw > 4) nw = 4;
  return tiles >= nw ? nw : tiles;
}

template <typename K>
static const char* set_smem(K kern, size_t bytes) {
  if (bytes > 227 * 1024) return "attention: sequence too long for the shared-memory resident kernel";
  if (bytes > 48 * 1024 &&
      cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)) != cudaSuccess)
    return "attention: cudaFuncSetAttribute failed";
  return nullptr;
}

const char* attention_fwd(const bf16* qkv, bf16* o, float* lse2, int S, int L, int H, int d, bool causal,
                          cudaStream_t stream) {
  if (S <= 0 || L <= 0) return nullptr;
  if (d != H * DH) return "attention: head width must be 64";
  const int nw = pick_warps(L), BQ = nw * 16, Lp = (L + 15) & ~15;
  const size_t smem = static_cast<size_t>(BQ + 2 * Lp) * ROW_BYTES;
  const float sl2 = 0.125f * 1.4426950408889634f;
  dim3 grid((L + BQ - 1) / BQ, S * H);
  const char* e;
  if (causal) {
    if ((e = set_smem(attn_fwd_kernel<true>, smem))) return e;
    attn_fwd_kernel<true><<<grid, nw * 32, smem, stream>>>(qkv, o, lse2, L, H, d, sl2);
  } else {
    if ((e = set_smem(attn_fwd_kernel<false>, smem))) return e;
    attn_fwd_kernel<false><<<grid, nw * 32, smem, stream>>>(qkv, o, lse2, L, H, d, sl2);
  }
  count_launch(1);
  return launch_status("attention fwd launch failed");
}

const char* attention_bwd(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv,
                          int S, int L, int H, int d, bool causal, cudaStream_t stream) {
  if (S <= 0 || L <= 0) return nullptr;
  if (d != H * DH) return "attention: head width must be 64";
  const int nw = pick_warps(L), BQ = nw * 16, Lp = (L + 15) & ~15;
  const size_t smem_q = static_cast<size_t>(3 * BQ + 2 * Lp) * ROW_BYTES;
  const size_t smem_kv = static_cast<size_t>(2 * BQ + 2 * Lp) * ROW_BYTES + 2 * Lp * sizeof(float);
  const float scale = 0.125f, sl2 = 0.125f * 1.4426950408889634f;
  dim3 grid((L + BQ - 1) / BQ, S * H);
  const char* e;
  if (causal) {
    if ((e = set_smem(attn_bwd_dq_kernel<true>, smem_q))) return e;
    if ((e = set_smem(attn_bwd_dkv_kernel<true>, smem_kv))) return e;
    attn_bwd_dq_kernel<true><<<grid, nw * 32, smem_q, stream>>>(qkv, o, d_o, lse2, dsum, dqkv, L, H, d, scale, sl2);
    attn_bwd_dkv_kernel<true><<<grid, nw * 32, smem_kv, stream>>>(qkv, d_o, lse2, dsum, dqkv, L, H, d, scale, sl2);
  } else {
    if ((e = set_smem(attn_bwd_dq_kernel<false>, smem_q))) return e;
    if ((e = set_smem(attn_bwd_dkv_kernel<false>, smem_kv))) return e;
    attn_bwd_dq_kernel<false><<<grid, nw * 32, smem_q, stream>>>(qkv, o, d_o, lse2, dsum, dqkv, L, H, d, scale, sl2);
    attn_bwd_dkv_kernel<false><<<grid, nw * 32, smem_kv, stream>>>(qkv, d_o, lse2, dsum, dqkv, L, H, d, scale, sl2);
  }
  count_launch(2);
  return launch_status("attention bwd launch failed");
}

}  // namespace mudpt
