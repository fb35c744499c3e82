// Blackwell-native attention for the CLIP towers: tcgen05.mma with the score / probability tiles in TMEM, operands
// staged by TMA.  Replaces the scaled_dot_product_attention inside nn.MultiheadAttention (clip/model.py:271-273;
// causal mask of the text tower :810-816) for sequences of up to 256 tokens, head width 64.
//
// Forward, one CTA per (sequence, head, block of 128 query rows), 288 threads:
//   warp 8 (one elected lane)  TMA: Q block [128 x 64], K and V [Lk x 64] (3D descriptors over [S][L][3d]: rows past
//                              the sequence end are zero-filled, never the next sequence's tokens)
//                              S = Q K^T : tcgen05.mma 128 x Lk x 64, fp32 scores in TMEM columns [0, Lk)
//                              O = P V   : tcgen05.mma 128 x 64 x Lk, V as the MN-major B operand (no transpose),
//                                          P from shared memory, fp32 output in TMEM columns [0, 64) (S is dead)
//   warps 0-7 (thread = row,   tcgen05.ld of the thread's own score row in 32-column pieces: row max, then
//    two warps per row: each   p = exp2(s c - m c), row sum -- no shuffles; the two halves of a row exchange their
//    half of the keys)         max / sum through 1 KB of shared memory; P as bf16 into
//                              the 128B-swizzled K-major layout the second MMA reads; log-sum-exp (log2 domain) saved;
//                              O / l -> bf16 -> TMA store (clipped at the sequence end)
// Shared memory: Q 16 KB | K | V | extra; the P slabs (64 keys = 16 KB each) reuse the Q tile, then the K tile (both
// dead once the scores are in TMEM), then the extra space: 100 KB for 199 tokens -> 2 CTAs per SM, each with 256 TMEM
// columns, so one CTA's softmax overlaps the other's loads and MMAs.
#include "attention.h"

#include <cstdlib>

#include "common.cuh"
#include "gemm.h"
#include "launch_count.h"

namespace mudpt {

static constexpr int TC_ROWS = 128;      // query rows per CTA = TMEM lanes
static constexpr int TC_SLAB = 16384;    // 128 rows x 64 bf16

__device__ __forceinline__ float f2_hsum_tc(f32x2 v) {
  float lo, hi;
  f2_unpack(v, lo, hi);
  return lo + hi;
}

__host__ __device__ inline int tc_smem_bytes(int Lk) {
  const int kv = Lk * 128;
  const int nslab = (Lk + 63) / 64;
  int extra = nslab - 1 - (kv >= TC_SLAB ? 1 : 0);  // slab 0 = Q tile, slab 1 = K tile when it is large enough
  if (extra < 0) extra = 0;
  return TC_SLAB + 2 * kv + extra * TC_SLAB + 1024 /*exchange*/ + 256 /*barriers*/ + 1024 /*alignment slack*/;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(288) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q,
                                                          const __grid_constant__ CUtensorMap map_kv,
                                                          const __grid_constant__ CUtensorMap map_o, float* __restrict__ lse2,
                                                          const int L, const int H, const int d, const int Lk,
                                                          const int tmem_cols, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x, sh = blockIdx.y, s = sh / H, h = sh - s * H;
  const int kv_bytes = Lk * 128;
  uint8_t* sQ = smem;
  uint8_t* sK = sQ + TC_SLAB;
  uint8_t* sV = sK + kv_bytes;
  uint8_t* sX = sV + kv_bytes;
  const int nslab = (Lk + 63) >> 6;
  const bool k_slab = kv_bytes >= TC_SLAB;
  auto slab = [&](int j) -> uint8_t* {  // P columns [64 j, 64 j + 64): 128 rows x 128 B, 128B swizzle
    if (j == 0) return sQ;
    if (k_slab) return j == 1 ? sK : sX + (j - 2) * TC_SLAB;
    return sX + (j - 1) * TC_SLAB;
  };
  int extra = nslab - 1 - (k_slab ? 1 : 0);
  if (extra < 0) extra = 0;
  float* sEx = reinterpret_cast<float*>(sX + extra * TC_SLAB);  // [2][128]: row max / row sum of the other column half
  uint64_t* bars = reinterpret_cast<uint64_t*>(sEx + 256);
  uint64_t* bar_load = bars;      // TMA bytes
  uint64_t* bar_s = bars + 1;     // scores in TMEM
  uint64_t* bar_p = bars + 2;     // P in shared memory (256 arrivals)
  uint64_t* bar_o = bars + 3;     // output in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_kv);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, 256);
      mbar_init(bar_o, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, static_cast<uint32_t>(tmem_cols));
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 8) {
    if (lane == 0) {
      mbar_expect_tx(bar_load, static_cast<uint32_t>(TC_SLAB + 2 * kv_bytes));
      tma_load_3d(sQ, &map_q, bar_load, h * 64, mb * TC_ROWS, s);
      tma_load_3d(sK, &map_kv, bar_load, d + h * 64, 0, s);
      tma_load_3d(sV, &map_kv, bar_load, 2 * d + h * 64, 0, s);
      mbar_wait(bar_load, 0);
      tc_fence_after();
      {  // S = Q K^T
        const uint32_t idesc = make_idesc_bf16(TC_ROWS, Lk);
        const uint64_t da = make_smem_desc_sw128(smem_u32(sQ)), db = make_smem_desc_sw128(smem_u32(sK));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_base, da + static_cast<uint64_t>(k * 2), db + static_cast<uint64_t>(k * 2), idesc, static_cast<uint32_t>(k != 0));
        umma_commit(bar_s);
      }
      mbar_wait(bar_p, 0);
      tc_fence_after();
      {  // O = P V: contraction over the keys, 16 per instruction
        const uint32_t idesc = make_idesc_bf16(TC_ROWS, 64) | kIdescBMnMajor;
        const uint64_t dv = make_smem_desc_sw128(smem_u32(sV));
        const int nk = Lk >> 4;
        for (int k = 0; k < nk; ++k) {
          const uint64_t da = make_smem_desc_sw128(smem_u32(slab(k >> 2))) + static_cast<uint64_t>((k & 3) * 2);
          umma_bf16(tmem_base, da, dv + static_cast<uint64_t>(k * 128), idesc, static_cast<uint32_t>(k != 0));
        }
        umma_commit(bar_o);
      }
    }
  } else {
    // ===================== softmax / epilogue: thread = query row, warp half = half of the key columns =====================
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;           // row inside the block = TMEM lane
    const int row = mb * TC_ROWS + t;         // token index of the row
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    // keys this row sees (rows past the sequence end compute on zero-filled Q and are never stored)
    const int nvis = CAUSAL ? (row < L ? row + 1 : L) : L;
    const int n32 = (Lk + 31) >> 5, nh0 = (n32 + 1) >> 1;
    const int j_lo = half ? nh0 : 0, j_hi = half ? n32 : nh0;  // this half's 32-column pieces
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int j = j_lo; j < j_hi; ++j) {
      uint32_t r[32];
      tmem_ld_32x32(trow + static_cast<uint32_t>(j * 32), r);
      tmem_ld_wait_regs(r);
      if (j * 32 + 32 <= nvis) {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, j * 32 + i < nvis ? __uint_as_float(r[i]) : -INFINITY);
      }
    }
    sEx[half * 128 + t] = mx;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    mx = fmaxf(mx, sEx[(half ^ 1) * 128 + t]);
    const float m2 = mx * scale_log2e;  // scaled max, log2 domain (nvis >= 1: finite)
    float l = 0.f;
    const f32x2 sc2 = f2_pack(scale_log2e, scale_log2e), nm2 = f2_pack(-m2, -m2);
    for (int j = j_lo; j < j_hi; ++j) {
      uint32_t r[32];
      tmem_ld_32x32(trow + static_cast<uint32_t>(j * 32), r);
      tmem_ld_wait_regs(r);
      uint32_t pk[16];
      const bool full = j * 32 + 32 <= nvis;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        float a, b;
        f2_unpack(f2_fma(f2_pack_u(r[2 * i], r[2 * i + 1]), sc2, nm2), a, b);
        a = exp2f(a);
        b = exp2f(b);
        if (!full) {
          a = j * 32 + 2 * i < nvis ? a : 0.f;
          b = j * 32 + 2 * i + 1 < nvis ? b : 0.f;
        }
        l += a + b;
        pk[i] = pack_bf16(a, b);
      }
      // 32 keys = four 16 B chunks of this row in slab j / 2 (64 keys per slab)
      uint8_t* ps = slab(j >> 1) + t * 128;
      const int c = (j & 1) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(ps + (((c + q) ^ (t & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    tc_fence_before();          // every tcgen05.ld of the scores has completed: the MMA may overwrite them with O
    fence_proxy_async_smem();   // P is visible to the tensor core's shared-memory reads
    mbar_arrive(bar_p);
    asm volatile("bar.sync 1, 256;" ::: "memory");  // (the max exchange has been read by everyone)
    sEx[half * 128 + t] = l;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += sEx[(half ^ 1) * 128 + t];
    if (half == 0 && row < L) lse2[(static_cast<size_t>(s) * H + h) * L + row] = m2 + log2f(l);
    const float inv = 1.f / l;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    // O / l as bf16 into the (dead) Q tile: 128 rows x 128 B, 128B swizzle; this half's 32 head columns
    {
      uint32_t r[32];
      tmem_ld_32x32(trow + static_cast<uint32_t>(half * 32), r);
      tmem_ld_wait_regs(r);
      uint8_t* po = sQ + t * 128;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(r[8 * q + 2 * e]) * inv, __uint_as_float(r[8 * q + 2 * e + 1]) * inv);
        *reinterpret_cast<uint4*>(po + (((half * 4 + q) ^ (t & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0) {
      tma_store_3d(&map_o, sQ, h * 64, mb * TC_ROWS, s);
      bulk_commit();
      bulk_wait_read<0>();  // the tile has been read out of shared memory; the global writes complete on their own
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
  }
}

// =======================================================================================
// Backward on tcgen05.  Two launches of one kernel template, as in the warp-MMA path:
//   KV = false  rows = 128 queries of the block:  S = Q K^T, dP = dO V^T, dS = P (dP - D),  dQ = scale dS K
//   KV = true   rows = 128 keys of the block:     S^T = K Q^T, dP^T = V dO^T, P^T, dS^T,     dV = P^T dO, dK = scale dS^T Q
// The other index runs in ROUNDS of 64 columns: per round two score MMAs (128 x 64 x 64) into TMEM columns [0, 64) /
// [64, 128), the element-wise step by 8 warps (thread = row, each warp half takes 32 columns; P recomputed from the
// saved log-sum-exp: no row reduction at all), the bf16 P / dS tiles into one 16 KB swizzled slab each, and the output
// MMAs accumulating in TMEM columns [128, 192) / [192, 256) over the rounds, with the round's K / V (resp. Q / dO)
// tile as the MN-major B operand.  256 TMEM columns and 96 KB of shared memory per CTA: two CTAs per SM.
// D = rowsum(dO * O) is computed by the KV = false launch (thread-local: the thread owns the row) and handed to the
// KV = true launch through `dsum`.  Without a mask nothing needs masking: padded key / query rows are zero-filled by
// TMA, so their contributions vanish in the output MMAs whatever their (finite) P.
// =======================================================================================
static constexpr int TCB_THREADS = 288;  // 8 element-wise warps + 1 control warp
static constexpr int TCB_TILE = 8192;    // 64 rows x 64 bf16

__host__ __device__ inline int tcb_smem_bytes(int Lpad) {
  return 2 * TC_SLAB /*A tiles*/ + 4 * TCB_TILE /*round tiles x2*/ + 2 * TC_SLAB /*slabs*/ + 2 * Lpad * 4 /*lse, D*/ + 2 * 128 * 4 /*exchange*/ +
         256 /*barriers*/ + 1024;
}

template <bool KV, bool CAUSAL>
__global__ void __launch_bounds__(TCB_THREADS) attn_tc_bwd_kernel(const __grid_constant__ CUtensorMap map_qkv128,  // row blocks of qkv
                                                                  const __grid_constant__ CUtensorMap map_qkv64,   // round tiles of qkv
                                                                  const __grid_constant__ CUtensorMap map_do128,
                                                                  const __grid_constant__ CUtensorMap map_do64,
                                                                  const __grid_constant__ CUtensorMap map_dqkv,    // output, 128-row boxes
                                                                  const bf16* __restrict__ o, const float* __restrict__ lse2,
                                                                  float* __restrict__ dsum, const int L, const int H, const int d,
                                                                  const float scale, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int blk = blockIdx.x, sh = blockIdx.y, s = sh / H, h = sh - s * H;
  const int Lpad = (L + 63) & ~63;
  uint8_t* sA0 = smem;                 // KV ? K block : Q block
  uint8_t* sA1 = sA0 + TC_SLAB;        // KV ? V block : dO block
  uint8_t* sU = sA1 + TC_SLAB;         // [2] KV ? Q round : K round
  uint8_t* sW = sU + 2 * TCB_TILE;     // [2] KV ? dO round : V round
  uint8_t* slabP = sW + 2 * TCB_TILE;  // P^T of the round (KV only)
  uint8_t* slabS = slabP + TC_SLAB;    // dS / dS^T of the round
  float* sLse = reinterpret_cast<float*>(slabS + TC_SLAB);  // KV: per query column
  float* sD = sLse + Lpad;
  float* sX = sD + Lpad;               // [2][128] exchange between the two warp halves of a row
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 256);
  uint64_t* bar_a = bars;          // block tiles loaded
  uint64_t* bar_t = bars + 1;      // [2] round tiles loaded
  uint64_t* bar_s = bars + 3;      // scores of the round in TMEM
  uint64_t* bar_c = bars + 4;      // scores consumed (256 arrivals)
  uint64_t* bar_p = bars + 5;      // slabs written (256 arrivals)
  uint64_t* bar_f = bars + 6;      // output MMAs of the round done: slabs and the round's tile buffer are free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

  // rounds: queries of the block see keys [0, kv_end), keys of the block are seen by queries [q_begin, L)
  const int row_blk0 = blk * TC_ROWS;
  int c_begin = 0, c_end = L;
  if (CAUSAL) {
    if (KV) c_begin = row_blk0 & ~63;                      // queries before the block's first key never see it
    else c_end = min(L, row_blk0 + TC_ROWS);               // keys after the block's last query are invisible
  }
  const int r_begin = c_begin >> 6, r_end = (c_end + 63) >> 6;

  if (warp == 8) {
    if (lane == 0) {
      tma_prefetch_desc(&map_qkv128);
      tma_prefetch_desc(&map_qkv64);
      tma_prefetch_desc(&map_do128);
      tma_prefetch_desc(&map_do64);
      mbar_init(bar_a, 1);
      mbar_init(&bar_t[0], 1);
      mbar_init(&bar_t[1], 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_c, 256);
      mbar_init(bar_p, 256);
      mbar_init(bar_f, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 256u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  const int colA0 = (KV ? d : 0) + h * 64, colU = (KV ? 0 : d) + h * 64;  // columns of qkv: A0 / U tiles
  if (warp == 8) {
    if (lane == 0) {
      auto load_round = [&](int r) {
        const int b = (r - r_begin) & 1;
        mbar_expect_tx(&bar_t[b], 2 * TCB_TILE);
        tma_load_3d(sU + b * TCB_TILE, &map_qkv64, &bar_t[b], colU, r * 64, s);
        if (KV) tma_load_3d(sW + b * TCB_TILE, &map_do64, &bar_t[b], h * 64, r * 64, s);
        else tma_load_3d(sW + b * TCB_TILE, &map_qkv64, &bar_t[b], 2 * d + h * 64, r * 64, s);
      };
      mbar_expect_tx(bar_a, 2 * TC_SLAB);
      tma_load_3d(sA0, &map_qkv128, bar_a, colA0, row_blk0, s);
      if (KV) tma_load_3d(sA1, &map_qkv128, bar_a, 2 * d + h * 64, row_blk0, s);
      else tma_load_3d(sA1, &map_do128, bar_a, h * 64, row_blk0, s);
      load_round(r_begin);
      if (r_begin + 1 < r_end) load_round(r_begin + 1);
      mbar_wait(bar_a, 0);
      const uint64_t dA0 = make_smem_desc_sw128(smem_u32(sA0)), dA1 = make_smem_desc_sw128(smem_u32(sA1));
      const uint64_t dP = make_smem_desc_sw128(smem_u32(slabP)), dS = make_smem_desc_sw128(smem_u32(slabS));
      const uint32_t idesc_out = make_idesc_bf16(TC_ROWS, 64) | kIdescBMnMajor;
      for (int r = r_begin; r < r_end; ++r) {
        const int i = r - r_begin, b = i & 1;
        const int nr = min(64, ((c_end + 15) & ~15) - r * 64);  // columns of this round (multiple of 16)
        mbar_wait(&bar_t[b], (i >> 1) & 1);
        if (i > 0) mbar_wait(bar_c, (i - 1) & 1);  // the previous round's scores have been read
        tc_fence_after();
        const uint64_t dU = make_smem_desc_sw128(smem_u32(sU + b * TCB_TILE)), dW = make_smem_desc_sw128(smem_u32(sW + b * TCB_TILE));
        const uint32_t idesc_s = make_idesc_bf16(TC_ROWS, nr);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base, dA0 + static_cast<uint64_t>(k * 2), dU + static_cast<uint64_t>(k * 2), idesc_s, static_cast<uint32_t>(k != 0));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem_base + 64u, dA1 + static_cast<uint64_t>(k * 2), dW + static_cast<uint64_t>(k * 2), idesc_s, static_cast<uint32_t>(k != 0));
        umma_commit(bar_s);
        // the buffer of round r - 1 is free once its output MMAs are done: refill it with round r + 1
        if (i >= 1 && r + 1 < r_end) {
          mbar_wait(bar_f, (i - 1) & 1);
          load_round(r + 1);
        }
        mbar_wait(bar_p, i & 1);  // the round's slabs are written
        tc_fence_after();
        const int nk = nr >> 4;
        for (int k = 0; k < nk; ++k) {
          const uint32_t acc = static_cast<uint32_t>(i != 0 || k != 0);
          if (KV) {
            umma_bf16(tmem_base + 128u, dP + static_cast<uint64_t>(k * 2), dW + static_cast<uint64_t>(k * 128), idesc_out, acc);  // dV += P^T dO
            umma_bf16(tmem_base + 192u, dS + static_cast<uint64_t>(k * 2), dU + static_cast<uint64_t>(k * 128), idesc_out, acc);  // dK += dS^T Q
          } else {
            umma_bf16(tmem_base + 128u, dS + static_cast<uint64_t>(k * 2), dU + static_cast<uint64_t>(k * 128), idesc_out, acc);  // dQ += dS K
          }
        }
        umma_commit(bar_f);
      }
    }
  } else {
    // ===================== element-wise warps: thread = row of the block, half = 32 of the round's 64 columns =====================
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;
    const int row = row_blk0 + t;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const size_t stat_base = (static_cast<size_t>(s) * H + h) * L;
    float lse_r = 0.f, D_r = 0.f;  // KV = false: of this thread's query row
    if (KV) {
      for (int i = threadIdx.x; i < Lpad; i += 256) {
        sLse[i] = i < L ? lse2[stat_base + i] : 0.f;
        sD[i] = i < L ? dsum[stat_base + i] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
    } else {
      // D = rowsum(dO * O): this half's 32 of the 64 head columns, exchanged through shared memory
      mbar_wait(bar_a, 0);
      float part = 0.f;
      if (row < L) {
        const bf16* orow = o + (static_cast<size_t>(s) * L + row) * d + h * 64 + half * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const uint4 a = *reinterpret_cast<const uint4*>(sA1 + t * 128 + (((half * 4 + c) ^ (t & 7)) << 4));
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(orow) + c);
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float2 x = unpack_bf16(aw[q]), y = unpack_bf16(bw[q]);
            part += x.x * y.x + x.y * y.y;
          }
        }
        lse_r = lse2[stat_base + row];
      }
      sX[half * 128 + t] = part;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      D_r = part + sX[(half ^ 1) * 128 + t];
      if (half == 0 && row < L) dsum[stat_base + row] = D_r;
    }
    const f32x2 c2 = f2_pack(scale_log2e, scale_log2e);
    for (int r = r_begin; r < r_end; ++r) {
      const int i = r - r_begin;
      mbar_wait(bar_s, i & 1);
      tc_fence_after();
      uint32_t sv[32], dv[32];
      tmem_ld_32x32(trow + static_cast<uint32_t>(half * 32), sv);
      tmem_ld_32x32(trow + 64u + static_cast<uint32_t>(half * 32), dv);
      tmem_ld_wait_regs(sv);
      tmem_ld_wait_regs(dv);
      tc_fence_before();
      mbar_arrive(bar_c);  // the next round's score MMAs may overwrite the columns
      const int col0 = r * 64 + half * 32;  // first column (key for KV = false, query for KV = true) of this thread's 32
      uint32_t pp[16], ds[16];
#pragma unroll
      for (int e = 0; e < 16; ++e) {
        f32x2 nl, nD;
        if (KV) {
          const float2 l2 = *reinterpret_cast<const float2*>(sLse + col0 + 2 * e);
          const float2 d2 = *reinterpret_cast<const float2*>(sD + col0 + 2 * e);
          nl = f2_pack(-l2.x, -l2.y);
          nD = f2_pack(-d2.x, -d2.y);
        } else {
          nl = f2_pack(-lse_r, -lse_r);
          nD = f2_pack(-D_r, -D_r);
        }
        float a, b;
        f2_unpack(f2_fma(f2_pack_u(sv[2 * e], sv[2 * e + 1]), c2, nl), a, b);
        a = exp2f(a);
        b = exp2f(b);
        if (CAUSAL) {  // key <= query
          const int c = col0 + 2 * e;
          if (KV) { a = row <= c ? a : 0.f; b = row <= c + 1 ? b : 0.f; }
          else { a = c <= row ? a : 0.f; b = c + 1 <= row ? b : 0.f; }
        }
        float x, y;
        f2_unpack(f2_mul(f2_pack(a, b), f2_add(f2_pack_u(dv[2 * e], dv[2 * e + 1]), nD)), x, y);  // dS (unscaled)
        pp[e] = pack_bf16(a, b);
        ds[e] = pack_bf16(x, y);
      }
      if (i > 0) mbar_wait(bar_f, (i - 1) & 1);  // the previous round's output MMAs have read the slabs
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * 4 + q) ^ (t & 7)) << 4));
        if (KV) *reinterpret_cast<uint4*>(slabP + off) = make_uint4(pp[4 * q], pp[4 * q + 1], pp[4 * q + 2], pp[4 * q + 3]);
        *reinterpret_cast<uint4*>(slabS + off) = make_uint4(ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
      }
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
    }
    // outputs: this half's 32 of the 64 head columns, scaled, as bf16 into the (dead) block tiles, then TMA stores
    mbar_wait(bar_f, (r_end - r_begin - 1) & 1);
    tc_fence_after();
    {
      uint32_t o0[32], o1[32];
      tmem_ld_32x32(trow + 128u + static_cast<uint32_t>(half * 32), o0);
      if (KV) tmem_ld_32x32(trow + 192u + static_cast<uint32_t>(half * 32), o1);
      tmem_ld_wait_regs(o0);
      if (KV) tmem_ld_wait_regs(o1);
      const float sc0 = KV ? 1.f : scale;  // dV unscaled; dQ, dK carry the softmax scale
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * 4 + q) ^ (t & 7)) << 4));
        uint32_t w[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o0[8 * q + 2 * e]) * sc0, __uint_as_float(o0[8 * q + 2 * e + 1]) * sc0);
        *reinterpret_cast<uint4*>(sA0 + off) = make_uint4(w[0], w[1], w[2], w[3]);
        if (KV) {
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o1[8 * q + 2 * e]) * scale, __uint_as_float(o1[8 * q + 2 * e + 1]) * scale);
          *reinterpret_cast<uint4*>(sA1 + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0) {
      if (KV) {
        tma_store_3d(&map_dqkv, sA0, 2 * d + h * 64, row_blk0, s);  // dV
        tma_store_3d(&map_dqkv, sA1, d + h * 64, row_blk0, s);      // dK
      } else {
        tma_store_3d(&map_dqkv, sA0, h * 64, row_blk0, s);          // dQ
      }
      bulk_commit();
      bulk_wait_read<0>();
    }
  }
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 256u);
  }
}

// =======================================================================================
// Backward for single-block sequences (L <= 128: the text tower), PERSISTENT: one CTA per SM walks the (sequence, head)
// problems; a TMA ring of 3 stages (Q, dO, K, V tiles of the problem, rows clipped to the sequence) keeps HBM busy while
// the current problem computes.  Per problem, FIVE tcgen05 products and no recomputation:
//     S = Q K^T, dP = dO V^T                       (TMEM columns [0,128) / [128,256))
//     element-wise, thread = query row (8 warps, two per row): P = exp2(S c - lse), D = rowsum(P dP) (the halves of a
//         row exchange through shared memory), dS = P (dP - D); P and dS as bf16 into 128B-swizzled slabs [query][key]
//     dQ = dS K       A = dS slab K-major,               B = K tile MN-major      (columns [256,320))
//     dK = dS^T Q     A = dS slab read MN-major (M = keys), B = Q tile MN-major   (columns [320,384))
//     dV = P^T dO     A = P slab read MN-major,          B = dO tile MN-major     (columns [384,448))
//   the transposes are free: the same slab serves as K-major and as MN-major operand.  Outputs leave through the
//   problem's (dead) Q / K / V tiles and three TMA stores.
// Padding needs no masks: rows past the sequence are zero-filled by TMA, so whatever (finite) P they get multiplies
// zero dO / Q / K rows; the causal mask is an index compare.
// =======================================================================================
static constexpr int TCS_THREADS = 320;  // 8 element-wise warps + TMA warp + MMA warp
static constexpr int TCS_MAX_STAGES = 3;
// stages of the TMA ring: 3 while they fit beside the four slabs (sequences of up to 96 tokens), else 2
__host__ __device__ inline int tcs_stages(int Lb) { return Lb <= 96 ? 3 : 2; }
__host__ __device__ inline int tcs_smem_bytes(int Lb) {
  return tcs_stages(Lb) * 4 * Lb * 128 + 4 * TC_SLAB + 2 * 128 * 4 + 256 + 1024;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(TCS_THREADS, 1) attn_tc_bwd_short_kernel(const __grid_constant__ CUtensorMap map_qkv,
                                                                           const __grid_constant__ CUtensorMap map_do,
                                                                           const __grid_constant__ CUtensorMap map_dqkv,
                                                                           const float* __restrict__ lse2, const int L, const int H,
                                                                           const int d, const int Lb, const int n_items,
                                                                           const float scale, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = Lb * 128;           // Lb = sequence length rounded up to 16 rows (<= 128)
  const int stage_bytes = 4 * tile_bytes;    // Q | dO | K | V
  const int TCS_STAGES = tcs_stages(Lb);
  uint8_t* slabP = smem + TCS_STAGES * stage_bytes;  // 2 slabs: keys [0,64), [64,128); 128 query rows each
  uint8_t* slabS = slabP + 2 * TC_SLAB;
  float* sX = reinterpret_cast<float*>(slabS + 2 * TC_SLAB);  // [2][128]: partial D of the other column half
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + 256);
  uint64_t* bar_full = bars;                      // [3] stage loaded
  uint64_t* bar_empty = bars + TCS_MAX_STAGES;        // [3] stage free again
  uint64_t* bar_s = bars + 2 * TCS_MAX_STAGES;        // scores in TMEM
  uint64_t* bar_sfree = bar_s + 1;                // scores consumed (256)
  uint64_t* bar_p = bar_s + 2;                    // slabs written (256)
  uint64_t* bar_o = bar_s + 3;                    // outputs in TMEM
  uint64_t* bar_ofree = bar_s + 4;                // outputs drained (256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_s + 6);

  if (warp == 9) {
    if (lane == 0) {
      tma_prefetch_desc(&map_qkv);
      tma_prefetch_desc(&map_do);
      for (int i = 0; i < TCS_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
      mbar_init(bar_s, 1);
      mbar_init(bar_sfree, 256);
      mbar_init(bar_p, 256);
      mbar_init(bar_o, 1);
      mbar_init(bar_ofree, 256);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int st = 0;
      uint32_t ph = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int s = it / H, h = it - s * H;
        mbar_wait(&bar_empty[st], ph ^ 1);
        uint8_t* base = smem + st * stage_bytes;
        mbar_expect_tx(&bar_full[st], static_cast<uint32_t>(stage_bytes));
        tma_load_3d(base, &map_qkv, &bar_full[st], h * 64, 0, s);                       // Q
        tma_load_3d(base + tile_bytes, &map_do, &bar_full[st], h * 64, 0, s);           // dO
        tma_load_3d(base + 2 * tile_bytes, &map_qkv, &bar_full[st], d + h * 64, 0, s);  // K
        tma_load_3d(base + 3 * tile_bytes, &map_qkv, &bar_full[st], 2 * d + h * 64, 0, s);  // V
        if (++st == TCS_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(TC_ROWS, Lb);
      const uint32_t idesc_dq = make_idesc_bf16(TC_ROWS, 64) | kIdescBMnMajor;
      const uint32_t idesc_dkv = make_idesc_bf16(TC_ROWS, 64) | kIdescAMnMajor | kIdescBMnMajor;
      // MN-major A over two 64-key slabs: leading-dimension byte offset = slab stride
      const uint64_t lbo_slab = static_cast<uint64_t>(TC_SLAB >> 4) << 16;
      const uint64_t dPk = make_smem_desc_sw128(smem_u32(slabP)), dSk = make_smem_desc_sw128(smem_u32(slabS));
      const uint64_t dPm = (dPk & ~(static_cast<uint64_t>(0x3FFF) << 16)) | lbo_slab;
      const uint64_t dSm = (dSk & ~(static_cast<uint64_t>(0x3FFF) << 16)) | lbo_slab;
      const int nk = Lb >> 4;
      int st = 0;
      uint32_t ph = 0;
      int k = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        uint8_t* base = smem + st * stage_bytes;
        const uint64_t dQt = make_smem_desc_sw128(smem_u32(base)), dOt = make_smem_desc_sw128(smem_u32(base + tile_bytes));
        const uint64_t dKt = make_smem_desc_sw128(smem_u32(base + 2 * tile_bytes)), dVt = make_smem_desc_sw128(smem_u32(base + 3 * tile_bytes));
        mbar_wait(&bar_full[st], ph);
        if (k > 0) mbar_wait(bar_sfree, (k - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16(tmem_base, dQt + static_cast<uint64_t>(j * 2), dKt + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16(tmem_base + 128u, dOt + static_cast<uint64_t>(j * 2), dVt + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
        umma_commit(bar_s);
        mbar_wait(bar_p, k & 1);
        if (k > 0) mbar_wait(bar_ofree, (k - 1) & 1);
        tc_fence_after();
        for (int j = 0; j < nk; ++j) {  // contraction over keys (dQ) / queries (dK, dV), 16 per instruction
          const uint32_t acc = static_cast<uint32_t>(j != 0);
          // dQ: A = dS [query][key] K-major -- keys 16 j .. : slab j / 4, 32 B step inside the swizzle row
          const uint64_t a_dq = dSk + static_cast<uint64_t>((j >> 2) * (TC_SLAB >> 4) + (j & 3) * 2);
          umma_bf16(tmem_base + 256u, a_dq, dKt + static_cast<uint64_t>(j * 128), idesc_dq, acc);
          // dK / dV: A = dS / P read MN-major (M = 128 keys over the two slabs), queries 16 j .. = 16 rows = 2048 B
          umma_bf16(tmem_base + 320u, dSm + static_cast<uint64_t>(j * 128), dQt + static_cast<uint64_t>(j * 128), idesc_dkv, acc);
          umma_bf16(tmem_base + 384u, dPm + static_cast<uint64_t>(j * 128), dOt + static_cast<uint64_t>(j * 128), idesc_dkv, acc);
        }
        umma_commit(bar_o);
        if (++st == TCS_STAGES) { st = 0; ph ^= 1; }
      }
    }
  } else {
    // ===================== element-wise + epilogue: thread = row (query, then output row), 2 warps per row =====================
    const int quad = warp & 3, half = warp >> 2;
    const int t = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const bool live = quad * 32 < Lb;                 // (warp-uniform) rows of this warp exist in the tiles
    const int n32 = (Lb + 31) >> 5;                   // 32-column pieces of the scores
    const int nh0 = (n32 + 1) >> 1;
    const int j_lo = half ? nh0 : 0, j_hi = half ? n32 : nh0;
    const f32x2 c2 = f2_pack(scale_log2e, scale_log2e);
    int st = 0;
    int k = 0;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
      const int s = it / H, h = it - s * H;
      uint8_t* base = smem + st * stage_bytes;
      const float lse_r = (live && t < L) ? lse2[(static_cast<size_t>(s) * H + h) * L + t] : 0.f;
      mbar_wait(bar_s, k & 1);
      tc_fence_after();
      // pass 1: P (kept as packed bf16) and this half's part of D = rowsum(P dP)
      uint32_t pk[2][16];
      float dpart = 0.f;
      if (live) {
        const f32x2 nl = f2_pack(-lse_r, -lse_r);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = j_lo + jj;
          if (j < j_hi) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(trow + static_cast<uint32_t>(j * 32), sv);
            tmem_ld_32x32(trow + 128u + static_cast<uint32_t>(j * 32), dv);
            tmem_ld_wait_regs(sv);
            tmem_ld_wait_regs(dv);
            f32x2 acc2 = f2_pack(0.f, 0.f);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float a, b;
              f2_unpack(f2_fma(f2_pack_u(sv[2 * e], sv[2 * e + 1]), c2, nl), a, b);
              a = exp2f(a);
              b = exp2f(b);
              const int c = j * 32 + 2 * e;
              if (CAUSAL) { a = c <= t ? a : 0.f; b = c + 1 <= t ? b : 0.f; }
              if (c >= Lb) { a = 0.f; b = 0.f; dv[2 * e] = 0u; dv[2 * e + 1] = 0u; }  // columns past the score tile hold stale TMEM (0 * Inf = NaN)
              acc2 = f2_fma(f2_pack(a, b), f2_pack_u(dv[2 * e], dv[2 * e + 1]), acc2);
              pk[jj][e] = pack_bf16(a, b);
            }
            dpart += f2_hsum_tc(acc2);
          }
        }
      }
      sX[half * 128 + t] = dpart;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      const float D_r = dpart + sX[(half ^ 1) * 128 + t];
      // pass 2: dS = P (dP - D); both tiles into the slabs [query row][key]
      if (live) {
        const f32x2 nD = f2_pack(-D_r, -D_r);
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
          const int j = j_lo + jj;
          if (j < j_hi) {
            uint32_t dv[32];
            tmem_ld_32x32(trow + 128u + static_cast<uint32_t>(j * 32), dv);
            tmem_ld_wait_regs(dv);
            uint32_t ds[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t w = pk[jj][e];
              const f32x2 p2 = f2_pack_u(w << 16, w & 0xffff0000u);
              float x, y;
              if (j * 32 + 2 * e >= Lb) { dv[2 * e] = 0u; dv[2 * e + 1] = 0u; }
              f2_unpack(f2_mul(p2, f2_add(f2_pack_u(dv[2 * e], dv[2 * e + 1]), nD)), x, y);
              ds[e] = pack_bf16(x, y);
            }
            const uint32_t off = static_cast<uint32_t>((j >> 1) * TC_SLAB + t * 128);
            const int cb = (j & 1) * 4;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t o2 = off + static_cast<uint32_t>(((cb + q) ^ (t & 7)) << 4);
              *reinterpret_cast<uint4*>(slabP + o2) = make_uint4(pk[jj][4 * q], pk[jj][4 * q + 1], pk[jj][4 * q + 2], pk[jj][4 * q + 3]);
              *reinterpret_cast<uint4*>(slabS + o2) = make_uint4(ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_sfree);        // the next problem's score MMAs may overwrite S / dP
      fence_proxy_async_smem();
      mbar_arrive(bar_p);
      // epilogue: dQ (row = query t), dK / dV (row = key t); this half's 32 head columns
      mbar_wait(bar_o, k & 1);
      tc_fence_after();
      if (live) {
        uint32_t o0[32], o1[32], o2[32];
        tmem_ld_32x32(trow + 256u + static_cast<uint32_t>(half * 32), o0);
        tmem_ld_32x32(trow + 320u + static_cast<uint32_t>(half * 32), o1);
        tmem_ld_32x32(trow + 384u + static_cast<uint32_t>(half * 32), o2);
        tmem_ld_wait_regs(o0);
        tmem_ld_wait_regs(o1);
        tmem_ld_wait_regs(o2);
        if (t < Lb) {  // the problem's tiles hold Lb rows: Q <- dQ, K <- dK, V <- dV
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * 4 + q) ^ (t & 7)) << 4));
            uint32_t w[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o0[8 * q + 2 * e]) * scale, __uint_as_float(o0[8 * q + 2 * e + 1]) * scale);
            *reinterpret_cast<uint4*>(base + off) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o1[8 * q + 2 * e]) * scale, __uint_as_float(o1[8 * q + 2 * e + 1]) * scale);
            *reinterpret_cast<uint4*>(base + 2 * tile_bytes + off) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o2[8 * q + 2 * e]), __uint_as_float(o2[8 * q + 2 * e + 1]));
            *reinterpret_cast<uint4*>(base + 3 * tile_bytes + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(bar_ofree);        // the next problem's output MMAs may overwrite dQ / dK / dV
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (threadIdx.x == 0) {
        tma_store_3d(&map_dqkv, base, h * 64, 0, s);
        tma_store_3d(&map_dqkv, base + 2 * tile_bytes, d + h * 64, 0, s);
        tma_store_3d(&map_dqkv, base + 3 * tile_bytes, 2 * d + h * 64, 0, s);
        bulk_commit();
        bulk_wait_read<0>();         // the tiles have been read: the stage may be refilled
        mbar_arrive(&bar_empty[st]);
      }
      if (++st == TCS_STAGES) st = 0;
    }
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// =======================================================================================
// The same backward with TWO problems in flight per SM (sequences of up to 80 tokens: the 77-token text tower).
// One problem's chain has ~10 barrier hops (MMA commit -> tcgen05.ld -> element-wise -> slab -> MMA -> epilogue -> TMA
// store), ~3 us that a single problem in flight leaves exposed (176 us per layer, no better than the warp-MMA kernel).
// Here two groups of 4 warps (thread = query row, the whole row: no exchange between halves) each own a TMEM slot
// (256 columns: S at +0, dP at +96; the outputs dQ / dK / dV at +0 / +64 / +128 reuse the score columns once the group
// has consumed them) and a set of P / dS slabs (Lb rows each -- the 128-row reads of the dQ product run into the
// following tile and only produce rows that are never stored), and the MMA warp serves both, so one problem's hops
// hide behind the other's work.  3-stage TMA ring as above.
// =======================================================================================
static constexpr int TCP_THREADS = 320;
static constexpr int TCP_STAGES = 3;

__host__ __device__ inline int tcp_smem_bytes(int Lb) {
  // (+16 KB: the 128-row A reads of the score products start inside the last stage's Q / dO tiles and run past short ones)
  return 2 * 4 * Lb * 128 /*slabs of both slots*/ + TCP_STAGES * 4 * Lb * 128 + 256 + 1024 + TC_SLAB;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(TCP_THREADS, 1) attn_tc_bwd_pp_kernel(const __grid_constant__ CUtensorMap map_qkv,
                                                                        const __grid_constant__ CUtensorMap map_do,
                                                                        const __grid_constant__ CUtensorMap map_dqkv,
                                                                        const float* __restrict__ lse2, const int L, const int H,
                                                                        const int d, const int Lb, const int n_items,
                                                                        const float scale, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = Lb * 128;           // Lb <= 80 rows
  const int stage_bytes = 4 * tile_bytes;    // Q | dO | K | V
  // slabs first (their 128-row over-reads land in the stages): slot g: P[2] then dS[2], each Lb rows x 128 B
  uint8_t* slabs = smem;
  uint8_t* stages = smem + 2 * 4 * tile_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stages + TCP_STAGES * stage_bytes);
  uint64_t* bar_full = bars;                 // [3]
  uint64_t* bar_empty = bars + 3;            // [3]
  uint64_t* bar_s = bars + 6;                // [2] scores of the slot in TMEM
  uint64_t* bar_p = bars + 8;                // [2] slabs of the slot written (128)
  uint64_t* bar_o = bars + 10;               // [2] outputs of the slot in TMEM
  uint64_t* bar_ofree = bars + 12;           // [2] outputs drained (128)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);

  if (warp == 9) {
    if (lane == 0) {
      tma_prefetch_desc(&map_qkv);
      tma_prefetch_desc(&map_do);
      for (int i = 0; i < TCP_STAGES; ++i) { mbar_init(&bar_full[i], 1); mbar_init(&bar_empty[i], 1); }
      for (int g = 0; g < 2; ++g) {
        mbar_init(&bar_s[g], 1);
        mbar_init(&bar_p[g], 128);
        mbar_init(&bar_o[g], 1);
        mbar_init(&bar_ofree[g], 128);
      }
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();
  const int n_local = (n_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == 8) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int k = 0; k < n_local; ++k) {
        const int it = blockIdx.x + k * gridDim.x;
        const int s = it / H, h = it - s * H;
        const int st = k % TCP_STAGES;
        mbar_wait(&bar_empty[st], ((k / TCP_STAGES) & 1) ^ 1);
        uint8_t* base = stages + st * stage_bytes;
        mbar_expect_tx(&bar_full[st], static_cast<uint32_t>(stage_bytes));
        tma_load_3d(base, &map_qkv, &bar_full[st], h * 64, 0, s);                          // Q
        tma_load_3d(base + tile_bytes, &map_do, &bar_full[st], h * 64, 0, s);              // dO
        tma_load_3d(base + 2 * tile_bytes, &map_qkv, &bar_full[st], d + h * 64, 0, s);     // K
        tma_load_3d(base + 3 * tile_bytes, &map_qkv, &bar_full[st], 2 * d + h * 64, 0, s); // V
      }
    }
  } else if (warp == 9) {
    // ===================== MMA issuer: scores of two problems, then their outputs =====================
    if (lane == 0) {
      const uint32_t idesc_s = make_idesc_bf16(TC_ROWS, Lb);
      const uint32_t idesc_dq = make_idesc_bf16(TC_ROWS, 64) | kIdescBMnMajor;
      const uint32_t idesc_dkv = make_idesc_bf16(TC_ROWS, 64) | kIdescAMnMajor | kIdescBMnMajor;
      const uint64_t lbo_mask = ~(static_cast<uint64_t>(0x3FFF) << 16);
      const uint64_t lbo_slab = static_cast<uint64_t>(tile_bytes >> 4) << 16;  // MN-major A: the second 64 keys = next slab
      const int nk = Lb >> 4;
      auto scores = [&](int k) {
        const int g = k & 1, kk = k >> 1, st = k % TCP_STAGES;
        uint8_t* base = stages + st * stage_bytes;
        const uint64_t dQt = make_smem_desc_sw128(smem_u32(base)), dOt = make_smem_desc_sw128(smem_u32(base + tile_bytes));
        const uint64_t dKt = make_smem_desc_sw128(smem_u32(base + 2 * tile_bytes)), dVt = make_smem_desc_sw128(smem_u32(base + 3 * tile_bytes));
        const uint32_t tb = tmem_base + static_cast<uint32_t>(g * 256);
        mbar_wait(&bar_full[st], (k / TCP_STAGES) & 1);
        if (kk > 0) mbar_wait(&bar_ofree[g], (kk - 1) & 1);  // the slot's previous outputs (same columns) are drained
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16(tb, dQt + static_cast<uint64_t>(j * 2), dKt + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
#pragma unroll
        for (int j = 0; j < 4; ++j) umma_bf16(tb + 96u, dOt + static_cast<uint64_t>(j * 2), dVt + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
        umma_commit(&bar_s[g]);
      };
      auto outputs = [&](int k) {
        const int g = k & 1, kk = k >> 1, st = k % TCP_STAGES;
        uint8_t* base = stages + st * stage_bytes;
        const uint64_t dQt = make_smem_desc_sw128(smem_u32(base)), dOt = make_smem_desc_sw128(smem_u32(base + tile_bytes));
        const uint64_t dKt = make_smem_desc_sw128(smem_u32(base + 2 * tile_bytes));
        uint8_t* sp = slabs + g * 4 * tile_bytes;  // P[0], P[1], dS[0], dS[1]
        const uint64_t dPm = (make_smem_desc_sw128(smem_u32(sp)) & lbo_mask) | lbo_slab;
        const uint64_t dSk = make_smem_desc_sw128(smem_u32(sp + 2 * tile_bytes));
        const uint64_t dSm = (dSk & lbo_mask) | lbo_slab;
        const uint32_t tb = tmem_base + static_cast<uint32_t>(g * 256);
        mbar_wait(&bar_p[g], kk & 1);
        tc_fence_after();
        for (int j = 0; j < nk; ++j) {
          const uint32_t acc = static_cast<uint32_t>(j != 0);
          const uint64_t a_dq = dSk + static_cast<uint64_t>((j >> 2) * (tile_bytes >> 4) + (j & 3) * 2);
          umma_bf16(tb, a_dq, dKt + static_cast<uint64_t>(j * 128), idesc_dq, acc);                                          // dQ += dS K
          umma_bf16(tb + 64u, dSm + static_cast<uint64_t>(j * 128), dQt + static_cast<uint64_t>(j * 128), idesc_dkv, acc);   // dK += dS^T Q
          umma_bf16(tb + 128u, dPm + static_cast<uint64_t>(j * 128), dOt + static_cast<uint64_t>(j * 128), idesc_dkv, acc);  // dV += P^T dO
        }
        umma_commit(&bar_o[g]);
      };
      for (int k0 = 0; k0 < n_local; k0 += 2) {
        scores(k0);
        if (k0 + 1 < n_local) scores(k0 + 1);
        outputs(k0);
        if (k0 + 1 < n_local) outputs(k0 + 1);
      }
    }
  } else {
    // ===================== group g = warp / 4: problems k = g, g + 2, ...; thread = row =====================
    const int g = warp >> 2, quad = warp & 3;
    const int t = quad * 32 + lane;
    const uint32_t trow = tmem_base + static_cast<uint32_t>(g * 256) + (static_cast<uint32_t>(quad * 32) << 16);
    const bool live = quad * 32 < Lb;   // (warp-uniform)
    const bool mine = t < Lb;           // the row exists in the tiles / slabs
    const int n32 = (Lb + 31) >> 5;     // <= 3
    uint8_t* sp = slabs + g * 4 * tile_bytes;
    const f32x2 c2 = f2_pack(scale_log2e, scale_log2e);
    for (int k = g; k < n_local; k += 2) {
      const int kk = k >> 1, st = k % TCP_STAGES;
      const int it = blockIdx.x + k * gridDim.x;
      const int s = it / H, h = it - s * H;
      uint8_t* base = stages + st * stage_bytes;
      const float lse_r = (live && t < L) ? lse2[(static_cast<size_t>(s) * H + h) * L + t] : 0.f;
      mbar_wait(&bar_s[g], kk & 1);
      tc_fence_after();
      if (live) {
        // pass 1: P (packed bf16) and D = rowsum(P dP)
        uint32_t pk[3][16];
        const f32x2 nl = f2_pack(-lse_r, -lse_r);
        f32x2 acc2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < n32) {
            uint32_t sv[32], dv[32];
            tmem_ld_32x32(trow + static_cast<uint32_t>(j * 32), sv);
            tmem_ld_32x32(trow + 96u + static_cast<uint32_t>(j * 32), dv);
            tmem_ld_wait_regs(sv);
            tmem_ld_wait_regs(dv);
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              float a, b;
              f2_unpack(f2_fma(f2_pack_u(sv[2 * e], sv[2 * e + 1]), c2, nl), a, b);
              a = exp2f(a);
              b = exp2f(b);
              const int c = j * 32 + 2 * e;
              if (CAUSAL) { a = c <= t ? a : 0.f; b = c + 1 <= t ? b : 0.f; }
              if (c >= Lb) { a = 0.f; b = 0.f; dv[2 * e] = 0u; dv[2 * e + 1] = 0u; }  // columns past the score tile hold other data (0 * Inf = NaN)
              acc2 = f2_fma(f2_pack(a, b), f2_pack_u(dv[2 * e], dv[2 * e + 1]), acc2);
              pk[j][e] = pack_bf16(a, b);
            }
          }
        }
        const float D_r = f2_hsum_tc(acc2);
        const f32x2 nD = f2_pack(-D_r, -D_r);
        // pass 2: dS = P (dP - D); rows of the slabs that exist (t < Lb)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          if (j < n32) {
            uint32_t dv[32];
            tmem_ld_32x32(trow + 96u + static_cast<uint32_t>(j * 32), dv);
            tmem_ld_wait_regs(dv);
            uint32_t ds[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
              const uint32_t w = pk[j][e];
              float x, y;
              if (j * 32 + 2 * e >= Lb) { dv[2 * e] = 0u; dv[2 * e + 1] = 0u; }
              f2_unpack(f2_mul(f2_pack_u(w << 16, w & 0xffff0000u), f2_add(f2_pack_u(dv[2 * e], dv[2 * e + 1]), nD)), x, y);
              ds[e] = pack_bf16(x, y);
            }
            if (mine) {
              const uint32_t off = static_cast<uint32_t>((j >> 1) * tile_bytes + t * 128);
              const int cb = (j & 1) * 4;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t o2 = off + static_cast<uint32_t>(((cb + q) ^ (t & 7)) << 4);
                *reinterpret_cast<uint4*>(sp + o2) = make_uint4(pk[j][4 * q], pk[j][4 * q + 1], pk[j][4 * q + 2], pk[j][4 * q + 3]);
                *reinterpret_cast<uint4*>(sp + 2 * tile_bytes + o2) = make_uint4(ds[4 * q], ds[4 * q + 1], ds[4 * q + 2], ds[4 * q + 3]);
              }
            }
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      mbar_arrive(&bar_p[g]);
      // epilogue: dQ (row = query t), dK / dV (row = key t): 64 head columns each, in two halves
      mbar_wait(&bar_o[g], kk & 1);
      tc_fence_after();
      if (live) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t o0[32], o1[32], o2[32];
          tmem_ld_32x32(trow + static_cast<uint32_t>(hh * 32), o0);
          tmem_ld_32x32(trow + 64u + static_cast<uint32_t>(hh * 32), o1);
          tmem_ld_32x32(trow + 128u + static_cast<uint32_t>(hh * 32), o2);
          tmem_ld_wait_regs(o0);
          tmem_ld_wait_regs(o1);
          tmem_ld_wait_regs(o2);
          if (mine) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint32_t off = static_cast<uint32_t>(t * 128 + (((hh * 4 + q) ^ (t & 7)) << 4));
              uint32_t w[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o0[8 * q + 2 * e]) * scale, __uint_as_float(o0[8 * q + 2 * e + 1]) * scale);
              *reinterpret_cast<uint4*>(base + off) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
              for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o1[8 * q + 2 * e]) * scale, __uint_as_float(o1[8 * q + 2 * e + 1]) * scale);
              *reinterpret_cast<uint4*>(base + 2 * tile_bytes + off) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
              for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o2[8 * q + 2 * e]), __uint_as_float(o2[8 * q + 2 * e + 1]));
              *reinterpret_cast<uint4*>(base + 3 * tile_bytes + off) = make_uint4(w[0], w[1], w[2], w[3]);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bar_ofree[g]);
      fence_proxy_async_smem();
      if (g == 0) asm volatile("bar.sync 1, 128;" ::: "memory");
      else asm volatile("bar.sync 2, 128;" ::: "memory");
      if (quad == 0 && lane == 0) {
        tma_store_3d(&map_dqkv, base, h * 64, 0, s);
        tma_store_3d(&map_dqkv, base + 2 * tile_bytes, d + h * 64, 0, s);
        tma_store_3d(&map_dqkv, base + 3 * tile_bytes, 2 * d + h * 64, 0, s);
        bulk_commit();
        bulk_wait_read<0>();
        mbar_arrive(&bar_empty[st]);
      }
    }
  }
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

// =======================================================================================
// Backward for two-block sequences (128 < L <= 256: the vision tower, 199 tokens) in ONE launch, PERSISTENT: one CTA per
// SM walks the (sequence, head) problems with the whole problem resident -- Q and dO tiles [Lb x 64], the K / V rows of
// the current key block in shared memory, every accumulator in the 512 TMEM columns -- and nothing is recomputed (the
// two-launch kernel above computes every score and every exp twice).  Rows of the score tiles are KEYS (two blocks kb of
// up to 128), columns are QUERIES in chunks qc of 64; one step = (kb, qc):
//     S^T = K_kb Q_qc^T, dP^T = V_kb dO_qc^T        128 x 64 each, TMEM columns [128 b, +64) / [128 b + 64, +64), b = step & 1
//     element-wise, thread = key row, 8 warps (each half of the CTA takes 32 of the 64 query columns):
//         P^T = exp2(S^T c - lse_q), dS^T = P^T (dP^T - D_q)   -> bf16 slabs [key][query] (128B swizzle)
//     dV_kb += P^T dO_qc, dK_kb += dS^T Q_qc         A = slab K-major, B = dO / Q chunk MN-major       (columns 320 / 256)
//     after the second chunk of a query block qb:  dQ_qb += dS K_kb   A = the two dS^T slabs read MN-major (M = 128
//         queries: the transpose is free), B = K_kb MN-major                                        (columns 384 + 64 qb)
// The score buffers are double: the MMA thread issues the scores of step i + 2 as soon as step i's slabs are written, so
// the element-wise warps -- the bottleneck -- never wait for a product.  dK / dV leave one step after the key block's last
// step, dQ at the end, through the problem's own (dead) K / V / Q tiles and TMA stores.
// Loads never sit on the critical path when the tiles fit twice (Lb <= 208): the Q / dO tiles are double-buffered over
// problems, the two K / V slots are a ring over key blocks (the next problem's first block is loaded while the second
// block of this one computes), and D_q = rowsum(dO_q * O_q) of the NEXT problem is computed in the middle of this one
// (thread q; O row loaded from global memory a step earlier) into a second copy of the per-query arrays.
// Padding: TMA zero-fills rows past the sequence; (key, query) pairs outside it are forced to P = dS = 0 by a select
// wherever a warp's 32 x 32 piece is not entirely inside.
// =======================================================================================
// EW element-wise warps (8 or 16) + TMA warp + MMA warp.  16 warps (4 per scheduler, <= 112 registers, 16 query columns per
// thread and step) hide the tcgen05.ld / MUFU / shared-memory latencies of the per-element chain that 8 warps leave exposed.
__host__ __device__ constexpr int tcl_threads(int EW) { return EW * 32 + 64; }
template <int NC> __device__ __forceinline__ void tcl_ld(uint32_t taddr, uint32_t (&r)[NC]);
template <> __device__ __forceinline__ void tcl_ld<32>(uint32_t taddr, uint32_t (&r)[32]) { tmem_ld_32x32(taddr, r); }
template <> __device__ __forceinline__ void tcl_ld<16>(uint32_t taddr, uint32_t (&r)[16]) { tmem_ld_32x16(taddr, r); }
template <int NC> __device__ __forceinline__ void tcl_wait(uint32_t (&r)[NC]);
template <> __device__ __forceinline__ void tcl_wait<32>(uint32_t (&r)[32]) { tmem_ld_wait_regs(r); }
template <> __device__ __forceinline__ void tcl_wait<16>(uint32_t (&r)[16]) { tmem_ld_wait_regs16(r); }
static constexpr int TCL_KV_SLOT = 2 * TC_SLAB;  // K block | V block, 128 rows each
__host__ __device__ inline int tcl_fixed_bytes() { return 2 * TCL_KV_SLOT + 3 * TC_SLAB + 2 * 2 * 256 * 4 + 256 + 1024; }
__host__ __device__ inline int tcl_qdo_stages(int Lb) { return 2 * 2 * Lb * 128 + tcl_fixed_bytes() <= 227 * 1024 ? 2 : 1; }
__host__ __device__ inline int tcl_smem_bytes(int Lb) { return tcl_qdo_stages(Lb) * 2 * Lb * 128 + tcl_fixed_bytes(); }

template <bool CAUSAL, int EW>
__global__ void __launch_bounds__(tcl_threads(EW), 1) attn_tc_bwd_long_kernel(const __grid_constant__ CUtensorMap map_q,    // qkv, Lb-row boxes
                                                                          const __grid_constant__ CUtensorMap map_do,   // d_o, Lb-row boxes
                                                                          const __grid_constant__ CUtensorMap map_kv,   // qkv, 128-row boxes
                                                                          const __grid_constant__ CUtensorMap map_dq,   // dqkv, Lb-row boxes
                                                                          const __grid_constant__ CUtensorMap map_dkv,  // dqkv, 128-row boxes
                                                                          const bf16* __restrict__ o, const float* __restrict__ lse2,
                                                                          const int L, const int H, const int d, const int Lb,
                                                                          const int n_items, const float scale, const float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_bytes = Lb * 128;  // Lb = sequence length rounded up to 16 rows (144..256): a multiple of 2048 B
  const int QS = tcl_qdo_stages(Lb);
  const bool PF = QS == 2;          // the next problem is prefetched while this one computes
  uint8_t* qdo = smem;                             // [QS][Q tile | dO tile]
  uint8_t* kv = qdo + QS * 2 * tile_bytes;         // [2 key blocks][K rows | V rows]
  uint8_t* slabP = kv + 2 * TCL_KV_SLOT;           // P^T of the step          [128 keys][64 queries]
  uint8_t* slabS = slabP + TC_SLAB;                // dS^T, slot = qc & 1: the two slabs are the 128 queries of a query block
  float* sLse = reinterpret_cast<float*>(slabS + 2 * TC_SLAB);  // [2][256] per query: -lse
  float* sD = sLse + 512;                                       // [2][256]            -D
  uint64_t* bars = reinterpret_cast<uint64_t*>(sD + 512);
  uint64_t* bar_qfull = bars;         // [2] Q / dO tiles of a problem loaded
  uint64_t* bar_qempty = bars + 2;    // [2] ... and free again (dQ stored)
  uint64_t* bar_kvfull = bars + 4;    // [2] K / V rows of key block kb loaded
  uint64_t* bar_kvempty = bars + 6;   // [2] ... and free again (dK / dV stored)
  uint64_t* bar_s = bars + 8;         // [2] scores of a step in TMEM
  uint64_t* bar_p = bars + 10;        // slabs of the step written (256)
  uint64_t* bar_f = bars + 11;        // output MMAs of the step done
  uint64_t* bar_kv = bars + 12;       // dK / dV of the key block complete
  uint64_t* bar_kvfree = bars + 13;   // ... and drained from TMEM (256)
  uint64_t* bar_qfree = bars + 14;    // dQ drained from TMEM (256)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int nq = (Lb + 63) >> 6;  // query chunks (3 or 4)
  const int n_steps = 2 * nq;
  constexpr int NEW = EW * 32;         // element-wise threads
  constexpr int NC = 64 / (EW / 4);    // query columns per thread and step (32 or 16)
  constexpr int TPQ = NEW / 256;       // threads per query in the D computation

  if (warp == EW + 1) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_do);
      tma_prefetch_desc(&map_kv);
      tma_prefetch_desc(&map_dq);
      tma_prefetch_desc(&map_dkv);
      for (int i = 0; i < 2; ++i) {
        mbar_init(&bar_qfull[i], 1);
        mbar_init(&bar_qempty[i], 1);
        mbar_init(&bar_kvfull[i], 1);
        mbar_init(&bar_kvempty[i], 1);
        mbar_init(&bar_s[i], 1);
      }
      mbar_init(bar_p, NEW);
      mbar_init(bar_f, 1);
      mbar_init(bar_kv, 1);
      mbar_init(bar_kvfree, NEW);
      mbar_init(bar_qfree, NEW);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, 512u);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();
  pdl_trigger();

  if (warp == EW) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int k = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        const int s = it / H, h = it - s * H;
        const int qs = PF ? (k & 1) : 0, u = PF ? (k >> 1) : k;  // Q / dO stage and how often it has been used
        if (u > 0) mbar_wait(&bar_qempty[qs], (u - 1) & 1);
        uint8_t* tQ = qdo + qs * 2 * tile_bytes;
        mbar_expect_tx(&bar_qfull[qs], static_cast<uint32_t>(2 * tile_bytes));
        tma_load_3d(tQ, &map_q, &bar_qfull[qs], h * 64, 0, s);
        tma_load_3d(tQ + tile_bytes, &map_do, &bar_qfull[qs], h * 64, 0, s);
        for (int kb = 0; kb < 2; ++kb) {
          if (k > 0) mbar_wait(&bar_kvempty[kb], (k - 1) & 1);
          uint8_t* tK = kv + kb * TCL_KV_SLOT;
          mbar_expect_tx(&bar_kvfull[kb], static_cast<uint32_t>(TCL_KV_SLOT));
          tma_load_3d(tK, &map_kv, &bar_kvfull[kb], d + h * 64, kb * TC_ROWS, s);
          tma_load_3d(tK + TC_SLAB, &map_kv, &bar_kvfull[kb], 2 * d + h * 64, kb * TC_ROWS, s);
        }
      }
    }
  } else if (warp == EW + 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc_out = make_idesc_bf16(TC_ROWS, 64) | kIdescBMnMajor;
      const uint32_t idesc_dq = make_idesc_bf16(TC_ROWS, 64) | kIdescAMnMajor | kIdescBMnMajor;
      const uint64_t dK0 = make_smem_desc_sw128(smem_u32(kv));  // K rows of key block 0; V: + TC_SLAB, block 1: + TCL_KV_SLOT
      const uint64_t dPk = make_smem_desc_sw128(smem_u32(slabP)), dSk = make_smem_desc_sw128(smem_u32(slabS));
      // MN-major A over the two dS^T slabs (M = 128 queries): leading-dimension byte offset = slab stride
      const uint64_t dSm = (dSk & ~(static_cast<uint64_t>(0x3FFF) << 16)) | (static_cast<uint64_t>(TC_SLAB >> 4) << 16);
      uint32_t g = 0;  // steps issued so far (all problems)
      int k = 0;
      for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
        const int qs = PF ? (k & 1) : 0, u = PF ? (k >> 1) : k;
        const uint64_t dQt = make_smem_desc_sw128(smem_u32(qdo + qs * 2 * tile_bytes));
        const uint64_t dOt = dQt + static_cast<uint64_t>(tile_bytes >> 4);
        auto issue_scores = [&](int i) {
          const int kb = i >= nq ? 1 : 0, qc = i - kb * nq, b = i & 1;
          if (qc == 0) {
            mbar_wait(&bar_kvfull[kb], k & 1);
            tc_fence_after();
          }
          const int nr = min(64, Lb - qc * 64);
          const uint32_t idesc_s = make_idesc_bf16(TC_ROWS, nr);
          const uint64_t aK = dK0 + static_cast<uint64_t>(kb * (TCL_KV_SLOT >> 4)), aV = aK + static_cast<uint64_t>(TC_SLAB >> 4);
          const uint64_t bQ = dQt + static_cast<uint64_t>(qc * 512), bO = dOt + static_cast<uint64_t>(qc * 512);
          const uint32_t col = tmem_base + static_cast<uint32_t>(b * 128);
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(col, aK + static_cast<uint64_t>(j * 2), bQ + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
#pragma unroll
          for (int j = 0; j < 4; ++j) umma_bf16(col + 64u, aV + static_cast<uint64_t>(j * 2), bO + static_cast<uint64_t>(j * 2), idesc_s, static_cast<uint32_t>(j != 0));
          umma_commit(&bar_s[b]);
        };
        mbar_wait(&bar_qfull[qs], u & 1);
        issue_scores(0);
        issue_scores(1);
        for (int i = 0; i < n_steps; ++i, ++g) {
          const int kb = i >= nq ? 1 : 0, qc = i - kb * nq;
          const int nr = min(64, Lb - qc * 64);
          mbar_wait(bar_p, g & 1);  // the step's slabs are written (and its score columns read)
          if (qc == 0) {
            const int m = 2 * k + kb;
            if (m > 0) mbar_wait(bar_kvfree, (m - 1) & 1);  // the previous key block's dK / dV have left TMEM
          }
          tc_fence_after();
          const uint64_t bQ = dQt + static_cast<uint64_t>(qc * 512), bO = dOt + static_cast<uint64_t>(qc * 512);
          const uint64_t aS = dSk + static_cast<uint64_t>((qc & 1) * (TC_SLAB >> 4));
          const int nj = nr >> 4;
          for (int j = 0; j < nj; ++j) {  // contraction over the chunk's queries, 16 per instruction
            const uint32_t acc = static_cast<uint32_t>(qc != 0 || j != 0);
            umma_bf16(tmem_base + 256u, aS + static_cast<uint64_t>(j * 2), bQ + static_cast<uint64_t>(j * 128), idesc_out, acc);   // dK += dS^T Q
            umma_bf16(tmem_base + 320u, dPk + static_cast<uint64_t>(j * 2), bO + static_cast<uint64_t>(j * 128), idesc_out, acc);  // dV += P^T dO
          }
          if ((qc & 1) || qc == nq - 1) {  // both chunks of the query block are in the slabs: dQ_qb += dS K_kb
            const int qb = qc >> 1;
            if (kb == 0 && qb == 0 && k > 0) {
              mbar_wait(bar_qfree, (k - 1) & 1);  // the previous problem's dQ has left TMEM
              tc_fence_after();
            }
            const int nkk = (kb ? Lb - TC_ROWS : TC_ROWS) >> 4;
            const uint64_t bK = dK0 + static_cast<uint64_t>(kb * (TCL_KV_SLOT >> 4));
            for (int j = 0; j < nkk; ++j)  // contraction over the block's keys
              umma_bf16(tmem_base + 384u + static_cast<uint32_t>(qb * 64), dSm + static_cast<uint64_t>(j * 128), bK + static_cast<uint64_t>(j * 128), idesc_dq,
                        static_cast<uint32_t>(kb != 0 || j != 0));
          }
          umma_commit(bar_f);
          if (qc == nq - 1) umma_commit(bar_kv);
          if (i + 2 < n_steps) issue_scores(i + 2);
        }
      }
    }
  } else {
    // ===================== element-wise + epilogue: thread = key row of the block; part = NC of the chunk's 64 query columns =====================
    const int quad = warp & 3, half = warp >> 2;  // ("half": which NC-column part of a 64-column chunk / of the 64 head columns)
    const int t = quad * 32 + lane;
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const f32x2 c2 = f2_pack(scale_log2e, scale_log2e);
    // lse and D = rowsum(dO * O) of query q = threadIdx.x of problem `item`, in two halves: the O row is fetched from global
    // memory first, the products with the dO tile (shared memory) come later
    // (TPQ threads per query, adjacent lanes: each takes 8 / TPQ of the row's eight 16-byte pieces)
    constexpr int OV = 8 / TPQ;
    uint4 ov[OV];
    float lse_q = 0.f;
    const int dq_q = threadIdx.x / TPQ, dq_h = threadIdx.x % TPQ;
    auto d_fetch = [&](int item) {
      const int s = item / H, h = item - s * H, q = dq_q;
      lse_q = 0.f;
      if (q < L) {
        const uint4* op = reinterpret_cast<const uint4*>(o + (static_cast<size_t>(s) * L + q) * d + h * 64) + dq_h * OV;
#pragma unroll
        for (int c = 0; c < OV; ++c) ov[c] = __ldg(op + c);
        lse_q = lse2[(static_cast<size_t>(s) * H + h) * L + q];
      }
    };
    auto d_finish = [&](int kk) {  // kk = per-CTA index of the problem
      const int qs = PF ? (kk & 1) : 0, u = PF ? (kk >> 1) : kk, q = dq_q;
      mbar_wait(&bar_qfull[qs], u & 1);
      const uint8_t* tdO = qdo + qs * 2 * tile_bytes + tile_bytes;
      float D = 0.f;
      if (q < L) {
#pragma unroll
        for (int c = 0; c < OV; ++c) {
          const uint4 a = *reinterpret_cast<const uint4*>(tdO + q * 128 + (((dq_h * OV + c) ^ (q & 7)) << 4));
          const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {ov[c].x, ov[c].y, ov[c].z, ov[c].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 x = unpack_bf16(aw[e]), y = unpack_bf16(bw[e]);
            D += x.x * y.x + x.y * y.y;
          }
        }
      }
      if (TPQ == 2) D += __shfl_xor_sync(0xffffffffu, D, 1);
      if (dq_h == 0) {
        sLse[(kk & 1) * 256 + q] = -lse_q;  // both are only ever subtracted
        sD[(kk & 1) * 256 + q] = -D;
      }
    };
    uint32_t g = 0;
    int k = 0;
    bool release_pending = false;  // (thread 0) the previous problem's last stores have not been waited for yet
    int prev_qs = 0;
    if (blockIdx.x < n_items) {
      d_fetch(blockIdx.x);
      d_finish(0);
      asm volatile("bar.sync 1, %0;" ::"n"(NEW) : "memory");
    }
    for (int it = blockIdx.x; it < n_items; it += gridDim.x, ++k) {
      const int s = it / H, h = it - s * H;
      const int qs = PF ? (k & 1) : 0;
      uint8_t* tQ = qdo + qs * 2 * tile_bytes;
      const float* nLse = sLse + (k & 1) * 256;
      const float* nDs = sD + (k & 1) * 256;
      const bool has_next = it + static_cast<int>(gridDim.x) < n_items;
      if (!PF && k > 0) {  // single-buffered tiles: this problem's per-query arrays can only be built now
        d_fetch(it);
        d_finish(k);
        asm volatile("bar.sync 1, %0;" ::"n"(NEW) : "memory");
      }
      // dK, dV of a complete key block (row = key, this half's 32 head columns) into the block's dead K / V rows
      auto drain_kv = [&](int kb) {
        uint8_t* tK = kv + kb * TCL_KV_SLOT;
        mbar_wait(bar_kv, (2 * k + kb) & 1);
        tc_fence_after();
        uint32_t o0[NC], o1[NC];
        tcl_ld<NC>(trow + 256u + static_cast<uint32_t>(half * NC), o0);
        tcl_ld<NC>(trow + 320u + static_cast<uint32_t>(half * NC), o1);
        tcl_wait<NC>(o0);
        tcl_wait<NC>(o1);
        tc_fence_before();
        mbar_arrive(bar_kvfree);
#pragma unroll
        for (int q4 = 0; q4 < NC / 8; ++q4) {
          const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * (NC / 8) + q4) ^ (t & 7)) << 4));
          uint32_t w[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o0[8 * q4 + 2 * e]) * scale, __uint_as_float(o0[8 * q4 + 2 * e + 1]) * scale);
          *reinterpret_cast<uint4*>(tK + off) = make_uint4(w[0], w[1], w[2], w[3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o1[8 * q4 + 2 * e]), __uint_as_float(o1[8 * q4 + 2 * e + 1]));
          *reinterpret_cast<uint4*>(tK + TC_SLAB + off) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      };
      for (int i = 0; i < n_steps; ++i, ++g) {
        const int kb = i >= nq ? 1 : 0, qc = i - kb * nq, b = i & 1;
        const int key = kb * TC_ROWS + t;
        const int key0 = kb * TC_ROWS + quad * 32;  // first key of the warp
        const int col0 = qc * 64 + half * NC;       // first query of this thread's NC columns
        const bool active = col0 < Lb;              // (warp-uniform) the chunk has columns for this part
        // (warp-uniform) every (key, query) of the warp's 32 x NC piece is live / none is: no masks, or no math at all
        const bool all_live = key0 + 32 <= L && col0 + NC <= L && (!CAUSAL || key0 + 31 <= col0);
        const bool none_live = key0 >= L || col0 >= L || (CAUSAL && key0 > col0 + NC - 1);
        if (PF && i == nq && has_next) d_fetch(it + gridDim.x);  // the next problem's O row: in flight during this step
        mbar_wait(&bar_s[b], (g >> 1) & 1);
        tc_fence_after();
        uint32_t pp[NC / 2], ds[NC / 2];
        if (active && !none_live) {
          uint32_t sv[NC], dv[NC];
          tcl_ld<NC>(trow + static_cast<uint32_t>(b * 128 + half * NC), sv);
          tcl_ld<NC>(trow + static_cast<uint32_t>(b * 128 + 64 + half * NC), dv);
          tcl_wait<NC>(sv);
          tcl_wait<NC>(dv);
          if (all_live) {
#pragma unroll
            for (int e4 = 0; e4 < NC / 4; ++e4) {  // 4 queries per iteration
              const float4 nl = *reinterpret_cast<const float4*>(nLse + col0 + 4 * e4);
              const float4 nD = *reinterpret_cast<const float4*>(nDs + col0 + 4 * e4);
              float p0, p1, p2, p3;
              f2_unpack(f2_fma(f2_pack_u(sv[4 * e4], sv[4 * e4 + 1]), c2, f2_pack(nl.x, nl.y)), p0, p1);
              f2_unpack(f2_fma(f2_pack_u(sv[4 * e4 + 2], sv[4 * e4 + 3]), c2, f2_pack(nl.z, nl.w)), p2, p3);
              p0 = exp2f(p0); p1 = exp2f(p1); p2 = exp2f(p2); p3 = exp2f(p3);
              float x0, x1, x2, x3;
              f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack_u(dv[4 * e4], dv[4 * e4 + 1]), f2_pack(nD.x, nD.y))), x0, x1);
              f2_unpack(f2_mul(f2_pack(p2, p3), f2_add(f2_pack_u(dv[4 * e4 + 2], dv[4 * e4 + 3]), f2_pack(nD.z, nD.w))), x2, x3);
              pp[2 * e4] = pack_bf16(p0, p1);
              pp[2 * e4 + 1] = pack_bf16(p2, p3);
              ds[2 * e4] = pack_bf16(x0, x1);
              ds[2 * e4 + 1] = pack_bf16(x2, x3);
            }
          } else {
            // live columns of this thread's key: [c_lo, c_hi)
            const int c_hi = key < L ? L : 0;
            const int c_lo = CAUSAL ? key : 0;
#pragma unroll
            for (int e = 0; e < NC / 2; ++e) {
              const int c = col0 + 2 * e;
              const float2 nl = *reinterpret_cast<const float2*>(nLse + c);
              const float2 nD = *reinterpret_cast<const float2*>(nDs + c);
              float p0, p1;
              f2_unpack(f2_fma(f2_pack_u(sv[2 * e], sv[2 * e + 1]), c2, f2_pack(nl.x, nl.y)), p0, p1);
              p0 = exp2f(p0);
              p1 = exp2f(p1);
              float x, y;
              f2_unpack(f2_mul(f2_pack(p0, p1), f2_add(f2_pack_u(dv[2 * e], dv[2 * e + 1]), f2_pack(nD.x, nD.y))), x, y);  // dS (unscaled)
              // selects, not products: masked positions may hold stale TMEM columns (Inf, NaN)
              const bool ok0 = c >= c_lo && c < c_hi, ok1 = c + 1 >= c_lo && c + 1 < c_hi;
              pp[e] = pack_bf16(ok0 ? p0 : 0.f, ok1 ? p1 : 0.f);
              ds[e] = pack_bf16(ok0 ? x : 0.f, ok1 ? y : 0.f);
            }
          }
        } else {
#pragma unroll
          for (int e = 0; e < NC / 2; ++e) { pp[e] = 0u; ds[e] = 0u; }
        }
        if (g > 0) mbar_wait(bar_f, (g - 1) & 1);  // the previous step's output MMAs have read the slabs
        if (active) {
          uint8_t* sS = slabS + (qc & 1) * TC_SLAB;
#pragma unroll
          for (int q4 = 0; q4 < NC / 8; ++q4) {
            const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * (NC / 8) + q4) ^ (t & 7)) << 4));
            *reinterpret_cast<uint4*>(slabP + off) = make_uint4(pp[4 * q4], pp[4 * q4 + 1], pp[4 * q4 + 2], pp[4 * q4 + 3]);
            *reinterpret_cast<uint4*>(sS + off) = make_uint4(ds[4 * q4], ds[4 * q4 + 1], ds[4 * q4 + 2], ds[4 * q4 + 3]);
          }
        }
        tc_fence_before();
        fence_proxy_async_smem();
        mbar_arrive(bar_p);
        if (i == 0 && threadIdx.x == 0 && release_pending) {
          // the previous problem's last stores have been read by now: its second K / V slot and its Q / dO stage are free
          bulk_wait_read<0>();
          mbar_arrive(&bar_kvempty[1]);
          mbar_arrive(&bar_qempty[prev_qs]);
          release_pending = false;
        }
        if (i == nq) {
          // the first key block's dK / dV leave one step late (its last output MMAs ran under this step's math) ...
          drain_kv(0);
          fence_proxy_async_smem();
          asm volatile("bar.sync 1, %0;" ::"n"(NEW) : "memory");
          if (threadIdx.x == 0) {
            tma_store_3d(&map_dkv, kv, d + h * 64, 0, s);
            tma_store_3d(&map_dkv, kv + TC_SLAB, 2 * d + h * 64, 0, s);
            bulk_commit();
          }
        }
        if (i == nq + 1) {
          // ... and one step after that the slot is free for the next problem's first key block
          if (threadIdx.x == 0) {
            bulk_wait_read<0>();
            mbar_arrive(&bar_kvempty[0]);
          }
          if (PF && has_next) d_finish(k + 1);  // ordered before its readers by the barrier at the end of this problem
        }
      }
      drain_kv(1);
      // dQ of both query blocks (row = query t and 128 + t) into the dead Q tile
      mbar_wait(bar_f, (g - 1) & 1);
      tc_fence_after();
      {
        uint32_t o0[NC], o1[NC];
        tcl_ld<NC>(trow + 384u + static_cast<uint32_t>(half * NC), o0);
        tcl_ld<NC>(trow + 448u + static_cast<uint32_t>(half * NC), o1);
        tcl_wait<NC>(o0);
        tcl_wait<NC>(o1);
        tc_fence_before();
        mbar_arrive(bar_qfree);
#pragma unroll
        for (int q4 = 0; q4 < NC / 8; ++q4) {
          const uint32_t off = static_cast<uint32_t>(t * 128 + (((half * (NC / 8) + q4) ^ (t & 7)) << 4));
          uint32_t w[4];
          if (t < Lb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o0[8 * q4 + 2 * e]) * scale, __uint_as_float(o0[8 * q4 + 2 * e + 1]) * scale);
            *reinterpret_cast<uint4*>(tQ + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
          if (TC_ROWS + t < Lb) {
#pragma unroll
            for (int e = 0; e < 4; ++e) w[e] = pack_bf16(__uint_as_float(o1[8 * q4 + 2 * e]) * scale, __uint_as_float(o1[8 * q4 + 2 * e + 1]) * scale);
            *reinterpret_cast<uint4*>(tQ + TC_SLAB + off) = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      }
      fence_proxy_async_smem();
      asm volatile("bar.sync 1, %0;" ::"n"(NEW) : "memory");
      if (threadIdx.x == 0) {
        tma_store_3d(&map_dq, tQ, h * 64, 0, s);
        tma_store_3d(&map_dkv, kv + TCL_KV_SLOT, d + h * 64, TC_ROWS, s);
        tma_store_3d(&map_dkv, kv + TCL_KV_SLOT + TC_SLAB, 2 * d + h * 64, TC_ROWS, s);
        bulk_commit();
        if (PF) {
          release_pending = true;  // waited for after the next problem's first step
          prev_qs = qs;
        } else {
          bulk_wait_read<0>();     // the tiles have been read: the next problem may be loaded over them
          mbar_arrive(&bar_kvempty[1]);
          mbar_arrive(&bar_qempty[0]);
        }
      }
    }
    if (threadIdx.x == 0) bulk_wait<0>();
  }
  __syncthreads();
  if (warp == EW + 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512u);
  }
}

static int g_tc_mode = -1;
static int tc_enabled() {
  if (g_tc_mode < 0) {
    const char* e = getenv("MUDPT_ATTN_TC");
    g_tc_mode = e ? atoi(e) : 1;
  }
  return g_tc_mode;
}
void attention_tc_set_mode(int mode) { g_tc_mode = mode; }

// Which sequences take the tcgen05 path (MUDPT_ATTN_TC: 0 = none, 1 = default, 2 = every L <= 256)
bool attention_tc_fwd_eligible(int L, bool causal) {
  const int en = tc_enabled();
  if (en == 0 || L > 256 || L < 1) return false;
  if (en == 2) return true;
  return !causal && L > 128;  // the vision tower; short causal sequences keep the single-pass warp-MMA kernels
}

const char* attention_tc_fwd(const bf16* qkv, bf16* o, float* lse2, int S, int L, int H, int d, bool causal,
                             cudaStream_t stream) {
  const int Lk = (L + 15) & ~15;
  CUtensorMap mq, mkv, mo;
  const char* e;
  const long long ld = 3LL * d;
  if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, TC_ROWS, &mq))) return e;
  if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, Lk, &mkv))) return e;
  if ((e = tensor_map_3d_bf16(o, d, L, S, d, static_cast<long long>(d) * L, TC_ROWS, &mo))) return e;
  const int smem = tc_smem_bytes(Lk);
  const int tmem_cols = Lk <= 64 ? 64 : Lk <= 128 ? 128 : 256;
  const float sl2 = 0.125f * 1.4426950408889634f;
  static bool attr_done[2] = {false, false};
  auto kern = causal ? attn_tc_fwd_kernel<true> : attn_tc_fwd_kernel<false>;
  if (!attr_done[causal ? 1 : 0]) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc_smem_bytes(256)) != cudaSuccess)
      return "attention (tcgen05): cudaFuncSetAttribute failed";
    attr_done[causal ? 1 : 0] = true;
  }
  launch_pdl(kern, dim3((L + TC_ROWS - 1) / TC_ROWS, S * H), dim3(288), static_cast<size_t>(smem), stream, mq, mkv, mo, lse2, L, H, d,
             Lk, tmem_cols, sl2);
  count_launch(1);
  return launch_status("attention fwd (tcgen05) launch failed");
}

// The round-based tcgen05 backward for longer sequences is parity-tested on every shape but, at 199 tokens, bound by its
// per-round hand-over chain (score MMA -> tcgen05.ld -> element-wise -> slab -> output MMA, two CTAs per SM by TMEM):
// 82 us per vision layer against 75 us for the warp-MMA pair (profiles/r02_attention_tc.txt): on request only (mode 2).
bool attention_tc_bwd_eligible(int L, bool causal) {
  (void)causal;
  const int en = tc_enabled();
  if (en == 0 || L < 1) return false;
  if (en == 2) return true;
  // default: the two-problems-in-flight persistent kernel where it wins -- 65..80-token sequences (the 77-token text
  // tower: 141 us per layer against 171 us for the warp-MMA kernel, 96 us of Q / K / V / dO / dQKV traffic)
  // ... and the resident-problem kernel for two-block sequences (the 199-token vision tower)
  return (L > 64 && L <= 80) || (L > TC_ROWS && L <= 2 * TC_ROWS);
}

const char* attention_tc_bwd(const bf16* qkv, const bf16* o, const bf16* d_o, const float* lse2, float* dsum, bf16* dqkv,
                             int S, int L, int H, int d, bool causal, cudaStream_t stream) {
  const char* e;
  const long long ld = 3LL * d;
  if (L <= TC_ROWS) {  // single-block sequences: the persistent kernel
    const int Lb = (L + 15) & ~15;
    CUtensorMap mq, mdo, mo;
    if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, Lb, &mq))) return e;
    if ((e = tensor_map_3d_bf16(d_o, d, L, S, d, static_cast<long long>(d) * L, Lb, &mdo))) return e;
    if ((e = tensor_map_3d_bf16(dqkv, 3 * d, L, S, ld, ld * L, Lb, &mo))) return e;
    const bool pp = Lb <= 80;  // two problems in flight per SM
    const int smem = pp ? tcp_smem_bytes(Lb) : tcs_smem_bytes(Lb);
    static int n_sms = 0;
    if (n_sms == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sms, cudaDevAttrMultiProcessorCount, dev);
      if (n_sms <= 0) n_sms = 148;
    }
    const int n_items = S * H;
    const int grid = n_items < n_sms ? n_items : n_sms;
    const float scale_s = 0.125f, sl2_s = 0.125f * 1.4426950408889634f;
    auto kern = pp ? (causal ? attn_tc_bwd_pp_kernel<true> : attn_tc_bwd_pp_kernel<false>)
                   : (causal ? attn_tc_bwd_short_kernel<true> : attn_tc_bwd_short_kernel<false>);
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)
      return "attention (tcgen05 backward, short): cudaFuncSetAttribute failed";
    launch_pdl(kern, dim3(grid), dim3(pp ? TCP_THREADS : TCS_THREADS), static_cast<size_t>(smem), stream, mq, mdo, mo, lse2, L, H, d, Lb,
               n_items, scale_s, sl2_s);
    count_launch(1);
    return launch_status("attention bwd (tcgen05, short) launch failed");
  }
  if (L <= 2 * TC_ROWS) {  // two-block sequences: the whole problem resident, one launch
    const int Lb = (L + 15) & ~15;
    CUtensorMap mq, mdo, mkv, mdq, mdkv;
    if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, Lb, &mq))) return e;
    if ((e = tensor_map_3d_bf16(d_o, d, L, S, d, static_cast<long long>(d) * L, Lb, &mdo))) return e;
    if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, TC_ROWS, &mkv))) return e;
    if ((e = tensor_map_3d_bf16(dqkv, 3 * d, L, S, ld, ld * L, Lb, &mdq))) return e;
    if ((e = tensor_map_3d_bf16(dqkv, 3 * d, L, S, ld, ld * L, TC_ROWS, &mdkv))) return e;
    const int smem = tcl_smem_bytes(Lb);
    static int n_sms_l = 0;
    if (n_sms_l == 0) {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&n_sms_l, cudaDevAttrMultiProcessorCount, dev);
      if (n_sms_l <= 0) n_sms_l = 148;
    }
    const int n_items = S * H;
    const int grid = n_items < n_sms_l ? n_items : n_sms_l;
    static const int ew = getenv("MUDPT_ATTN_TCL_WARPS") ? atoi(getenv("MUDPT_ATTN_TCL_WARPS")) : 16;
    auto kern = ew == 8 ? (causal ? attn_tc_bwd_long_kernel<true, 8> : attn_tc_bwd_long_kernel<false, 8>)
                        : (causal ? attn_tc_bwd_long_kernel<true, 16> : attn_tc_bwd_long_kernel<false, 16>);
    static bool attr_long[2] = {false, false};
    if (!attr_long[causal ? 1 : 0]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess)
        return "attention (tcgen05 backward, resident): cudaFuncSetAttribute failed";
      attr_long[causal ? 1 : 0] = true;
    }
    launch_pdl(kern, dim3(grid), dim3(tcl_threads(ew == 8 ? 8 : 16)), static_cast<size_t>(smem), stream, mq, mdo, mkv, mdq, mdkv, o, lse2, L, H, d, Lb,
               n_items, 0.125f, 0.125f * 1.4426950408889634f);
    count_launch(1);
    return launch_status("attention bwd (tcgen05, resident) launch failed");
  }
  CUtensorMap m128, m64, mdo128, mdo64, mout;
  if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, TC_ROWS, &m128))) return e;
  if ((e = tensor_map_3d_bf16(qkv, 3 * d, L, S, ld, ld * L, 64, &m64))) return e;
  if ((e = tensor_map_3d_bf16(d_o, d, L, S, d, static_cast<long long>(d) * L, TC_ROWS, &mdo128))) return e;
  if ((e = tensor_map_3d_bf16(d_o, d, L, S, d, static_cast<long long>(d) * L, 64, &mdo64))) return e;
  if ((e = tensor_map_3d_bf16(dqkv, 3 * d, L, S, ld, ld * L, TC_ROWS, &mout))) return e;
  const int Lpad = (L + 63) & ~63;
  const int smem = tcb_smem_bytes(Lpad);
  if (smem > 227 * 1024) return "attention (tcgen05 backward): sequence too long";
  const float scale = 0.125f, sl2 = 0.125f * 1.4426950408889634f;
  const dim3 grid((L + TC_ROWS - 1) / TC_ROWS, S * H);
#define MUDPT_TCB_LAUNCH(KVF, CF)                                                                                        \
  do {                                                                                                                   \
    auto kern = attn_tc_bwd_kernel<KVF, CF>;                                                                             \
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess)                    \
      return "attention (tcgen05 backward): cudaFuncSetAttribute failed";                                                \
    launch_pdl(kern, grid, dim3(TCB_THREADS), static_cast<size_t>(smem), stream, m128, m64, mdo128, mdo64, mout, o, lse2, dsum, L, \
               H, d, scale, sl2);                                                                                        \
  } while (0)
  if (causal) { MUDPT_TCB_LAUNCH(false, true); MUDPT_TCB_LAUNCH(true, true); }
  else { MUDPT_TCB_LAUNCH(false, false); MUDPT_TCB_LAUNCH(true, false); }
#undef MUDPT_TCB_LAUNCH
  count_launch(2);
  return launch_status("attention bwd (tcgen05) launch failed");
}

}  // namespace mudpt
