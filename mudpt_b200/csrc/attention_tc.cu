// Blackwell-native attention for the CLIP towers: tcgen05.mma with the score / probability tiles in TMEM, operands
// staged by TMA.  Replaces the scaled_dot_product_attention inside nn.MultiheadAttention (clip/model.py:271-273;
// causal mask of the text tower :810-816) for sequences of up to 256 tokens, head width 64.
//
// Forward, one CTA per (sequence, head, block of 128 query rows), 160 threads:
//   warp 4 (one elected lane)  TMA: Q block [128 x 64], K and V [Lk x 64] (3D descriptors over [S][L][3d]: rows past
//                              the sequence end are zero-filled, never the next sequence's tokens)
//                              S = Q K^T : tcgen05.mma 128 x Lk x 64, fp32 scores in TMEM columns [0, Lk)
//                              O = P V   : tcgen05.mma 128 x 64 x Lk, V as the MN-major B operand (no transpose),
//                                          P from shared memory, fp32 output in TMEM columns [0, 64) (S is dead)
//   warps 0-3 (thread = row)   tcgen05.ld of the thread's own score row in 16-column pieces: row max, then
//                              p = exp2(s c - m c), row sum -- no shuffles, no cross-thread reduction; P as bf16 into
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

__host__ __device__ inline int tc_smem_bytes(int Lk) {
  const int kv = Lk * 128;
  const int nslab = (Lk + 63) / 64;
  int extra = nslab - 1 - (kv >= TC_SLAB ? 1 : 0);  // slab 0 = Q tile, slab 1 = K tile when it is large enough
  if (extra < 0) extra = 0;
  return TC_SLAB + 2 * kv + extra * TC_SLAB + 256 /*barriers*/ + 1024 /*alignment slack*/;
}

template <bool CAUSAL>
__global__ void __launch_bounds__(160) attn_tc_fwd_kernel(const __grid_constant__ CUtensorMap map_q,
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
  uint64_t* bars = reinterpret_cast<uint64_t*>(sX + extra * TC_SLAB);
  uint64_t* bar_load = bars;      // TMA bytes
  uint64_t* bar_s = bars + 1;     // scores in TMEM
  uint64_t* bar_p = bars + 2;     // P in shared memory (128 arrivals)
  uint64_t* bar_o = bars + 3;     // output in TMEM
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  if (warp == 4) {
    if (lane == 0) {
      tma_prefetch_desc(&map_q);
      tma_prefetch_desc(&map_kv);
      mbar_init(bar_load, 1);
      mbar_init(bar_s, 1);
      mbar_init(bar_p, TC_ROWS);
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

  if (warp == 4) {
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
    // ===================== softmax / epilogue: thread = query row =====================
    const int t = warp * 32 + lane;           // row inside the block = TMEM lane
    const int row = mb * TC_ROWS + t;         // token index of the row
    const uint32_t trow = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    // keys this row sees (rows past the sequence end compute on zero-filled Q and are never stored)
    const int nvis = CAUSAL ? (row < L ? row + 1 : L) : L;
    const int n16 = Lk >> 4;
    mbar_wait(bar_s, 0);
    tc_fence_after();
    float mx = -INFINITY;
    for (int j = 0; j < n16; ++j) {
      uint32_t r[16];
      tmem_ld_32x16(trow + static_cast<uint32_t>(j * 16), r);
      tmem_ld_wait_regs16(r);
      if (j * 16 + 16 <= nvis) {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) mx = fmaxf(mx, j * 16 + i < nvis ? __uint_as_float(r[i]) : -INFINITY);
      }
    }
    const float m2 = mx * scale_log2e;  // scaled max, log2 domain (nvis >= 1: finite)
    float l = 0.f;
    const f32x2 sc2 = f2_pack(scale_log2e, scale_log2e), nm2 = f2_pack(-m2, -m2);
    for (int j = 0; j < n16; ++j) {
      uint32_t r[16];
      tmem_ld_32x16(trow + static_cast<uint32_t>(j * 16), r);
      tmem_ld_wait_regs16(r);
      uint32_t pk[8];
      const bool full = j * 16 + 16 <= nvis;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float a, b;
        f2_unpack(f2_fma(f2_pack_u(r[2 * i], r[2 * i + 1]), sc2, nm2), a, b);
        a = exp2f(a);
        b = exp2f(b);
        if (!full) {
          a = j * 16 + 2 * i < nvis ? a : 0.f;
          b = j * 16 + 2 * i + 1 < nvis ? b : 0.f;
        }
        l += a + b;
        pk[i] = pack_bf16(a, b);
      }
      // 16 keys = two 16 B chunks of this row in slab j / 4 (64 keys per slab)
      uint8_t* ps = slab(j >> 2) + t * 128;
      const int c = (j & 3) * 2;
      *reinterpret_cast<uint4*>(ps + (((c) ^ (t & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(ps + (((c + 1) ^ (t & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    tc_fence_before();          // every tcgen05.ld of the scores has completed: the MMA may overwrite them with O
    fence_proxy_async_smem();   // P is visible to the tensor core's shared-memory reads
    mbar_arrive(bar_p);
    if (row < L) lse2[(static_cast<size_t>(s) * H + h) * L + row] = m2 + log2f(l);
    const float inv = 1.f / l;
    mbar_wait(bar_o, 0);
    tc_fence_after();
    // O / l as bf16 into the (dead) Q tile: 128 rows x 128 B, 128B swizzle, then one TMA store
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint32_t r[16];
      tmem_ld_32x16(trow + static_cast<uint32_t>(j * 16), r);
      tmem_ld_wait_regs16(r);
      uint32_t pk[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) pk[i] = pack_bf16(__uint_as_float(r[2 * i]) * inv, __uint_as_float(r[2 * i + 1]) * inv);
      uint8_t* po = sQ + t * 128;
      *reinterpret_cast<uint4*>(po + (((2 * j) ^ (t & 7)) << 4)) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      *reinterpret_cast<uint4*>(po + (((2 * j + 1) ^ (t & 7)) << 4)) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
    }
    tc_fence_before();
    fence_proxy_async_smem();
    asm volatile("bar.sync 1, 128;" ::: "memory");  // the four softmax warps
    if (t == 0) {
      tma_store_3d(&map_o, sQ, h * 64, mb * TC_ROWS, s);
      bulk_commit();
      bulk_wait<0>();
    }
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, static_cast<uint32_t>(tmem_cols));
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
  launch_pdl(kern, dim3((L + TC_ROWS - 1) / TC_ROWS, S * H), dim3(160), static_cast<size_t>(smem), stream, mq, mkv, mo, lse2, L, H, d,
             Lk, tmem_cols, sl2);
  count_launch(1);
  return launch_status("attention fwd (tcgen05) launch failed");
}

}  // namespace mudpt
