// Internal interface of the tcgen05 GEMM (see gemm.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace mudpt {

enum GemmEpilogueMode {
  EPI_BF16 = 0,       // out0(bf16) = acc (+ bias)                          QKV in-proj; dgrad of out-proj
  EPI_F32 = 1,        // out0(f32)  = acc (+ bias)                          dgrad into the LN backward
  EPI_RESID_F32 = 2,  // out0(f32)  = acc + bias + resid(f32)               out-proj / c_proj + residual
  EPI_GELU = 3,       // out0(bf16) = h = acc + bias; out1(bf16) = QuickGELU(h)   c_fc
  EPI_GELU_BWD = 4,   // out0(bf16) = acc * QuickGELU'(aux(bf16))           dgrad of c_proj
  EPI_PATCH = 5,      // out0(f32)[img*L + 1 + p] = acc + resid[1 + p]       conv1 patch embedding
};

struct GemmEpilogue {
  int mode = EPI_BF16;
  void* out0 = nullptr;
  void* out1 = nullptr;
  const float* bias = nullptr;   // [N] or null
  const float* resid = nullptr;  // EPI_RESID_F32: [M, ldc] f32; EPI_PATCH: positional embedding [np+1, ldc]
  const void* aux = nullptr;     // EPI_GELU_BWD: h [M, ldc] bf16
  int ldc = 0;                   // leading dimension (elements) of out0/out1/resid/aux
  int patch_np = 1;              // EPI_PATCH: patches per image
  int patch_L = 1;               // EPI_PATCH: tokens per image (np + 1 + n_ctx)
};

// C = A[M,K] * B[N,K]^T with the fused epilogue. Returns nullptr on success, else a static
// error string. Asynchronous on `stream`.
const char* gemm_bf16_tn(const __nv_bfloat16* A, int lda, const __nv_bfloat16* B, int ldb, const GemmEpilogue& ep,
                         int M, int N, int K, cudaStream_t stream);
void gemm_clear_tensor_map_cache();
#ifdef MUDPT_BRINGUP
void gemm_set_bringup_simt(bool on);
#endif

}  // namespace mudpt
