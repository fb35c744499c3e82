// Internal interface of the row kernels (see rowops.cu). All return nullptr on success.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace mudpt {
const char* layernorm_fwd(const float* x, const float* gamma, const float* beta, void* out, bool out_bf16, int M, int d,
                          float eps, cudaStream_t stream);
const char* layernorm_bwd(const void* dy, bool dy_bf16, const float* x, const float* gamma, const float* resid, float* dx,
                          __nv_bfloat16* dx_bf16, int M, int d, float eps, cudaStream_t stream);
const char* splice_fwd(float* x, const float* prompt, int S, int L, int row0, int n, int d, cudaStream_t stream);
size_t splice_bwd_workspace_floats(int n, int d);
const char* splice_bwd(float* dx, __nv_bfloat16* dx_bf16, float* dprompt, float* workspace, int S, int L, int row0, int n,
                       int d, bool zero_rows, cudaStream_t stream);
const char* im2col_bf16(const float* img, __nv_bfloat16* out, int B, int R, int p, int ldo, cudaStream_t stream);
const char* write_cls_rows(float* x, const float* cls, const float* pos, int S, int L, int d, cudaStream_t stream);
const char* add_positional(float* x0, const float* emb, const float* pos, int S, int L, int Lsrc, int d, cudaStream_t stream);
const char* cast_to_bf16(const float* in, __nv_bfloat16* out, size_t n, cudaStream_t stream);
const char* transpose_cast_bf16(const float* in, __nv_bfloat16* out, int rows, int cols, cudaStream_t stream);
}  // namespace mudpt
