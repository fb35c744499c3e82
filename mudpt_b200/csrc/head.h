// Internal interface of the head kernels (see head.cu). All return nullptr on success.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>

namespace mudpt {
size_t feature_head_workspace_floats(int S, int d, int e);
// ws: feature_head_workspace_floats(S, d, e) floats (gathered + normalised rows / their gradient, split-K partials)
const char* feature_head_fwd(const float* x, const int* rows, const float* gamma, const float* beta, const float* proj,
                             float* f, float* ws, int S, int L, int d, int e, float eps, cudaStream_t stream);
const char* feature_head_bwd(const float* df, const float* x, const int* rows, const float* gamma, const float* proj,
                             float* dx, __nv_bfloat16* dx_bf16, float* ws, int S, int L, int d, int e, float eps,
                             cudaStream_t stream);
size_t logits_head_workspace_floats(int B, int C, int e);
const char* logits_head(const float* f_img, const float* f_txt, const long long* labels, float scale, int B, int C, int e,
                        float inv_global_batch, float* ws, float* logits, float* loss, float* d_f_img, float* d_f_txt,
                        cudaStream_t stream);
const char* logits_head_bwd(const float* f_img, const float* f_txt, const float* dlogits, float scale, int B, int C, int e,
                            float* ws, float* d_f_img, float* d_f_txt, cudaStream_t stream);

// fused multi-tensor SGD (torch.optim.SGD semantics); pointer arrays live on the host
static constexpr int SGD_MAX_TENSORS = 32;
const char* sgd_step(void* const* params, const void* const* grads, void* const* bufs, const long long* numel, int n, float lr,
                     float momentum, float dampening, float weight_decay, bool nesterov, bool first_step, cudaStream_t stream);

}  // namespace mudpt
