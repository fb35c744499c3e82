// Internal interface of the prompt-algebra kernels (see prompt.cu).
#pragma once
#include <cuda_runtime.h>

namespace mudpt {

// Plain pointers into the caller's fp32 tensors (all contiguous).  Shapes: n = n_ctx, D = depth, dt / dv = widths.
struct PromptArgs {
  int n, depth, dt, dv;
  float eps;
  // trainable (forward inputs)
  const float *ctx, *deep, *We, *be, *Wd, *bd, *vctx, *vdeep, *Wv, *bv;
  // frozen
  const float *ln_g, *ln_b, *pos;  // ln_pre [dv]; positional_embedding[1:1+n] [n, dt]
  // forward outputs / saved
  float *P_v, *P_t;   // [D, n, dv], [D, n, dt]
  float* ln_in;       // [n, dv] = visual_ctx + shared_ctx (saved for the backward)
  // backward inputs / outputs
  const float *dP_v, *dP_t;
  float* u;           // [n, dv] scratch
  float *d_ctx, *d_deep, *d_We, *d_be, *d_Wd, *d_bd, *d_vctx, *d_vdeep, *d_Wv, *d_bv;
};

const char* prompt_forward(const PromptArgs& a, cudaStream_t stream);
const char* prompt_backward(const PromptArgs& a, cudaStream_t stream);

}  // namespace mudpt
