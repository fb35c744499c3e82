// Internal interface of the input-pipeline kernels (see augment.cu). Return nullptr on success.
#pragma once
#include <cuda_runtime.h>

#include "../../include/mudpt_b200.h"

namespace mudpt {
// bytes of device workspace augment_images needs for this batch; -1 and *err set on a bad descriptor
long long augment_workspace_bytes(const mudpt_image_desc* descs_host, int n, int out_h, int out_w, const char** err);
const char* augment_images(const mudpt_image_desc* descs, const mudpt_image_desc* descs_host, int n, int out_h, int out_w,
                           const float* mean_host, const float* std_host, void* workspace, long long ws_bytes, float* out,
                           cudaStream_t stream);
}  // namespace mudpt
