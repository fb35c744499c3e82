// Internal interface of the peer-memory collectives (see peer.cu).  nullptr on success.
#pragma once
#include <cuda_runtime.h>

namespace mudpt {
const char* peer_gather_rows(const float* const* peers_dev, int world, int n_total, int width, float* out, cudaStream_t stream);
const char* peer_reduce_scatter_rows(const float* const* peers_dev, int world, int rank, int n_total, int width, float* out,
                                     cudaStream_t stream);
}  // namespace mudpt
