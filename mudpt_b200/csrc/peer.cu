// Own collectives over NVLink peer memory for the logits head's exchange: the ranks' text-feature shards
// [C / G, e] and their gradients [C, e] live in symmetric allocations that every GPU of the node maps (the caller passes the
// device array of peer base pointers; torch.distributed._symmetric_memory supplies it and the stream-ordered barrier
// that precedes each call).  Instead of an NCCL all-gather / reduce-scatter launch, each rank PULLS what it needs with
// plain loads over NVLink:
//   gather:          out[c, :] = peers[owner(c)][c - lo(owner), :]                (2 MB at C = 1000, e = 512)
//   reduce-scatter:  out[i, :] = sum over ranks r (fixed order) of peers[r][lo(me) + i, :]
// Row shards follow mudpt_b200/dist.py:shard_bounds (the first n % world ranks hold one extra row).
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch_count.h"
#include "peer.h"

namespace mudpt {

__device__ __forceinline__ void shard_of_row(int row, int n_total, int world, int& r, int& li) {
  const int q = n_total / world, rem = n_total - q * world;
  const int b = rem * (q + 1);
  if (row < b) {
    r = row / (q + 1);
    li = row - r * (q + 1);
  } else {
    r = rem + (row - b) / q;
    li = (row - b) - (r - rem) * q;
  }
}

__global__ void __launch_bounds__(128) peer_gather_rows_kernel(const float* const* __restrict__ peers, int world, int n_total,
                                                               int width4, float4* __restrict__ out) {
  const int row = blockIdx.x;
  int r, li;
  shard_of_row(row, n_total, world, r, li);
  const float4* src = reinterpret_cast<const float4*>(peers[r]) + static_cast<size_t>(li) * width4;
  for (int t = threadIdx.x; t < width4; t += blockDim.x) out[static_cast<size_t>(row) * width4 + t] = __ldcg(src + t);
}

__global__ void __launch_bounds__(128) peer_reduce_rows_kernel(const float* const* __restrict__ peers, int world, int lo,
                                                               int width4, float4* __restrict__ out) {
  const int li = blockIdx.x;
  const size_t off = static_cast<size_t>(lo + li) * width4;
  for (int t = threadIdx.x; t < width4; t += blockDim.x) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < world; ++r) {  // fixed order: the sum is deterministic and identical to a rank-ordered reduction
      const float4 v = __ldcg(reinterpret_cast<const float4*>(peers[r]) + off + t);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    out[static_cast<size_t>(li) * width4 + t] = acc;
  }
}

const char* peer_gather_rows(const float* const* peers_dev, int world, int n_total, int width, float* out, cudaStream_t stream) {
  if (world < 1 || n_total < world || width % 4 != 0) return "peer gather: bad geometry";
  peer_gather_rows_kernel<<<n_total, 128, 0, stream>>>(peers_dev, world, n_total, width / 4, reinterpret_cast<float4*>(out));
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? nullptr : "peer gather launch failed";
}

const char* peer_reduce_scatter_rows(const float* const* peers_dev, int world, int rank, int n_total, int width, float* out,
                                     cudaStream_t stream) {
  if (world < 1 || rank < 0 || rank >= world || n_total < world || width % 4 != 0) return "peer reduce-scatter: bad geometry";
  const int q = n_total / world, rem = n_total - q * world;
  const int lo = rank * q + (rank < rem ? rank : rem), rows = q + (rank < rem ? 1 : 0);
  peer_reduce_rows_kernel<<<rows, 128, 0, stream>>>(peers_dev, world, lo, width / 4, reinterpret_cast<float4*>(out));
  count_launch(1);
  return cudaGetLastError() == cudaSuccess ? nullptr : "peer reduce-scatter launch failed";
}

}  // namespace mudpt
