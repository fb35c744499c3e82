// Global count of kernel launches issued by the library (bench.py reports it as gpu_launches).
#pragma once
#include <atomic>
namespace mudpt {
extern std::atomic<long long> g_launch_counter;
inline void count_launch(int n = 1) { g_launch_counter.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mudpt

#include <cuda_runtime.h>
#include <cstdio>
namespace mudpt {
// Launch with programmatic stream serialization (see pdl_wait() in common.cuh).  Only for kernels that
// call pdl_wait() before their first global-memory access.
template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);  // errors surface through launch_status()
}

// nullptr when the last launch succeeded, else "<what>: <cuda error string>" (thread-local buffer).
inline const char* launch_status(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return nullptr;
  static thread_local char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  return buf;
}
}  // namespace mudpt
