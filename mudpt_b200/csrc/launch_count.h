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
// nullptr when the last launch succeeded, else "<what>: <cuda error string>" (thread-local buffer).
inline const char* launch_status(const char* what) {
  const cudaError_t e = cudaPeekAtLastError();
  if (e == cudaSuccess) return nullptr;
  static thread_local char buf[256];
  snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
  return buf;
}
}  // namespace mudpt
