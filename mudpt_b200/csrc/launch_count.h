// Global count of kernel launches issued by the library (bench.py reports it as gpu_launches).
#pragma once
#include <atomic>
namespace mudpt {
extern std::atomic<long long> g_launch_counter;
inline void count_launch(int n = 1) { g_launch_counter.fetch_add(n, std::memory_order_relaxed); }
}  // namespace mudpt
