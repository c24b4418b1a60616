#pragma once
#include <cstdarg>
#include <cstdint>
#include <cuda_runtime.h>

#include <atomic>

#ifndef OPADPO_EINVAL
#define OPADPO_OK 0
#define OPADPO_EINVAL (-1)
#define OPADPO_ECUDA (-2)
#define OPADPO_ENOMEM (-3)
#define OPADPO_ESTATE (-4)
#define OPADPO_ENOENT (-5)
#endif

namespace opadpo {
extern thread_local char g_last_error[1024];
void set_error(const char* fmt, ...);
// every kernel launch of this library is counted (bench.py reports it as gpu_launches)
extern std::atomic<int64_t> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace opadpo

#define OPADPO_CUDA_CHECK(expr)                                                                     \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::opadpo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return OPADPO_ECUDA;                                                                          \
    }                                                                                               \
  } while (0)
