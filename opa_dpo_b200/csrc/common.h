#pragma once
#include <cstdarg>
#include <cstdint>
#include <cuda_runtime.h>

namespace opadpo {
extern thread_local char g_last_error[1024];
void set_error(const char* fmt, ...);
}  // namespace opadpo

#define OPADPO_CUDA_CHECK(expr)                                                                     \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess) {                                                                        \
      ::opadpo::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));    \
      return OPADPO_ECUDA;                                                                          \
    }                                                                                               \
  } while (0)
