// Internal helpers shared by the translation units of libapg_b200.so (not part of the C-ABI).
#pragma once

#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>

#include "apg_b200.h"

namespace apg {

// thread-local error text behind apg_last_error()
std::string& last_error();
int fail(int code, const char* fmt, ...);
void count_launch(int64_t n = 1);

#define APG_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::apg::fail(APG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                         __FILE__, __LINE__);                                                       \
  } while (0)

#define APG_CHECK(cond, code, ...)                  \
  do {                                              \
    if (!(cond)) return ::apg::fail(code, __VA_ARGS__); \
  } while (0)

template <typename T>
inline cudaError_t dev_alloc(T** p, size_t count) {
  if (count == 0) count = 1;
  cudaError_t e = cudaMalloc(reinterpret_cast<void**>(p), count * sizeof(T));
  if (e == cudaSuccess) e = cudaMemset(*p, 0, count * sizeof(T));
  return e;
}

template <typename T>
inline void dev_free(T*& p) {
  if (p) cudaFree(p);
  p = nullptr;
}

inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

}  // namespace apg
