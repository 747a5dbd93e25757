// Shared host-side helpers for the C-ABI translation units (error text, CUDA checks).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdio>

#include "../../include/cmos_b200.h"

namespace cmos {

void set_error(const char* fmt, ...);

#define CMOS_CUDA_OK(expr)                                                                      \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) {                                                                    \
      cmos::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CMOS_ERR_CUDA;                                                                     \
    }                                                                                           \
  } while (0)

#define CMOS_REQUIRE(cond, ...)              \
  do {                                       \
    if (!(cond)) {                           \
      cmos::set_error(__VA_ARGS__);          \
      return CMOS_ERR_ARG;                   \
    }                                        \
  } while (0)

template <typename T>
static inline T* dev_alloc(size_t n, cudaError_t* err) {
  void* p = nullptr;
  size_t bytes = n * sizeof(T);
  cudaError_t e = cudaMalloc(&p, bytes > 0 ? bytes : 1);
  if (e != cudaSuccess && err) *err = e;
  return (T*)p;
}

}  // namespace cmos
