// Shared host-side helpers for the C-ABI translation units (error text, CUDA checks).
#pragma once
#include <nvtx3/nvToolsExt.h>
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

// NVTX range over a host-side enqueue section (header-only NVTX v3: the injection library is looked up at run time, nothing is
// linked).  Ranges name the stages in an nsys / ncu timeline: cmos.orb.pyramid, cmos.ba.lm_iteration, ...
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

// Per-stage device timing with CUDA events recorded on the launching stream (bench.py's roofline reads this).
// When disabled nothing is recorded.  Events of up to kRing calls are kept and folded into the totals when the
// caller asks for them (which synchronises the events, never the launches).
struct StageTimer {
  static constexpr int kMaxStages = 12, kRing = 128;
  bool enabled = false;
  int n_stages = 0, used = 0, cur = 0;
  cudaEvent_t ev[kRing][kMaxStages + 1] = {};
  int marks[kRing] = {};
  double total_ms[kMaxStages] = {};
  long long calls = 0;

  void fold() {
    for (int c = 0; c < used; c++) {
      if (marks[c] < 2) continue;
      cudaEventSynchronize(ev[c][marks[c] - 1]);
      for (int s = 0; s + 1 < marks[c]; s++) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[c][s], ev[c][s + 1]) == cudaSuccess) total_ms[s] += ms;
      }
      n_stages = marks[c] - 1 > n_stages ? marks[c] - 1 : n_stages;
      calls++;
    }
    used = 0;
  }
  void begin(cudaStream_t st) {
    if (!enabled) return;
    if (used == kRing) fold();
    cur = used++;
    marks[cur] = 0;
    mark(st);
  }
  void mark(cudaStream_t st) {
    if (!enabled || marks[cur] > kMaxStages) return;
    cudaEvent_t& e = ev[cur][marks[cur]];
    if (!e) cudaEventCreate(&e);
    cudaEventRecord(e, st);
    marks[cur]++;
  }
  void reset() {
    fold();
    for (double& t : total_ms) t = 0.0;
    calls = 0;
    n_stages = 0;
  }
  void destroy() {
    for (auto& row : ev)
      for (cudaEvent_t& e : row)
        if (e) { cudaEventDestroy(e); e = nullptr; }
  }
};

}  // namespace cmos
