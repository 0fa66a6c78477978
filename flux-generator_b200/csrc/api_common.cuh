// Host-side plumbing shared by every C-ABI entry point: status codes, thread-local error text,
// launch counter, checked launches.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/flux_b200.h"

namespace fx {

extern thread_local char g_err[512];
extern std::atomic<uint64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// call right after a kernel launch
inline int launched(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaPeekAtLastError();
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(FX_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  }
  return FX_OK;
}

#define FX_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return ::fx::fail(FX_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define FX_CUDA(expr)                                                                   \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess) return ::fx::fail(FX_ERR_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
inline int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

}  // namespace fx
