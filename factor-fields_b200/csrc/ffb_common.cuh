// ffb_common.cuh — error plumbing and launch helpers shared by the .cu files of libffb200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>
#include "../../include/ffb200.h"

namespace ffb {
void set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_launches;
int sm_count();             // of the current device (cached per device)
int smem_optin_bytes();     // cudaDevAttrMaxSharedMemoryPerBlockOptin of the current device (cached per device)
int current_device();
extern int g_deterministic;   // knob "field_deterministic": bit-reproducible (serial) scatter-add for debugging / gradient tests
void keep_pool_cached();

// Function attributes (cudaFuncSetAttribute) are per device: `static PerDeviceOnce once; if (once.first()) { ... }`
struct PerDeviceOnce {
  bool done[64] = {};
  bool first() {
    const int d = current_device();
    if (d < 0 || d >= 64) return true;
    if (done[d]) return false;
    done[d] = true;
    return true;
  }
};

inline int check_cuda(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return FFB_OK;
  set_error("%s: %s", what, cudaGetErrorString(e));
  return FFB_ECUDA;
}

#define FFB_CUDA(call)                                          \
  do {                                                          \
    int _rc = ::ffb::check_cuda((call), #call);                 \
    if (_rc != FFB_OK) return _rc;                              \
  } while (0)

#define FFB_REQUIRE(cond, msg)                                  \
  do {                                                          \
    if (!(cond)) {                                              \
      ::ffb::set_error("%s: %s", __func__, msg);                \
      return FFB_EINVAL;                                        \
    }                                                           \
  } while (0)

// call after every kernel launch
#define FFB_LAUNCHED()                                          \
  do {                                                          \
    ::ffb::g_launches.fetch_add(1, std::memory_order_relaxed);  \
    FFB_CUDA(cudaGetLastError());                               \
  } while (0)

__device__ __forceinline__ int64_t resolve_n(int64_t n, const int32_t* n_dev) {
  if (n_dev) {
    int64_t m = (int64_t)(*n_dev);
    return m < n ? m : n;
  }
  return n;
}

inline unsigned blocks_for(int64_t n, int threads, int64_t cap = (1 << 30)) {
  int64_t b = (n + threads - 1) / threads;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (unsigned)b;
}
}  // namespace ffb
