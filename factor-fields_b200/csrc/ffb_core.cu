// ffb_core.cu — error state, device info, launch counter.
#include <stdarg.h>
#include <string.h>
#include "ffb_common.cuh"

namespace ffb {
static thread_local char g_err[512] = "";
std::atomic<uint64_t> g_launches{0};
int g_deterministic = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int current_device() {
  int dev = 0;
  return cudaGetDevice(&dev) == cudaSuccess ? dev : -1;
}

static int cached_attr(int (&cache)[64], cudaDeviceAttr attr, int fallback) {
  const int dev = current_device();
  if (dev < 0) return fallback;
  if (dev < 64 && cache[dev]) return cache[dev];
  int v = 0;
  if (cudaDeviceGetAttribute(&v, attr, dev) != cudaSuccess || v <= 0) return fallback;
  if (dev < 64) cache[dev] = v;
  return v;
}

int sm_count() {
  static int cache[64] = {};
  return cached_attr(cache, cudaDevAttrMultiProcessorCount, 148);
}

int smem_optin_bytes() {
  static int cache[64] = {};
  return cached_attr(cache, cudaDevAttrMaxSharedMemoryPerBlockOptin, 48 * 1024);
}
// Stream-ordered allocations (generic field backward, grid_mapping) come from the device's default memory pool; keep
// freed blocks cached across synchronisation points instead of returning them to the driver (default threshold 0),
// otherwise every host sync makes the next cudaMallocAsync a multi-millisecond driver call.
void keep_pool_cached() {
  static bool done[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
  done[dev] = true;
}
}  // namespace ffb

extern "C" {
const char* ffb_last_error(void) { return ffb::g_err; }
int ffb_abi_version(void) { return FFB_ABI_VERSION; }
uint64_t ffb_launch_count(void) { return ffb::g_launches.load(); }

int ffb_device_info(int* sm, int* major, int* minor) {
  int dev = 0;
  FFB_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  FFB_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm) *sm = p.multiProcessorCount;
  if (major) *major = p.major;
  if (minor) *minor = p.minor;
  return FFB_OK;
}
}
