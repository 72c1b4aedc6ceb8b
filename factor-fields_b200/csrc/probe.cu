// probe.cu — measurement probes (bench.py): the L2 reduction throughput that bounds the scatter kernels.
//
// The field backward pass is a stream of red.global.add.v2/v4.f32 into L2-resident gradient grids; HBM bandwidth is the
// wrong yardstick for it (ncu: DRAM 12 %, L2 reduction sectors the busiest unit).  This probe issues the same instruction
// against an L2-resident buffer of the gradient arena's size and reports how many 16-byte vector reductions per second
// the memory system retires, for the two address patterns that bracket the real kernel:
//   pattern 0: every lane a random 16-byte slot (one 32-byte sector per lane — the fine basis levels);
//   pattern 1: a warp covers 512 contiguous bytes (16 sectors per instruction — the run-aggregated coefficient rows).
#include "ffb_common.cuh"

namespace ffb {

__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

__global__ void __launch_bounds__(256) red_probe_kernel(float* __restrict__ buf, uint32_t n_slots, int iters, int pattern) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31, warp = tid >> 5;
  for (int it = 0; it < iters; ++it) {
    uint32_t slot;
    if (pattern == 0) slot = mix32(tid * 0x9e3779b9u + (uint32_t)it * 0x85ebca6bu) % n_slots;
    else slot = ((mix32(warp * 0x9e3779b9u + (uint32_t)it) % (n_slots / 32)) * 32 + lane) % n_slots;
    float* p = buf + (size_t)slot * 4;
    asm volatile("red.global.add.v4.f32 [%0], {%1, %1, %1, %1};" ::"l"(p), "f"(1.0f) : "memory");
  }
}

}  // namespace ffb

extern "C" {

/* Issues blocks*256*iters 16-byte vector reductions into buf[0 .. n_floats) (n_floats a multiple of 128, 16-byte aligned).
 * Time it with CUDA events on `stream`; *n_ops_out receives the number of reductions issued. */
int ffb_probe_red(float* buf, int64_t n_floats, int32_t blocks, int32_t iters, int32_t pattern, int64_t* n_ops_out, void* stream) {
  FFB_REQUIRE(buf && n_floats >= 128 && (n_floats % 128) == 0 && blocks > 0 && iters > 0, "bad argument");
  FFB_REQUIRE(((uintptr_t)buf & 15) == 0, "buffer must be 16-byte aligned");
  ffb::red_probe_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(buf, (uint32_t)(n_floats / 4), iters, pattern);
  FFB_LAUNCHED();
  if (n_ops_out) *n_ops_out = (int64_t)blocks * 256 * iters;
  return FFB_OK;
}

}
