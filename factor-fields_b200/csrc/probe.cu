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

// ---------------------------------------------------------------------------------------------------------
// tcgen05.mma issue / execution rate as a function of N (M = 128, K = 16, bf16, SWIZZLE_NONE operands in shared memory):
// one elected thread per CTA issues `count` accumulating MMAs back to back, commits and waits; cycles by clock64.
// The MLPs of this path are made of small-N MMAs (N = 32 / 64 / 128), so this — not the dense peak — is their roofline.
// ---------------------------------------------------------------------------------------------------------
#include "tc_common.cuh"
namespace ffb {
__global__ void __launch_bounds__(128) mma_probe_kernel(int n_cols, int count, long long* out) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (uint32_t o = threadIdx.x * 16u; o < 64u * 1024u; o += 128u * 16u) *reinterpret_cast<uint4*>(smem + o) = make_uint4(0, 0, 0, 0);
  if (warp == 0) tmem_alloc(&tmem_slot, 256u);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 0) {
    if (elect_one()) {
      const uint32_t idesc = make_idesc(n_cols, 0, 0);
      const DescBase a = desc_base(smem_u32(smem), 2048u, 128u), b = desc_base(smem_u32(smem) + 32768u, 4096u, 128u);
      const long long t0 = clock64();
      for (int i = 0; i < count; i += 4) {
        umma_f16_c<true>(tmem, desc_at(a, 0), desc_at(b, 0), idesc);
        umma_f16_c<true>(tmem, desc_at(a, 4096), desc_at(b, 8192), idesc);
        umma_f16_c<true>(tmem, desc_at(a, 8192), desc_at(b, 0), idesc);
        umma_f16_c<true>(tmem, desc_at(a, 12288), desc_at(b, 8192), idesc);
      }
      const long long t1 = clock64();
      umma_commit(&bar);
      mbar_wait(&bar, 0);
      const long long t2 = clock64();
      if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    __syncwarp();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256u);
}
}  // namespace ffb

extern "C" int ffb_probe_mma(int32_t n_cols, int32_t count, long long* d_out, void* stream) {
  FFB_REQUIRE(n_cols >= 16 && n_cols <= 256 && (n_cols % 16) == 0 && count > 0 && d_out, "bad argument");
  static ffb::PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(ffb::mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  ffb::mma_probe_kernel<<<ffb::sm_count(), 128, 64 * 1024, (cudaStream_t)stream>>>(n_cols, count, d_out);
  FFB_LAUNCHED();
  return FFB_OK;
}
