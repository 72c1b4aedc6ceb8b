// allreduce.cu — in-place sum all-reduce of the flat fp32 gradient arena over NVLink 5 / NVSwitch peer memory, as ONE kernel
// on the training stream (so the whole data-parallel step stays a single CUDA graph: no NCCL call between two graphs).
//
// The arena lives in symmetric memory (every rank maps every peer's buffer, and — where the fabric offers it — one multicast
// address for all of them).  Rank r owns the r-th slice of the arena:
//   barrier (all ranks have finished writing their gradients)
//   NVLS path : v = multimem.ld_reduce.add [mc + i]   (the SWITCH sums the eight copies; one 16-byte load per element group)
//               multimem.st [mc + i] = v               (the switch writes the sum back into all eight copies)
//   P2P path  : v = sum over peers of ld [peer_p + i] (fixed rank order), then st [peer_p + i] = v for every peer
//   barrier (every slice has been written everywhere)
// Each element is reduced once, by its owner, so all ranks end with bit-identical sums.  The two cross-GPU barriers are run
// by CTA 0 (one 32-bit slot per peer in the ranks' signal pads, release / acquire at system scope) and fanned out to the other
// CTAs through a local flag; the epoch counter lives in device memory and only grows, so there is nothing to reset and a
// captured graph replays correctly.
#include "ffb_common.cuh"

namespace ffb {

struct ArArgs {
  float* local;                    // this rank's arena (symmetric buffer base + offset)
  const uint64_t* peers;           // device array [world]: every rank's arena address as mapped here
  uint64_t mc;                     // multicast address of the arena (0: no multicast -> P2P path)
  const uint64_t* pads;            // device array [world]: every rank's signal pad as mapped here
  uint32_t* epoch;                 // device: [0] barriers completed so far, [1] CTAs done storing (this launch), [2] go flag
  int64_t n4;                      // arena size in float4
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Cross-rank barrier, executed by CTA 0 only: slot r of rank p's pad is written only by rank r.
__device__ __forceinline__ void rank_barrier(const ArArgs& a, uint32_t value) {
  if ((int)threadIdx.x < a.world) {
    const int p = threadIdx.x;
    st_release_sys(reinterpret_cast<uint32_t*>(a.pads[p]) + a.rank, value);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.pads[a.rank]) + p;
    while ((int32_t)(ld_acquire_sys(mine) - value) < 0) {
    }
  }
  __syncthreads();
}

// epoch[0]: barriers completed before this launch; epoch[1]: CTAs of this launch done with their stores;
// epoch[2]: "go" flag — CTA 0 publishes the epoch of the opening barrier here for the other CTAs of the grid.
template <bool NVLS>
__global__ void __launch_bounds__(512) allreduce_symm_kernel(const ArArgs a) {
  volatile uint32_t* ep = a.epoch;
  const uint32_t e0 = ep[0];                 // same value in every CTA: bumped only at the very end of a launch
  if (blockIdx.x == 0) {
    rank_barrier(a, e0 + 1);                 // every rank has finished writing its gradients (stream order + this barrier)
    if (threadIdx.x == 0) {
      __threadfence();
      ep[2] = e0 + 1;
    }
  } else {
    if (threadIdx.x == 0)
      while ((int32_t)(ep[2] - (e0 + 1)) < 0) {
      }
    __syncthreads();
  }
  const int64_t per = (a.n4 + a.world - 1) / a.world;
  const int64_t lo = per * a.rank, hi = lo + per < a.n4 ? lo + per : a.n4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int U = 4;
  for (int64_t i0 = lo + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < hi; i0 += U * stride) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= hi) break;
      if (NVLS) {
        const float4* src = reinterpret_cast<const float4*>(a.mc) + i;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[u].x), "=f"(v[u].y), "=f"(v[u].z), "=f"(v[u].w)
                     : "l"(src)
                     : "memory");
      } else {
        v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = 0; p < a.world; ++p) {
          const float4 t = __ldcv(reinterpret_cast<const float4*>(a.peers[p]) + i);      // volatile: never a stale cached line
          v[u].x += t.x; v[u].y += t.y; v[u].z += t.z; v[u].w += t.w;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t i = i0 + u * stride;
      if (i >= hi) break;
      if (NVLS) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<float4*>(a.mc) + i), "f"(v[u].x),
                     "f"(v[u].y), "f"(v[u].z), "f"(v[u].w)
                     : "memory");
      } else {
        for (int p = 0; p < a.world; ++p) __stcg(reinterpret_cast<float4*>(a.peers[p]) + i, v[u]);
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (blockIdx.x != 0) {
    if (threadIdx.x == 0) atomicAdd(a.epoch + 1, 1u);
    return;
  }
  if (threadIdx.x == 0)
    while (ep[1] < gridDim.x - 1) {          // every other CTA of this rank has stored (and fenced) its part
    }
  __syncthreads();
  __threadfence_system();
  rank_barrier(a, e0 + 2);                   // every slice has been written on every rank
  if (threadIdx.x == 0) {
    ep[1] = 0;
    __threadfence();
    ep[0] = e0 + 2;
  }
}

}  // namespace ffb

using namespace ffb;

extern "C" {

/* In-place sum all-reduce of local[0 .. n_floats) across `world` ranks whose arenas are mapped at d_peer_ptrs[0 .. world)
 * (device array of 64-bit addresses; d_peer_ptrs[rank] == local) and, optionally, at one multicast address (0: none).
 * d_signal_pads: device array of the ranks' signal pads (>= world * 4 bytes each, zero-initialised once);
 * d_epoch: 4 zero-initialised uint32 on this device.  n_floats must be a multiple of 4 and the arena 16-byte aligned.
 * `blocks` CTAs must be co-resident (<= the SM count).  Every rank launches it in the same order relative to other launches
 * on these pads. */
int ffb_allreduce_symm(float* local, const uint64_t* d_peer_ptrs, uint64_t multicast_ptr, const uint64_t* d_signal_pads, uint32_t* d_epoch,
                       int32_t rank, int32_t world, int64_t n_floats, int32_t blocks, void* stream) {
  FFB_REQUIRE(local && d_peer_ptrs && d_signal_pads && d_epoch, "null argument");
  FFB_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && blocks >= 1 && blocks <= sm_count(), "bad rank / world / blocks");
  FFB_REQUIRE((n_floats & 3) == 0 && ((uintptr_t)local & 15) == 0 && (multicast_ptr & 15) == 0, "arena must be 16-byte aligned with a multiple of 4 floats");
  if (n_floats == 0 || world == 1) return FFB_OK;
  ArArgs a;
  a.local = local; a.peers = d_peer_ptrs; a.mc = multicast_ptr; a.pads = d_signal_pads; a.epoch = d_epoch; a.n4 = n_floats / 4;
  a.rank = rank; a.world = world;
  if (multicast_ptr) allreduce_symm_kernel<true><<<(unsigned)blocks, 512, 0, (cudaStream_t)stream>>>(a);
  else allreduce_symm_kernel<false><<<(unsigned)blocks, 512, 0, (cudaStream_t)stream>>>(a);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
