// sampler.cu — ray sampling + alpha-mask stream compaction.
// Replaces FactorFields.sample_point (FactorFields.py:586-602), sample_point_ndc (:575-584), sample_point_unbound
// (:604-633), AlphaGridMask.sample_alpha (:103-110) and
// the boolean-mask gathers of forward (:864-867, :874) — which in the reference materialise dense
// [rays, samples, 3] tensors, call nonzero()/index() and sync the host on `.any()`.
//
// Design: one warp per ray.  Pass 1 counts valid samples per ray with warp ballots, a device-wide
// exclusive scan turns counts into offsets, pass 2 recomputes the (cheap) positions and writes the
// compacted list in row-major (ray, sample) order — the order torch boolean-mask indexing produces, so
// indices compare bit-exactly.  All decision arithmetic is the unfused fp32 sequence of ffb_math.h.
#include "ffb_common.cuh"
#include "ffb_math.h"
#include <map>
#include <mutex>
#include <utility>

namespace ffb {

struct RaySetup {
  float o[3], d[3];
  float tmin;
};

// interpx of sample s: marched from the box entry (mode 0) or read from the per-call table shared by all rays (modes 1, 2)
__device__ __forceinline__ float sample_z(const ffb_sampler_desc& D, const RaySetup& r, int s, float jit, bool train) {
  return D.mode == FFB_SAMPLE_BOUNDED ? sample_t(r.tmin, D.step_size, s, jit, train) : __ldg(D.z_table + s);
}

// forward differences of interpx (FactorFields.py:850 unbounded: last = previous; :854-856 NDC: last 0, times |d|; :861 last 0)
__device__ __forceinline__ float sample_dist(const ffb_sampler_desc& D, const RaySetup& r, int s, float t, float jit, bool train,
                                             float dnorm) {
  if (D.mode == FFB_SAMPLE_UNBOUND) {
    if (s + 1 < D.n_samples) return FFB_SUB(__ldg(D.z_table + s + 1), t);
    return D.n_samples > 1 ? FFB_SUB(t, __ldg(D.z_table + s - 1)) : 0.0f;
  }
  if (s + 1 >= D.n_samples) return 0.0f;
  const float dz = FFB_SUB(sample_z(D, r, s + 1, jit, train), t);
  return D.mode == FFB_SAMPLE_NDC ? FFB_MUL(dz, dnorm) : dz;
}

__device__ __forceinline__ bool sample_valid(const ffb_sampler_desc& D, const RaySetup& r, int s, float jit, bool train, float p[3],
                                             float* t_out, bool* inner_out = nullptr) {
  const float t = sample_z(D, r, s, jit, train);
  if (t_out) *t_out = t;
  if (D.mode == FFB_SAMPLE_UNBOUND) {
    // every sample is kept; the alpha mask only prunes samples inside the unit cube (:864-867 with ray_valid = ones)
    const bool inner = sample_pos_unbound(r.o, r.d, t, D.bg_len, p);
    if (inner_out) *inner_out = inner;
    if (inner && D.alpha_volume) return alpha_lookup(D.alpha_volume, D.alpha_size, D.alpha_aabb_min, D.alpha_inv_size, p) > D.alpha_thres;
    return true;
  }
  bool ok = sample_pos(r.o, r.d, t, D.aabb_min, D.aabb_max, p);
  if (inner_out) *inner_out = ok;
  // forward() looks the mask up for in-box samples only (:864-867); filtering_rays looks it up for EVERY sample (:832-833: a
  // sample just outside the box still interpolates the boundary voxels under zeros padding)
  if (D.alpha_volume && (ok || D.alpha_outside)) ok = alpha_lookup(D.alpha_volume, D.alpha_size, D.alpha_aabb_min, D.alpha_inv_size, p) > D.alpha_thres;
  return ok;
}

__device__ __forceinline__ float dir_norm(const RaySetup& r) {
  return sqrtf(FFB_ADD(FFB_ADD(FFB_MUL(r.d[0], r.d[0]), FFB_MUL(r.d[1], r.d[1])), FFB_MUL(r.d[2], r.d[2])));
}

__device__ __forceinline__ void load_ray(const float* __restrict__ rays, int64_t r, RaySetup& rs) {
  for (int k = 0; k < 3; ++k) {
    rs.o[k] = rays[r * 6 + k];
    rs.d[k] = rays[r * 6 + 3 + k];
  }
}

__global__ void __launch_bounds__(256) sample_count_kernel(ffb_sampler_desc D, const float* __restrict__ rays,
                                                           const float* __restrict__ jitter, int64_t R, int32_t* __restrict__ counts,
                                                           float* __restrict__ tmin) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    RaySetup rs;
    load_ray(rays, r, rs);
    rs.tmin = D.mode == FFB_SAMPLE_BOUNDED ? ray_tmin(rs.o, rs.d, D.aabb_min, D.aabb_max) : 0.0f;
    const bool train = jitter != nullptr;
    const float jit = train ? jitter[r] : 0.0f;
    int cnt = 0;
    for (int s0 = 0; s0 < D.n_samples; s0 += 32) {
      const int s = s0 + lane;
      float p[3];
      const bool ok = s < D.n_samples && sample_valid(D, rs, s, jit, train, p, nullptr);
      cnt += __popc(__ballot_sync(0xffffffffu, ok));
    }
    if (lane == 0) {
      counts[r] = cnt;
      if (tmin) tmin[r] = rs.tmin;
    }
  }
}

__global__ void __launch_bounds__(256) sample_fill_kernel(ffb_sampler_desc D, const float* __restrict__ rays,
                                                          const float* __restrict__ jitter, const float* __restrict__ tmin,
                                                          const int32_t* __restrict__ offsets, int64_t R, int64_t cap,
                                                          float* __restrict__ xyz, int32_t* __restrict__ ray_id,
                                                          int32_t* __restrict__ sample_id, float* __restrict__ z,
                                                          float* __restrict__ dist) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    RaySetup rs;
    load_ray(rays, r, rs);
    rs.tmin = D.mode != FFB_SAMPLE_BOUNDED ? 0.0f : (tmin ? tmin[r] : ray_tmin(rs.o, rs.d, D.aabb_min, D.aabb_max));
    const bool train = jitter != nullptr;
    const float jit = train ? jitter[r] : 0.0f;
    const float dnorm = D.mode == FFB_SAMPLE_NDC ? dir_norm(rs) : 1.0f;
    int64_t base = offsets[r];
    for (int s0 = 0; s0 < D.n_samples; s0 += 32) {
      const int s = s0 + lane;
      float p[3], t = 0.0f;
      const bool ok = s < D.n_samples && sample_valid(D, rs, s, jit, train, p, &t);
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (ok) {
        const int64_t i = base + __popc(m & ((1u << lane) - 1u));
        if (i < cap) {
          xyz[i * 3 + 0] = p[0];
          xyz[i * 3 + 1] = p[1];
          xyz[i * 3 + 2] = p[2];
          if (ray_id) ray_id[i] = (int32_t)r;
          if (sample_id) sample_id[i] = s;
          if (z) z[i] = t;
          if (dist) dist[i] = sample_dist(D, rs, s, t, jit, train, dnorm);
        }
      }
      base += __popc(m);
    }
  }
}

__global__ void __launch_bounds__(256) sample_dense_kernel(ffb_sampler_desc D, const float* __restrict__ rays,
                                                           const float* __restrict__ jitter, int64_t R, uint8_t* __restrict__ mask,
                                                           float* __restrict__ z, float* __restrict__ pts) {
  const int64_t total = R * D.n_samples;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = t / D.n_samples;
    const int s = (int)(t % D.n_samples);
    RaySetup rs;
    load_ray(rays, r, rs);
    rs.tmin = D.mode == FFB_SAMPLE_BOUNDED ? ray_tmin(rs.o, rs.d, D.aabb_min, D.aabb_max) : 0.0f;
    const bool train = jitter != nullptr;
    float p[3], tt;
    bool inner;
    const bool ok = sample_valid(D, rs, s, train ? jitter[r] : 0.0f, train, p, &tt, &inner);
    // the reference's sample_point* return the in-box / inner mask; with an alpha volume attached this is ray_valid
    mask[t] = (D.alpha_volume ? ok : inner) ? 1 : 0;
    if (z) z[t] = tt;
    if (pts) {
      pts[t * 3 + 0] = p[0];
      pts[t * 3 + 1] = p[1];
      pts[t * 3 + 2] = p[2];
    }
  }
}

__global__ void alpha_sample_kernel(ffb_sampler_desc D, const float* __restrict__ xyz, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float p[3] = {xyz[i * 3], xyz[i * 3 + 1], xyz[i * 3 + 2]};
    out[i] = alpha_lookup(D.alpha_volume, D.alpha_size, D.alpha_aabb_min, D.alpha_inv_size, p);
  }
}

// ---- exclusive scan (single pass, decoupled look-back over 1024-element tiles) -----------------
constexpr int SCAN_T = 256, SCAN_PER = 4, SCAN_TILE = SCAN_T * SCAN_PER;

__global__ void __launch_bounds__(SCAN_T) scan_tiles_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
                                                            unsigned long long* __restrict__ tile_state /* zeroed */,
                                                            unsigned int* __restrict__ ticket) {
  __shared__ int warp_sums[SCAN_T / 32];
  __shared__ long long s_prefix;
  __shared__ unsigned s_tile;
  if (threadIdx.x == 0) s_tile = atomicAdd(ticket, 1u);   // dynamic tile id => forward-progress safe look-back
  __syncthreads();
  const unsigned tile = s_tile;
  const int64_t base = (int64_t)tile * SCAN_TILE + threadIdx.x * SCAN_PER;
  int v[SCAN_PER];
  int sum = 0;
#pragma unroll
  for (int j = 0; j < SCAN_PER; ++j) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    sum += v[j];
  }
  // block exclusive scan of `sum`
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = sum;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += y;
  }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = lane < SCAN_T / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_T / 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += y;
    }
    if (lane < SCAN_T / 32) warp_sums[lane] = ws;
  }
  __syncthreads();
  const int warp_off = wid ? warp_sums[wid - 1] : 0;
  const int excl = warp_off + inc - sum;
  const int tile_total = warp_sums[SCAN_T / 32 - 1];
  // state word: bits 62..63 flag (1 = aggregate, 2 = inclusive prefix), low bits value
  if (threadIdx.x == 0) {
    long long prefix = 0;
    if (tile == 0) {
      atomicExch(&tile_state[0], (2ull << 62) | (unsigned long long)(unsigned)tile_total);
    } else {
      atomicExch(&tile_state[tile], (1ull << 62) | (unsigned long long)(unsigned)tile_total);
      long long run = 0;
      int look = (int)tile - 1;
      while (true) {
        unsigned long long st = atomicAdd(&tile_state[look], 0ull);
        unsigned flag = (unsigned)(st >> 62);
        if (flag == 0) continue;
        run += (long long)(st & 0x3fffffffffffffffull);
        if (flag == 2) break;
        --look;
      }
      prefix = run;
      atomicExch(&tile_state[tile], (2ull << 62) | (unsigned long long)(prefix + tile_total));
    }
    s_prefix = prefix;
  }
  __syncthreads();
  int run = (int)s_prefix + excl;
#pragma unroll
  for (int j = 0; j < SCAN_PER; ++j) {
    if (base + j < n) out[base + j] = run;
    run += v[j];
  }
  // total at out[n]
  if (base <= n - 1 && n - 1 < base + SCAN_PER) out[n] = run;
}

// Single-CTA exclusive scan for small inputs (n <= SCAN_SINGLE_MAX): 1024 threads, 4 elements per thread per round,
// running carry between rounds.  No workspace, so it is safe inside CUDA-graph capture and on any stream.
constexpr int64_t SCAN_SINGLE_MAX = 1 << 16;

__global__ void __launch_bounds__(1024) scan_single_block_kernel(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n) {
  __shared__ int warp_sums[32];
  __shared__ int s_carry;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int64_t base0 = 0; base0 < n; base0 += 4096) {
    const int64_t base = base0 + threadIdx.x * 4;
    int v[4];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = (base + j < n) ? in[base + j] : 0;
      sum += v[j];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int y = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      int ws = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int y = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += y;
      }
      warp_sums[lane] = ws;
    }
    __syncthreads();
    const int carry = s_carry;
    int run = carry + (wid ? warp_sums[wid - 1] : 0) + inc - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (base + j < n) out[base + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = carry + warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[n] = s_carry;
}

// Look-back state for the multi-tile scan: one cached device buffer per stream (grown on demand, never freed), so the
// hot path never calls the allocator.
static unsigned long long* scan_workspace(cudaStream_t s, size_t words) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, std::pair<unsigned long long*, size_t>> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto& e = cache[{dev, s}];
  if (e.second < words) {
    if (e.first) {
      cudaStreamSynchronize(s);
      cudaFree(e.first);
      e = {nullptr, 0};
    }
    size_t cap = words < 4096 ? 4096 : words * 2;
    if (cudaMalloc(&e.first, cap * sizeof(unsigned long long)) != cudaSuccess) {
      e = {nullptr, 0};
      return nullptr;
    }
    e.second = cap;
  }
  return e.first;
}

}  // namespace ffb

using namespace ffb;

static int check_sampler(const ffb_sampler_desc* d) {
  FFB_REQUIRE(d, "null descriptor");
  FFB_REQUIRE(d->n_samples > 0, "n_samples must be positive");
  if (d->alpha_volume) FFB_REQUIRE(d->alpha_size[0] > 0 && d->alpha_size[1] > 0 && d->alpha_size[2] > 0, "bad alpha volume size");
  FFB_REQUIRE(d->mode >= FFB_SAMPLE_BOUNDED && d->mode <= FFB_SAMPLE_UNBOUND, "unknown sampling mode");
  if (d->mode != FFB_SAMPLE_BOUNDED) FFB_REQUIRE(d->z_table, "z_table is required for the NDC / unbounded sampling modes");
  return FFB_OK;
}

extern "C" {

int ffb_sample_count(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter, int64_t R, int32_t* counts, float* tmin,
                     void* stream) {
  int rc = check_sampler(h_desc);
  if (rc) return rc;
  FFB_REQUIRE(rays && counts, "null argument");
  if (R <= 0) return FFB_OK;
  sample_count_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, rays, jitter, R, counts, tmin);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_exclusive_scan_i32(const int32_t* counts, int32_t* offsets, int64_t R, void* stream) {
  FFB_REQUIRE(counts && offsets, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (R <= 0) {
    FFB_CUDA(cudaMemsetAsync(offsets, 0, sizeof(int32_t), s));
    return FFB_OK;
  }
  if (R <= SCAN_SINGLE_MAX) {   // one CTA, no workspace, no memset: the per-ray counts of a training batch
    scan_single_block_kernel<<<1, 1024, 0, s>>>(counts, offsets, R);
    FFB_LAUNCHED();
    return FFB_OK;
  }
  const int64_t tiles = (R + SCAN_TILE - 1) / SCAN_TILE;
  unsigned long long* state = scan_workspace(s, (size_t)(tiles + 1));
  FFB_REQUIRE(state, "cannot allocate the scan workspace");
  FFB_CUDA(cudaMemsetAsync(state, 0, sizeof(unsigned long long) * (tiles + 1), s));
  scan_tiles_kernel<<<(unsigned)tiles, SCAN_T, 0, s>>>(counts, offsets, R, state, reinterpret_cast<unsigned int*>(state + tiles));
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_sample_fill(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter, const float* tmin, const int32_t* offsets,
                    int64_t R, int64_t cap, float* xyz, int32_t* ray_id, int32_t* sample_id, float* z, float* dist, void* stream) {
  int rc = check_sampler(h_desc);
  if (rc) return rc;
  FFB_REQUIRE(rays && offsets && xyz, "null argument");
  if (R <= 0) return FFB_OK;
  sample_fill_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, rays, jitter, tmin, offsets, R, cap, xyz,
                                                                                          ray_id, sample_id, z, dist);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_sample_dense(const ffb_sampler_desc* h_desc, const float* rays, const float* jitter, int64_t R, uint8_t* mask, float* z,
                     float* pts, void* stream) {
  int rc = check_sampler(h_desc);
  if (rc) return rc;
  FFB_REQUIRE(rays && mask, "null argument");
  if (R <= 0) return FFB_OK;
  sample_dense_kernel<<<blocks_for(R * h_desc->n_samples, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, rays, jitter, R, mask, z, pts);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_alpha_sample(const ffb_sampler_desc* h_desc, const float* xyz, int64_t n, float* out, void* stream) {
  FFB_REQUIRE(h_desc && h_desc->alpha_volume && xyz && out, "null argument / no alpha volume");
  if (n <= 0) return FFB_OK;
  alpha_sample_kernel<<<blocks_for(n, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, xyz, n, out);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_sample_dense_host(const ffb_sampler_desc* h_desc, const float* h_rays, const float* h_jitter, int64_t R, uint8_t* h_mask,
                          float* h_z) {
  int rc = check_sampler(h_desc);
  if (rc) return rc;
  FFB_REQUIRE(h_rays && h_mask, "null argument");
  if (R <= 0) return FFB_OK;
  const int64_t S = h_desc->n_samples;
  float *d_rays = nullptr, *d_jit = nullptr, *d_z = nullptr;
  uint8_t* d_mask = nullptr;
  cudaStream_t s = nullptr;
  FFB_CUDA(cudaMalloc(&d_rays, sizeof(float) * R * 6));
  FFB_CUDA(cudaMalloc(&d_mask, R * S));
  if (h_jitter) FFB_CUDA(cudaMalloc(&d_jit, sizeof(float) * R));
  if (h_z) FFB_CUDA(cudaMalloc(&d_z, sizeof(float) * R * S));
  FFB_CUDA(cudaMemcpyAsync(d_rays, h_rays, sizeof(float) * R * 6, cudaMemcpyHostToDevice, s));
  if (h_jitter) FFB_CUDA(cudaMemcpyAsync(d_jit, h_jitter, sizeof(float) * R, cudaMemcpyHostToDevice, s));
  rc = ffb_sample_dense(h_desc, d_rays, d_jit, R, d_mask, d_z, nullptr, s);
  if (rc == FFB_OK) rc = check_cuda(cudaMemcpyAsync(h_mask, d_mask, R * S, cudaMemcpyDeviceToHost, s), "D2H mask");
  if (rc == FFB_OK && h_z) rc = check_cuda(cudaMemcpyAsync(h_z, d_z, sizeof(float) * R * S, cudaMemcpyDeviceToHost, s), "D2H z");
  if (rc == FFB_OK) rc = check_cuda(cudaStreamSynchronize(s), "sync");
  cudaFree(d_rays); cudaFree(d_mask); cudaFree(d_jit); cudaFree(d_z);
  return rc;
}

}  // extern "C"
