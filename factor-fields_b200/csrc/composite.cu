// composite.cu — per-ray emission-absorption composite over the COMPACTED sample list, forward and backward.
// Replaces basis2density (FactorFields.py:639-643), raw2alpha (:82-88), the weight-threshold mask (:879-881)
// and the accumulation (:887-896) — dense [rays, samples] tensors + cumprod + index_put in the reference.
// Invalid samples have sigma = 0 -> alpha = 0 -> a transmittance factor of exactly 1.0f in fp32
// ((1 - 0) + 1e-10 == 1), so running the recurrence over the valid samples only is exact.
// One warp per ray; transmittance by a multiplicative warp scan with a carried prefix.
#include "ffb_common.cuh"
#include "ffb_math.h"

namespace ffb {

__device__ __forceinline__ float density_act(const ffb_composite_desc& D, float f) {
  const float x = FFB_ADD(f, D.density_shift);
  return D.softplus ? softplus_f(x) : fmaxf(x, 0.0f);
}
__device__ __forceinline__ float density_act_grad(const ffb_composite_desc& D, float f) {
  const float x = FFB_ADD(f, D.density_shift);
  if (D.softplus) return x > 20.0f ? 1.0f : sigmoid_f(x);
  return x > 0.0f ? 1.0f : 0.0f;
}

constexpr int CU = 4;      // chunks of 32 samples whose loads a warp issues together

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(256) composite_weights_kernel(ffb_composite_desc D, const float* __restrict__ feat0, int ld_feat,
                                                                const float* __restrict__ dist, const int32_t* __restrict__ offsets,
                                                                int64_t R, float* __restrict__ sigma, float* __restrict__ trans,
                                                                float* __restrict__ weight, int32_t* __restrict__ app_counts) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t beg = offsets[r], end = offsets[r + 1];
    float carry = 1.0f;
    int napp = 0;
    for (int64_t s0 = beg; s0 < end; s0 += 32 * CU) {
      // the loads of CU chunks are issued together (a ray's chunks are otherwise a chain of load -> scan -> load ...: ~1 us each)
      float fv[CU], dv[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t i = s0 + u * 32 + lane;
        fv[u] = 0.0f; dv[u] = 0.0f;
        if (i < end) { fv[u] = feat0[i * ld_feat]; dv[u] = dist[i]; }
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t c0 = s0 + u * 32;
        if (c0 >= end) break;
        const int64_t i = c0 + lane;
        const bool act = i < end;
        float sg = 0.0f, alpha = 0.0f, fac = 1.0f;
        if (act) {
          sg = density_act(D, fv[u]);
          const float delta = FFB_MUL(dv[u], D.distance_scale);
          alpha = FFB_SUB(1.0f, expf(-FFB_MUL(sg, delta)));
          fac = FFB_ADD(FFB_SUB(1.0f, alpha), 1e-10f);
        }
        float p = fac;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float y = __shfl_up_sync(0xffffffffu, p, o);
          if (lane >= o) p *= y;
        }
        float excl = __shfl_up_sync(0xffffffffu, p, 1);
        if (lane == 0) excl = 1.0f;
        const float T = carry * excl;
        const float w = alpha * T;
        carry *= __shfl_sync(0xffffffffu, p, 31);
        if (act) {
          sigma[i] = sg;
          trans[i] = T;
          weight[i] = w;
        }
        napp += __popc(__ballot_sync(0xffffffffu, act && w > D.weight_thres));
      }
    }
    if (lane == 0) app_counts[r] = napp;
  }
}

__global__ void __launch_bounds__(256) composite_app_fill_kernel(const float* __restrict__ weight, float thres,
                                                                 const int32_t* __restrict__ offsets,
                                                                 const int32_t* __restrict__ app_offsets, int64_t R,
                                                                 int32_t* __restrict__ app_idx, int32_t* __restrict__ app_slot) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t beg = offsets[r], end = offsets[r + 1];
    int64_t base = app_offsets[r];
    for (int64_t s0 = beg; s0 < end; s0 += 32 * CU) {
      float wv[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t i = s0 + u * 32 + lane;
        wv[u] = i < end ? weight[i] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t i = s0 + u * 32 + lane;
        if (s0 + u * 32 >= end) break;
        const bool f = i < end && wv[u] > thres;
        const unsigned m = __ballot_sync(0xffffffffu, f);
        const int64_t j = base + __popc(m & ((1u << lane) - 1u));
        if (f) app_idx[j] = (int32_t)i;
        if (app_slot && i < end) app_slot[i] = f ? (int32_t)j : -1;      // the inverse map: shaded-sample slot of sample i, or -1
        base += __popc(m);
      }
    }
  }
}

__global__ void __launch_bounds__(256) composite_accum_kernel(ffb_composite_desc D, const float* __restrict__ weight,
                                                              const float* __restrict__ z, const float* __restrict__ rgb,
                                                              const int32_t* __restrict__ offsets, const int32_t* __restrict__ app_offsets,
                                                              int64_t R, float* __restrict__ rgb_map, float* __restrict__ pre_clamp,
                                                              float* __restrict__ acc_out, float* __restrict__ depth) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t beg = offsets[r], end = offsets[r + 1];
    int64_t abase = app_offsets[r];
    float acc = 0.0f, dep = 0.0f, c0s = 0.0f, c1s = 0.0f, c2s = 0.0f;
    for (int64_t s0 = beg; s0 < end; s0 += 32 * CU) {
      float wv[CU], zv[CU], cr[CU][3];
      bool fl[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t i = s0 + u * 32 + lane;
        const bool act = i < end;
        wv[u] = act ? weight[i] : 0.0f;
        zv[u] = (act && z) ? z[i] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {           // shaded-sample slots of the CU chunks, then their colours in one batch of loads
        fl[u] = (s0 + u * 32 + lane < end) && wv[u] > D.weight_thres;
        const unsigned m = __ballot_sync(0xffffffffu, fl[u]);
        cr[u][0] = cr[u][1] = cr[u][2] = 0.0f;
        if (fl[u]) {
          const int64_t j = abase + __popc(m & ((1u << lane) - 1u));
          cr[u][0] = rgb[j * 3 + 0]; cr[u][1] = rgb[j * 3 + 1]; cr[u][2] = rgb[j * 3 + 2];
        }
        abase += __popc(m);
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        if (s0 + u * 32 + lane < end) {
          acc += wv[u];
          if (z) dep += wv[u] * zv[u];
        }
        if (fl[u]) {
          c0s += wv[u] * cr[u][0];
          c1s += wv[u] * cr[u][1];
          c2s += wv[u] * cr[u][2];
        }
      }
    }
    acc = warp_sum(acc);
    dep = warp_sum(dep);
    c0s = warp_sum(c0s);
    c1s = warp_sum(c1s);
    c2s = warp_sum(c2s);
    if (lane == 0) {
      float c[3] = {c0s, c1s, c2s};
      for (int k = 0; k < 3; ++k) {
        float v = c[k];
        if (D.white_bg_dev ? (*D.white_bg_dev != 0) : (D.white_bg != 0)) v = v + (1.0f - acc);
        if (pre_clamp) pre_clamp[r * 3 + k] = v;
        rgb_map[r * 3 + k] = fminf(fmaxf(v, 0.0f), 1.0f);
      }
      if (acc_out) acc_out[r] = acc;
      if (depth) depth[r] = dep;
    }
  }
}

__global__ void __launch_bounds__(256) composite_bwd_kernel(ffb_composite_desc D, const float* __restrict__ g_rgb_map,
                                                            const float* __restrict__ pre_clamp, const float* __restrict__ feat0, int ld_feat,
                                                            const float* __restrict__ dist, const float* __restrict__ sigma,
                                                            const float* __restrict__ trans, const float* __restrict__ weight,
                                                            const float* __restrict__ rgb, const int32_t* __restrict__ offsets,
                                                            const int32_t* __restrict__ app_offsets, int64_t R, float* __restrict__ g_rgb,
                                                            float* __restrict__ g_feat0, int ld_g, int zero_rest) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = warp; r < R; r += nwarps) {
    const int64_t beg = offsets[r], end = offsets[r + 1];
    if (beg >= end) continue;
    float g[3];
    for (int k = 0; k < 3; ++k) {
      const float pc = pre_clamp[r * 3 + k];
      g[k] = (pc >= 0.0f && pc <= 1.0f) ? g_rgb_map[r * 3 + k] : 0.0f;   // clamp(0,1) backward
    }
    const float gsum = (D.white_bg_dev ? (*D.white_bg_dev != 0) : (D.white_bg != 0)) ? (g[0] + g[1] + g[2]) : 0.0f;
    int64_t aend = app_offsets[r + 1];   // one past the last shaded sample of this ray
    float suffix = 0.0f;                 // sum_{k > chunk} g_w_k * w_k
    const int64_t nchunks = (end - beg + 31) / 32;
    for (int64_t ch_hi = nchunks - 1; ch_hi >= 0; ch_hi -= CU) {
      // the loads of CU chunks (walking down from ch_hi) are issued together, then the shaded samples' colours, then the scans
      float wv[CU], sv[CU], dv[CU], tv[CU], fv[CU], cr[CU][3];
      int64_t jv[CU];
      bool fl[CU];
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t ch = ch_hi - u, i = beg + ch * 32 + lane;
        const bool act = ch >= 0 && i < end;
        wv[u] = sv[u] = dv[u] = tv[u] = fv[u] = 0.0f;
        if (act) { wv[u] = weight[i]; sv[u] = sigma[i]; dv[u] = dist[i]; tv[u] = trans[i]; fv[u] = feat0[i * ld_feat]; }
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t ch = ch_hi - u, i = beg + ch * 32 + lane;
        fl[u] = ch >= 0 && i < end && wv[u] > D.weight_thres;
        const unsigned m = __ballot_sync(0xffffffffu, fl[u]);
        jv[u] = 0;
        cr[u][0] = cr[u][1] = cr[u][2] = 0.0f;
        if (fl[u]) {
          const unsigned upper = (lane == 31) ? 0u : (m & ~((2u << lane) - 1u));
          jv[u] = aend - 1 - __popc(upper);
          cr[u][0] = rgb[jv[u] * 3 + 0]; cr[u][1] = rgb[jv[u] * 3 + 1]; cr[u][2] = rgb[jv[u] * 3 + 2];
        }
        aend -= __popc(m);
      }
#pragma unroll
      for (int u = 0; u < CU; ++u) {
        const int64_t ch = ch_hi - u;
        if (ch < 0) break;
        const int64_t i = beg + ch * 32 + lane;
        const bool act = i < end;
        const float w = wv[u];
        float gw = -gsum;
        if (fl[u]) {
          const int64_t j = jv[u];
          gw += g[0] * cr[u][0] + g[1] * cr[u][1] + g[2] * cr[u][2];
          g_rgb[j * 3 + 0] = w * g[0];
          g_rgb[j * 3 + 1] = w * g[1];
          g_rgb[j * 3 + 2] = w * g[2];
        }
        const float gww = act ? gw * w : 0.0f;
        // reverse inclusive scan over lanes
        float s = gww;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const float y = __shfl_down_sync(0xffffffffu, s, o);
          if (lane + o < 32) s += y;
        }
        const float after = suffix + (s - gww);   // sum over samples after i
        suffix += __shfl_sync(0xffffffffu, s, 0);
        float g0 = 0.0f;
        if (act) {
          const float sg = sv[u];
          const float delta = FFB_MUL(dv[u], D.distance_scale);
          const float e = expf(-FFB_MUL(sg, delta));          // 1 - alpha
          const float fac = FFB_ADD(FFB_SUB(1.0f, FFB_SUB(1.0f, e)), 1e-10f);
          const float g_alpha = gw * tv[u] - after / fac;
          const float g_sigma = g_alpha * e * delta;
          g0 = g_sigma * density_act_grad(D, fv[u]);
          if (!zero_rest) g_feat0[i * ld_g] = g0;
        }
        if (zero_rest) {
          // the whole gradient rows of this chunk: density column + zeros (the caller skips its memset of [Nv, ld_g]).  The chunk's
          // rows are contiguous in memory, so the warp writes them as coalesced 16-byte pieces; piece t = row t / q4, quad t % q4.
          const int q4 = ld_g >> 2;
          const int64_t cbase = beg + ch * 32;
          const int rows = (int)((end - cbase) < 32 ? (end - cbase) : 32);
          float4* dst = reinterpret_cast<float4*>(g_feat0 + cbase * ld_g);
          for (int t0 = 0; t0 < 32 * q4; t0 += 32) {
            const int t = t0 + lane, r = t / q4, q = t - r * q4;
            const float gr = __shfl_sync(0xffffffffu, g0, r & 31);
            if (r < rows) dst[t] = make_float4(q == 0 ? gr : 0.f, 0.f, 0.f, 0.f);
          }
        }
      }
    }
  }
}

__global__ void density_alpha_kernel(ffb_composite_desc D, const float* __restrict__ feat0, int ld_feat, float length, int64_t n,
                                     const int32_t* __restrict__ n_dev, float* __restrict__ alpha) {
  n = resolve_n(n, n_dev);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float sg = density_act(D, feat0[i * ld_feat]);
    alpha[i] = FFB_SUB(1.0f, expf(-FFB_MUL(sg, length)));
  }
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_composite_weights(const ffb_composite_desc* h_desc, const float* feat0, int32_t ld_feat, const float* dist,
                          const int32_t* offsets, int64_t R, float* sigma, float* trans, float* weight, int32_t* app_counts,
                          void* stream) {
  FFB_REQUIRE(h_desc && feat0 && dist && offsets && sigma && trans && weight && app_counts && ld_feat >= 1, "bad argument");
  if (R <= 0) return FFB_OK;
  composite_weights_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, feat0, ld_feat, dist, offsets, R,
                                                                                                sigma, trans, weight, app_counts);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_composite_app_fill_ex(const float* weight, float weight_thres, const int32_t* offsets, const int32_t* app_offsets, int64_t R,
                              int32_t* app_idx, int32_t* app_slot, void* stream) {
  FFB_REQUIRE(weight && offsets && app_offsets && app_idx, "null argument");
  if (R <= 0) return FFB_OK;
  composite_app_fill_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(weight, weight_thres, offsets, app_offsets,
                                                                                                 R, app_idx, app_slot);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_composite_app_fill(const float* weight, float weight_thres, const int32_t* offsets, const int32_t* app_offsets, int64_t R,
                           int32_t* app_idx, void* stream) {
  return ffb_composite_app_fill_ex(weight, weight_thres, offsets, app_offsets, R, app_idx, nullptr, stream);
}

int ffb_composite_accum(const ffb_composite_desc* h_desc, const float* weight, const float* z, const float* rgb,
                        const int32_t* offsets, const int32_t* app_offsets, int64_t R, float* rgb_map, float* pre_clamp, float* acc,
                        float* depth, void* stream) {
  FFB_REQUIRE(h_desc && weight && offsets && app_offsets && rgb_map, "null argument");   // rgb may be NULL when no sample is shaded
  if (R <= 0) return FFB_OK;
  composite_accum_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, weight, z, rgb, offsets, app_offsets, R,
                                                                                              rgb_map, pre_clamp, acc, depth);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_composite_bwd(const ffb_composite_desc* h_desc, const float* g_rgb_map, const float* pre_clamp, const float* feat0,
                      int32_t ld_feat, const float* dist, const float* sigma, const float* trans, const float* weight,
                      const float* rgb, const int32_t* offsets, const int32_t* app_offsets, int64_t R, float* g_rgb, float* g_feat0,
                      int32_t ld_g, int32_t zero_rest, void* stream) {
  FFB_REQUIRE(h_desc && g_rgb_map && pre_clamp && feat0 && dist && sigma && trans && weight && offsets && app_offsets && g_feat0,
              "null argument");   // rgb / g_rgb may be NULL when no sample is shaded
  FFB_REQUIRE(!zero_rest || ((ld_g & 3) == 0 && ((uintptr_t)g_feat0 & 15) == 0), "zero_rest needs 16-byte aligned gradient rows");
  if (R <= 0) return FFB_OK;
  composite_bwd_kernel<<<blocks_for(R * 32, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(
      *h_desc, g_rgb_map, pre_clamp, feat0, ld_feat, dist, sigma, trans, weight, rgb, offsets, app_offsets, R, g_rgb, g_feat0, ld_g, zero_rest);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_density_alpha(const ffb_composite_desc* h_desc, const float* feat0, int32_t ld_feat, float length, int64_t n,
                      const int32_t* n_dev, float* alpha, void* stream) {
  FFB_REQUIRE(h_desc && feat0 && alpha && ld_feat >= 1, "bad argument");
  if (n <= 0) return FFB_OK;
  density_alpha_kernel<<<blocks_for(n, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(*h_desc, feat0, ld_feat, length, n, n_dev, alpha);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
