// optim.cu — train-step glue: MSE loss (train_per_scene.py:158) and Adam (train_per_scene.py:124-132,160-162).
#include "ffb_common.cuh"

namespace ffb {

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                  float g_scale, const float* __restrict__ g_scale_dev, float* __restrict__ loss,
                                                  float* __restrict__ g_pred) {
  float s = 0.0f;
  const float inv_n = 1.0f / (float)n;
  if (g_scale_dev) g_scale *= __ldg(g_scale_dev);     // e.g. the decaying loss scale of the regression drivers
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = pred[i] - target[i];
    s += d * d;
    if (g_pred) g_pred[i] = 2.0f * d * inv_n * g_scale;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.0f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    atomicAdd(loss, t * inv_n);
  }
}

// torch.optim.Adam._single_tensor_adam, fp32:
//   m = b1 m + (1-b1) g ; v = b2 v + (1-b2) g^2 ; denom = sqrt(v)/sqrt(1-b2^t) + eps ; p -= lr/(1-b1^t) * m/denom
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, int64_t n, float step_size, float beta1, float beta2,
                                                   float eps, float inv_sqrt_bc2, float grad_scale) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    const float mi = m[i] * beta1 + gi * (1.0f - beta1);
    const float vi = v[i] * beta2 + gi * gi * (1.0f - beta2);
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) * inv_sqrt_bc2 + eps;
    p[i] -= step_size * (mi / denom);
  }
}

// Multi-tensor Adam: one launch for every parameter tensor.  Block b updates chunk b = CHUNK consecutive elements of
// tensor chunk_tensor[b] starting at chunk_start[b]; the step-dependent scalars are read from DEVICE memory
// (hyper[group] = {lr / (1 - b1^t), 1 / sqrt(1 - b2^t)}) so the launch can sit inside a replayed CUDA graph.
__global__ void __launch_bounds__(256) adam_multi_kernel(const int64_t* __restrict__ table /*[T][6]: p g m v n group*/,
                                                         const int32_t* __restrict__ chunk_tensor,
                                                         const int64_t* __restrict__ chunk_start, int chunk,
                                                         const float* __restrict__ hyper, float beta1, float beta2, float eps,
                                                         float grad_scale) {
  const int t = chunk_tensor[blockIdx.x];
  const int64_t* e = table + (int64_t)t * 6;
  float* __restrict__ p = reinterpret_cast<float*>(e[0]);
  const float* __restrict__ g = reinterpret_cast<const float*>(e[1]);
  float* __restrict__ m = reinterpret_cast<float*>(e[2]);
  float* __restrict__ v = reinterpret_cast<float*>(e[3]);
  const int64_t n = e[4];
  const int grp = (int)e[5];
  const float step_size = hyper[2 * grp], inv_sqrt_bc2 = hyper[2 * grp + 1];
  const int64_t begin = chunk_start[blockIdx.x];
  const int64_t end = begin + chunk < n ? begin + chunk : n;
  const float ob1 = 1.0f - beta1, ob2 = 1.0f - beta2;
  auto upd = [&](float pi, float gi, float& mi, float& vi) {
    gi *= grad_scale;
    mi = mi * beta1 + gi * ob1;
    vi = vi * beta2 + gi * gi * ob2;
    return pi - step_size * (mi / (sqrtf(vi) * inv_sqrt_bc2 + eps));
  };
  int64_t i0 = begin;
  const uintptr_t bits = reinterpret_cast<uintptr_t>(p + begin) | reinterpret_cast<uintptr_t>(g + begin) |
                         reinterpret_cast<uintptr_t>(m + begin) | reinterpret_cast<uintptr_t>(v + begin);
  if ((bits & 15) == 0) {      // 16-byte vector path (arena slices and torch allocations are 256-byte aligned)
    const int64_t n4 = (end - begin) >> 2;
    float4* p4 = reinterpret_cast<float4*>(p + begin);
    const float4* g4 = reinterpret_cast<const float4*>(g + begin);
    float4* m4 = reinterpret_cast<float4*>(m + begin);
    float4* v4 = reinterpret_cast<float4*>(v + begin);
    // the moments and the gradient are touched once per step: streaming (evict-first) accesses, so that this 7-stream sweep
    // leaves the PARAMETERS — what the next step's gathers read — in L2 rather than a mix of all four arrays
    for (int64_t i = threadIdx.x; i < n4; i += blockDim.x) {
      float4 pp = p4[i], mm = __ldcs(m4 + i), vv = __ldcs(v4 + i);
      const float4 gg = __ldcs(g4 + i);
      pp.x = upd(pp.x, gg.x, mm.x, vv.x);
      pp.y = upd(pp.y, gg.y, mm.y, vv.y);
      pp.z = upd(pp.z, gg.z, mm.z, vv.z);
      pp.w = upd(pp.w, gg.w, mm.w, vv.w);
      p4[i] = pp;
      __stcs(m4 + i, mm);
      __stcs(v4 + i, vv);
    }
    i0 = begin + n4 * 4;
  }
  for (int64_t i = i0 + threadIdx.x; i < end; i += blockDim.x) {
    float mi = m[i], vi = v[i];
    p[i] = upd(p[i], g[i], mi, vi);
    m[i] = mi;
    v[i] = vi;
  }
}

// One thread: t += 1; hyper[g] = {lr_g / (1 - b1^t), 1 / sqrt(1 - b2^t)}; lr_g *= decay  (torch.optim.Adam's scalar
// bookkeeping + train_per_scene.py:170-171, in double like the Python reference) -- device-side so that a replayed
// CUDA graph advances the optimiser without host traffic.
__global__ void adam_hyper_kernel(double* __restrict__ lr, long long* __restrict__ step, float* __restrict__ hyper, int n_groups,
                                  double beta1, double beta2, double decay) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const long long t = *step + 1;
  *step = t;
  const double bc1 = 1.0 - pow(beta1, (double)t), bc2 = 1.0 - pow(beta2, (double)t);
  for (int g = 0; g < n_groups; ++g) {
    hyper[2 * g] = (float)(lr[g] / bc1);
    hyper[2 * g + 1] = (float)(1.0 / sqrt(bc2));
    lr[g] *= decay;
  }
}

// value *= factor (double, like the Python scalar `loss_scale *= lr_factor` of scripts/2D_regression.ipynb cell 4 /
// sdf_regression.ipynb cell 2); the fp32 copy is what a captured graph's loss kernel reads.
__global__ void scalar_decay_kernel(double* __restrict__ value, double factor, float* __restrict__ out_f32) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  const double v = *value * factor;
  *value = v;
  if (out_f32) *out_f32 = (float)v;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_mse_fwd_bwd(const float* pred, const float* target, int64_t n, float g_scale, const float* g_scale_dev, float* loss,
                    float* g_pred, void* stream) {
  FFB_REQUIRE(pred && target && loss && n > 0, "bad argument");
  mse_kernel<<<blocks_for(n, 256, sm_count() * 4), 256, 0, (cudaStream_t)stream>>>(pred, target, n, g_scale, g_scale_dev, loss, g_pred);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_scalar_decay(double* d_value, double factor, float* d_out_f32, void* stream) {
  FFB_REQUIRE(d_value, "null argument");
  scalar_decay_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_value, factor, d_out_f32);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                  int32_t step, float grad_scale, void* stream) {
  FFB_REQUIRE(p && g && m && v && step >= 1, "bad argument");
  if (n <= 0) return FFB_OK;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  adam_kernel<<<blocks_for(n, 256, sm_count() * 8), 256, 0, (cudaStream_t)stream>>>(p, g, m, v, n, (float)(lr / bc1), beta1, beta2, eps,
                                                                                   (float)(1.0 / sqrt(bc2)), grad_scale);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_adam_hyper_advance(double* d_lr, int64_t* d_step, float* d_hyper, int32_t n_groups, double beta1, double beta2,
                           double lr_decay, void* stream) {
  FFB_REQUIRE(d_lr && d_step && d_hyper && n_groups > 0, "bad argument");
  adam_hyper_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(d_lr, reinterpret_cast<long long*>(d_step), d_hyper, n_groups, beta1, beta2, lr_decay);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_adam_multi(const int64_t* d_table, const int32_t* d_chunk_tensor, const int64_t* d_chunk_start, int32_t n_chunks,
                   int32_t chunk, const float* d_hyper, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  FFB_REQUIRE(d_table && d_chunk_tensor && d_chunk_start && d_hyper && chunk > 0, "bad argument");
  if (n_chunks <= 0) return FFB_OK;
  adam_multi_kernel<<<(unsigned)n_chunks, 256, 0, (cudaStream_t)stream>>>(d_table, d_chunk_tensor, d_chunk_start, chunk, d_hyper, beta1,
                                                                          beta2, eps, grad_scale);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
