// optim.cu — train-step glue: MSE loss (train_per_scene.py:158) and Adam (train_per_scene.py:124-132,160-162).
#include "ffb_common.cuh"

namespace ffb {

__global__ void __launch_bounds__(256) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int64_t n,
                                                  float g_scale, float* __restrict__ loss, float* __restrict__ g_pred) {
  float s = 0.0f;
  const float inv_n = 1.0f / (float)n;
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

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_mse_fwd_bwd(const float* pred, const float* target, int64_t n, float g_scale, float* loss, float* g_pred, void* stream) {
  FFB_REQUIRE(pred && target && loss && n > 0, "bad argument");
  mse_kernel<<<blocks_for(n, 256, sm_count() * 4), 256, 0, (cudaStream_t)stream>>>(pred, target, n, g_scale, loss, g_pred);
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

}  // extern "C"
