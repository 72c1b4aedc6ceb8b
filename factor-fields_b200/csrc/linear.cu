// linear.cu — fp32 dense layers for the decoder MLPs (MLPMixer FactorFields.py:113-159, MLPRender_Fea
// :162-203): SIMT register-tiled GEMM kernels that cover every layer shape the presets use.  The
// tcgen05 (tensor-core, bf16x3 split) path for the hot shapes lives in mlp_tc.cu; these kernels are the
// exact-fp32 general path and the building blocks of the backward pass.
#include "ffb_common.cuh"
#include "ffb_math.h"

namespace ffb {

constexpr int BM = 64, BN = 64, BK = 16;

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return sigmoid_f(v);
  return v;
}
__device__ __forceinline__ float act_mask(float y, int act) {
  if (act == 1) return y > 0.0f ? 1.0f : 0.0f;
  if (act == 2) return y * (1.0f - y);
  return 1.0f;
}

// C[n, N] = epilogue( (A .* mask(Y)) [n, Kin] * B ),  B(kin, col) = Bp[kin * sbk + col * sbc]
// A row stride = Kin, C row stride = N.  mask only when Y != nullptr (act_in).
__global__ void __launch_bounds__(256) gemm_rows_kernel(const float* __restrict__ A, const float* __restrict__ Y, int act_in,
                                                        const float* __restrict__ Bp, int64_t sbk, int64_t sbc,
                                                        const float* __restrict__ bias, int act_out, float* __restrict__ C,
                                                        int64_t n, const int32_t* __restrict__ n_dev, int Kin, int N) {
  n = resolve_n(n, n_dev);
  const int64_t row0 = (int64_t)blockIdx.x * BM;
  if (row0 >= n) return;
  const int col0 = blockIdx.y * BN;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int t = threadIdx.x, ty = t / 16, tx = t % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int k0 = 0; k0 < Kin; k0 += BK) {
    {  // A tile: 64 rows x 16 k ; thread -> (row = t/4, k = (t%4)*4 + j)
      const int r = t / 4, kq = (t % 4) * 4;
      const int64_t row = row0 + r;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + kq + j;
        float v = 0.0f;
        if (row < n && k < Kin) {
          v = A[row * Kin + k];
          if (Y) v *= act_mask(Y[row * Kin + k], act_in);
        }
        As[kq + j][r] = v;
      }
      // B tile: 16 k x 64 cols ; thread -> (col = t/4, k = (t%4)*4 + j) when sbk == 1 (contiguous in k),
      // else (k = t/16, col = (t%16)*4 + j) (contiguous in col)
      if (sbk == 1) {
        const int c = t / 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + kq + j;
          float v = 0.0f;
          if (col0 + c < N && k < Kin) v = __ldg(Bp + (int64_t)k * sbk + (int64_t)(col0 + c) * sbc);
          Bs[kq + j][c] = v;
        }
      } else {
        const int kk = t / 16, cq = (t % 16) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = k0 + kk, c = col0 + cq + j;
          float v = 0.0f;
          if (c < N && k < Kin) v = __ldg(Bp + (int64_t)k * sbk + (int64_t)c * sbc);
          Bs[kk][cq + j] = v;
        }
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t row = row0 + ty * 4 + i;
    if (row >= n) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = col0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias) v += __ldg(bias + c);
      C[row * N + c] = act_fwd(v, act_out);
    }
  }
}

// Skinny output (N <= 8): one thread per row, weights in shared memory.
template <int NMAX>
__global__ void __launch_bounds__(256) gemm_skinny_kernel(const float* __restrict__ A, const float* __restrict__ W /*[N,Kin]*/,
                                                          const float* __restrict__ bias, int act_out, float* __restrict__ C,
                                                          int64_t n, const int32_t* __restrict__ n_dev, int Kin, int N) {
  n = resolve_n(n, n_dev);
  extern __shared__ float sm[];
  float* Ws = sm;                       // [N][Kin]
  float* Xs = sm + NMAX * Kin;          // [256][33] staging of a 32-wide k chunk
  for (int i = threadIdx.x; i < N * Kin; i += blockDim.x) Ws[i] = W[i];
  const int64_t row0 = (int64_t)blockIdx.x * 256;
  if (row0 >= n) return;
  float acc[NMAX];
#pragma unroll
  for (int j = 0; j < NMAX; ++j) acc[j] = 0.0f;
  __syncthreads();
  for (int k0 = 0; k0 < Kin; k0 += 32) {
    // coalesced load of [256 rows][32 k]
    for (int e = threadIdx.x; e < 256 * 32; e += 256) {
      const int r = e / 32, k = e % 32;
      const int64_t row = row0 + r;
      Xs[r * 33 + k] = (row < n && k0 + k < Kin) ? A[row * Kin + k0 + k] : 0.0f;
    }
    __syncthreads();
    const int kmax = min(32, Kin - k0);
    for (int k = 0; k < kmax; ++k) {
      const float xv = Xs[threadIdx.x * 33 + k];
#pragma unroll
      for (int j = 0; j < NMAX; ++j)
        if (j < N) acc[j] = fmaf(xv, Ws[j * Kin + k0 + k], acc[j]);
    }
    __syncthreads();
  }
  const int64_t row = row0 + threadIdx.x;
  if (row < n) {
#pragma unroll
    for (int j = 0; j < NMAX; ++j)
      if (j < N) {
        float v = acc[j];
        if (bias) v += bias[j];
        C[row * N + j] = act_fwd(v, act_out);
      }
  }
}

// gW[M,K] += (gy .* mask(y))^T x over a chunk of rows; one 64x64 tile of gW per block.
__global__ void __launch_bounds__(256) wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ y, int act,
                                                    const float* __restrict__ x, float* __restrict__ gW, int64_t n,
                                                    const int32_t* __restrict__ n_dev, int K, int M, int64_t rows_per_chunk) {
  n = resolve_n(n, n_dev);
  const int tilesK = (K + BN - 1) / BN;
  const int m0 = (blockIdx.x / tilesK) * BM, k0 = (blockIdx.x % tilesK) * BN;
  const int64_t r_begin = (int64_t)blockIdx.y * rows_per_chunk;
  const int64_t r_end = min(n, r_begin + rows_per_chunk);
  if (r_begin >= r_end) return;
  __shared__ float Gs[BK][BM + 4];
  __shared__ float Xs[BK][BN + 4];
  const int t = threadIdx.x, ty = t / 16, tx = t % 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (int64_t r0 = r_begin; r0 < r_end; r0 += BK) {
    const int rr = t / 16, cq = (t % 16) * 4;
    const int64_t row = r0 + rr;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int m = m0 + cq + j, k = k0 + cq + j;
      float g = 0.0f, xv = 0.0f;
      if (row < r_end) {
        if (m < M) {
          g = gy[row * M + m];
          if (act) g *= act_mask(y[row * M + m], act);
        }
        if (k < K) xv = x[row * K + k];
      }
      Gs[rr][cq + j] = g;
      Xs[rr][cq + j] = xv;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Gs[kk][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[kk][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tx * 4 + j;
      if (k < K && acc[i][j] != 0.0f) atomicAdd(gW + (int64_t)m * K + k, acc[i][j]);
    }
  }
}

// gb[m] += sum_rows (gy .* mask(y))[r, m]
__global__ void __launch_bounds__(256) bgrad_kernel(const float* __restrict__ gy, const float* __restrict__ y, int act,
                                                    float* __restrict__ gb, int64_t n, const int32_t* __restrict__ n_dev, int M,
                                                    int64_t rows_per_block) {
  n = resolve_n(n, n_dev);
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  if (r_begin >= r_end) return;
  // thread -> column m = t % M' striding rows; M <= 256 handled by looping columns
  for (int m = threadIdx.x % 32; m < M; m += 32) {
    float s = 0.0f;
    for (int64_t r = r_begin + threadIdx.x / 32; r < r_end; r += blockDim.x / 32) {
      float g = gy[r * M + m];
      if (act) g *= act_mask(y[r * M + m], act);
      s += g;
    }
    if (s != 0.0f) atomicAdd(gb + m, s);
  }
}

// Skinny layer (M <= 8 outputs, e.g. the 128 -> 3 colour head) backward in ONE pass, exact fp32: thread <-> input
// column k, rows streamed in chunks of 32 whose masked output gradients gm = gy .* act'(y) sit in shared memory;
// gx[r,k] = sum_m gm[r,m] W[m,k] is written coalesced and gW[m,k] / gb[m] accumulate in registers (one atomic per
// block at the end).  Replaces a padded K=16 tensor-core GEMM plus a 128-row-padded weight-gradient GEMM.
template <int MMAX>
__global__ void __launch_bounds__(256) skinny_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ y, int act,
                                                         const float* __restrict__ x, const float* __restrict__ W, float* __restrict__ gx,
                                                         float* __restrict__ gW, float* __restrict__ gb, int64_t n,
                                                         const int32_t* __restrict__ n_dev, int K, int M, int64_t rows_per_block) {
  n = resolve_n(n, n_dev);
  const int64_t r_begin = (int64_t)blockIdx.x * rows_per_block;
  const int64_t r_end = min(n, r_begin + rows_per_block);
  if (r_begin >= r_end) return;
  __shared__ float gm[32][MMAX];
  const int t = threadIdx.x;
  for (int k0 = 0; k0 < K; k0 += blockDim.x) {
    const int k = k0 + t;
    float w[MMAX], acc[MMAX], accb = 0.0f;
#pragma unroll
    for (int m = 0; m < MMAX; ++m) {
      w[m] = (m < M && k < K) ? __ldg(W + (int64_t)m * K + k) : 0.0f;
      acc[m] = 0.0f;
    }
    for (int64_t r0 = r_begin; r0 < r_end; r0 += 32) {
      __syncthreads();
      if (t < 32 * MMAX) {
        const int rr = t / MMAX, m = t % MMAX;
        const int64_t row = r0 + rr;
        float g = 0.0f;
        if (row < r_end && m < M) {
          g = gy[row * M + m];
          if (act) g *= act_mask(y[row * M + m], act);
        }
        gm[rr][m] = g;
      }
      __syncthreads();
      const int rows = (int)min((int64_t)32, r_end - r0);
      if (k < K) {
#pragma unroll 4
        for (int rr = 0; rr < rows; ++rr) {
          const float xv = x ? x[(r0 + rr) * K + k] : 0.0f;
          float s = 0.0f;
#pragma unroll
          for (int m = 0; m < MMAX; ++m) {
            const float g = gm[rr][m];
            s = fmaf(g, w[m], s);
            acc[m] = fmaf(g, xv, acc[m]);
          }
          if (gx) gx[(r0 + rr) * K + k] = s;
        }
      }
      if (k0 == 0 && t < M && gb) {
        for (int rr = 0; rr < rows; ++rr) accb += gm[rr][t];
      }
    }
    if (k < K && gW) {
#pragma unroll
      for (int m = 0; m < MMAX; ++m)
        if (m < M && acc[m] != 0.0f) atomicAdd(gW + (int64_t)m * K + k, acc[m]);
    }
    if (k0 == 0 && t < M && gb && accb != 0.0f) atomicAdd(gb + t, accb);
  }
}

// ---- positional encoding ---------------------------------------------------------------------
__global__ void pe_concat_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, int64_t n, const int32_t* __restrict__ n_dev,
                                     int D, int pe) {
  n = resolve_n(n, n_dev);
  const int W = D + 2 * D * pe;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * W; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / W;
    const int c = (int)(t % W);
    float v;
    if (c < D) {
      v = x[i * D + c];
    } else {
      int e = c - D;
      const bool is_cos = e >= D * pe;
      if (is_cos) e -= D * pe;
      const float a = FFB_MUL(x[i * D + e / pe], (float)(1 << (e % pe)));
      v = is_cos ? cosf(a) : sinf(a);
    }
    out[t] = v;
  }
}

__global__ void pe_concat_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, float* __restrict__ gx, int64_t n,
                                     const int32_t* __restrict__ n_dev, int D, int pe) {
  n = resolve_n(n, n_dev);
  const int W = D + 2 * D * pe;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * D; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = t / D;
    const int d = (int)(t % D);
    const float xv = x[t];
    const float* gr = g + i * W;
    float s = gr[d];
    for (int k = 0; k < pe; ++k) {
      const float f = (float)(1 << k);
      const float a = FFB_MUL(xv, f);
      s += (gr[D + d * pe + k] * cosf(a) - gr[D + D * pe + d * pe + k] * sinf(a)) * f;
    }
    gx[t] = s;
  }
}

// MLPRender_Fea input: [features(C), viewdirs(3), PE(features, feape), PE(viewdirs, viewpe)]   (FactorFields.py:190-196)
// One work item = (row, input channel): ONE sincosf of the channel value, the higher octaves sin/cos(2^k v) by exact
// angle doubling (sin 2a = 2 sa ca, cos 2a = (ca - sa)(ca + sa); each step at most doubles the rounding error: <= 2^5
// ulp ~ 4e-6 absolute for viewpe = 6, far inside the 1e-4 bar) instead of 2*pe separate sinf / cosf range reductions.
// A CTA assembles RIN_ROWS rows in shared memory and writes them out as one contiguous, fully coalesced block.
constexpr int RIN_ROWS = 16;
__global__ void __launch_bounds__(256) render_input_fwd_kernel(const float* __restrict__ feat, int ld_feat, const float* __restrict__ rays,
                                                               const int32_t* __restrict__ ray_id, const int32_t* __restrict__ app_idx,
                                                               float* __restrict__ out, int64_t n, const int32_t* __restrict__ n_dev, int C,
                                                               int viewpe, int feape) {
  extern __shared__ float rin[];
  n = resolve_n(n, n_dev);
  const int W = 3 + C + 6 * viewpe + 2 * feape * C;
  const int oV = C, oPF = C + 3, oPV = C + 3 + 2 * feape * C, nch = C + 3;
  for (int64_t j0 = (int64_t)blockIdx.x * RIN_ROWS; j0 < n; j0 += (int64_t)gridDim.x * RIN_ROWS) {
    const int rows = (int)((n - j0) < RIN_ROWS ? (n - j0) : RIN_ROWS);
    for (int it = threadIdx.x; it < rows * nch; it += blockDim.x) {
      const int r = it / nch, ch = it % nch;
      const int64_t i = app_idx ? app_idx[j0 + r] : (j0 + r);
      float* o = rin + r * W;
      float v;
      int pe, o_sin, o_cos;
      if (ch < C) {
        v = feat[i * ld_feat + 1 + ch];
        o[ch] = v;
        pe = feape; o_sin = oPF + ch * feape; o_cos = oPF + C * feape + ch * feape;
      } else {
        const int d = ch - C;
        const int64_t rr = ray_id ? ray_id[i] : i;
        v = rays[rr * 6 + 3 + d];
        o[oV + d] = v;
        pe = viewpe; o_sin = oPV + d * viewpe; o_cos = oPV + 3 * viewpe + d * viewpe;
      }
      float sa, ca;
      sincosf(v, &sa, &ca);
      for (int k = 0; k < pe; ++k) {
        o[o_sin + k] = sa;
        o[o_cos + k] = ca;
        const float s2 = 2.0f * sa * ca, c2 = (ca - sa) * (ca + sa);
        sa = s2;
        ca = c2;
      }
    }
    __syncthreads();
    float* dst = out + j0 * W;
    for (int t = threadIdx.x; t < rows * W; t += blockDim.x) dst[t] = rin[t];
    __syncthreads();
  }
}

// g_app == nullptr: the gradient is ADDED into row app_idx[j] of the dense g_feat [*, ld_feat]; otherwise it is WRITTEN to row j of
// the compact g_app [n, ld_app] (column 0, the density gradient, is not touched: it travels separately)
__global__ void render_input_bwd_kernel(const float* __restrict__ feat, int ld_feat, const int32_t* __restrict__ app_idx,
                                        const float* __restrict__ g_in, int ld_gin, float* __restrict__ g_feat, int64_t n,
                                        const int32_t* __restrict__ n_dev, int C, int viewpe, int feape,
                                        float* __restrict__ g_app = nullptr, int ld_app = 0) {
  n = resolve_n(n, n_dev);
  const int W = ld_gin > 0 ? ld_gin : 3 + C + 6 * viewpe + 2 * feape * C;      // row stride of g_in
  const int oPF = C + 3;
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * C; t += (int64_t)gridDim.x * blockDim.x) {
    const int64_t j = t / C;
    const int c = (int)(t % C);
    const int64_t i = app_idx ? app_idx[j] : j;
    const float* gr = g_in + j * W;
    const float xv = feat[i * ld_feat + 1 + c];
    float s = gr[c];
    float sa, ca;
    sincosf(xv, &sa, &ca);                 // octaves by angle doubling, as in the forward kernel
    for (int k = 0; k < feape; ++k) {
      const float f = (float)(1 << k);
      s += (gr[oPF + c * feape + k] * ca - gr[oPF + C * feape + c * feape + k] * sa) * f;
      const float s2 = 2.0f * sa * ca, c2 = (ca - sa) * (ca + sa);
      sa = s2;
      ca = c2;
    }
    if (g_app) g_app[j * ld_app + 1 + c] = s;
    else g_feat[i * ld_feat + 1 + c] += s;
  }
}

}  // namespace ffb

using namespace ffb;

extern "C" {
int ffb_linear_tc_eligible(int32_t K, int32_t N);
int ffb_linear_tc_wgrad_eligible(int32_t K, int32_t M);
int ffb_linear_tc_fwd(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                      int32_t act, void* stream);
int ffb_linear_tc_bwd_input(const float* gy, const float* y, const float* W, float* gx, int64_t n, const int32_t* n_dev, int32_t K,
                            int32_t M, int32_t act, void* stream);
int ffb_linear_tc_bwd_weight(const float* gy, const float* y, int32_t act, const float* x, float* gW, float* gb, int64_t n,
                             const int32_t* n_dev, int32_t K, int32_t M, void* stream);

int ffb_linear_tc_fwd_ex(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                         int32_t act, int32_t split_terms, void* stream);

int ffb_linear_fwd(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K,
                   int32_t M, int32_t act, void* stream) {
  return ffb_linear_fwd_ex(x, W, b, y, n, n_dev, K, M, act, 3, stream);
}

int ffb_linear_fwd_ex(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K,
                      int32_t M, int32_t act, int32_t split_terms, void* stream) {
  FFB_REQUIRE(x && W && y && K > 0 && M > 0, "bad argument");
  if (n <= 0) return FFB_OK;
  const bool skinny = M <= 8 && (size_t)(8 * K + 256 * 33) * sizeof(float) <= 96 * 1024;   // exact fp32 and cheaper than a padded MMA
  if (!skinny && n >= 1024 && ffb_linear_tc_eligible(K, M)) return ffb_linear_tc_fwd_ex(x, W, b, y, n, n_dev, K, M, act, split_terms, stream);
  cudaStream_t s = (cudaStream_t)stream;
  if (skinny) {
    const size_t smem = (size_t)(8 * K + 256 * 33) * sizeof(float);
    static PerDeviceOnce attr_done;
    if (attr_done.first()) {
      FFB_CUDA(cudaFuncSetAttribute(gemm_skinny_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    }
    gemm_skinny_kernel<8><<<blocks_for(n, 256), 256, smem, s>>>(x, W, b, act, y, n, n_dev, K, M);
  } else {
    dim3 grid(blocks_for(n, BM), (M + BN - 1) / BN);
    gemm_rows_kernel<<<grid, 256, 0, s>>>(x, nullptr, 0, W, 1, K, b, act, y, n, n_dev, K, M);
  }
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_linear_bwd_input(float* gy, const float* y, const float* W, float* gx, int64_t n, const int32_t* n_dev, int32_t K,
                         int32_t M, int32_t act, void* stream) {
  FFB_REQUIRE(gy && W && K > 0 && M > 0 && (act == 0 || y), "bad argument");
  if (n <= 0 || !gx) return FFB_OK;
  if (n >= 1024 && ffb_linear_tc_eligible(M, K)) return ffb_linear_tc_bwd_input(gy, y, W, gx, n, n_dev, K, M, act, stream);
  cudaStream_t s = (cudaStream_t)stream;
  // gx[n,K] = (gy .* mask)[n,M] * W[M,K] : inner = M, B(kin=m, col=k) = W[m*K + k]
  dim3 grid(blocks_for(n, BM), (K + BN - 1) / BN);
  gemm_rows_kernel<<<grid, 256, 0, s>>>(gy, act ? y : nullptr, act, W, K, 1, nullptr, 0, gx, n, n_dev, M, K);
  FFB_LAUNCHED();
  return FFB_OK;
}

// y/act describe the activation that followed this layer in the forward pass (mask applied on the fly).
int ffb_linear_bwd_weight_act(const float* gy, const float* y, int32_t act, const float* x, float* gW, float* gb, int64_t n,
                              const int32_t* n_dev, int32_t K, int32_t M, void* stream) {
  FFB_REQUIRE(gy && x && gW && K > 0 && M > 0 && (act == 0 || y), "bad argument");
  if (n <= 0) return FFB_OK;
  if (n >= 1024 && ffb_linear_tc_wgrad_eligible(K, M)) return ffb_linear_tc_bwd_weight(gy, y, act, x, gW, gb, n, n_dev, K, M, stream);
  cudaStream_t s = (cudaStream_t)stream;
  const int tiles = ((M + BM - 1) / BM) * ((K + BN - 1) / BN);
  int64_t chunks = (4LL * sm_count() + tiles - 1) / tiles;
  int64_t rows = ((n + chunks - 1) / chunks + BK - 1) / BK * BK;
  if (rows < 256) rows = 256;
  chunks = (n + rows - 1) / rows;
  dim3 grid(tiles, (unsigned)chunks);
  wgrad_kernel<<<grid, 256, 0, s>>>(gy, y, act, x, gW, n, n_dev, K, M, rows);
  FFB_LAUNCHED();
  if (gb) {
    int64_t rpb = 256;
    bgrad_kernel<<<blocks_for(n, (int)rpb), 256, 0, s>>>(gy, y, act, gb, n, n_dev, M, rpb);
    FFB_LAUNCHED();
  }
  return FFB_OK;
}

int ffb_linear_bwd_skinny(const float* gy, const float* y, int32_t act, const float* x, const float* W, float* gx, float* gW, float* gb,
                          int64_t n, const int32_t* n_dev, int32_t K, int32_t M, void* stream) {
  FFB_REQUIRE(gy && W && K > 0 && M > 0 && M <= 8 && (act == 0 || y) && (x || !gW), "bad argument");
  if (n <= 0) return FFB_OK;
  const int threads = 256;   // also the size of the gm staging pass (32 rows x 8 outputs)
  const int64_t rows = 64;   // ~n/64 CTAs: enough resident warps to cover the streaming loads
  skinny_bwd_kernel<8><<<blocks_for(n, (int)rows), threads, 0, (cudaStream_t)stream>>>(gy, y, act, x, W, gx, gW, gb, n, n_dev, K, M, rows);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_linear_bwd_weight(const float* gy, const float* x, float* gW, float* gb, int64_t n, const int32_t* n_dev, int32_t K,
                          int32_t M, void* stream) {
  return ffb_linear_bwd_weight_act(gy, nullptr, 0, x, gW, gb, n, n_dev, K, M, stream);
}

int ffb_pe_concat_fwd(const float* x, float* out, int64_t n, const int32_t* n_dev, int32_t D, int32_t pe, void* stream) {
  FFB_REQUIRE(x && out && D > 0 && pe >= 0 && pe < 24, "bad argument");
  if (n <= 0) return FFB_OK;
  pe_concat_fwd_kernel<<<blocks_for(n * (D + 2 * D * pe), 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(x, out, n, n_dev, D, pe);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_pe_concat_bwd(const float* x, const float* g, float* gx, int64_t n, const int32_t* n_dev, int32_t D, int32_t pe, void* stream) {
  FFB_REQUIRE(x && g && gx && D > 0 && pe >= 0 && pe < 24, "bad argument");
  if (n <= 0) return FFB_OK;
  pe_concat_bwd_kernel<<<blocks_for(n * D, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(x, g, gx, n, n_dev, D, pe);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_render_input_fwd(const float* feat, int32_t ld_feat, const float* rays, const int32_t* ray_id, const int32_t* app_idx,
                         float* out, int64_t n, const int32_t* n_dev, int32_t C, int32_t viewpe, int32_t feape, void* stream) {
  FFB_REQUIRE(feat && rays && out && C > 0 && ld_feat >= C + 1, "bad argument");
  if (n <= 0) return FFB_OK;
  const int W = 3 + C + 6 * viewpe + 2 * feape * C;
  FFB_REQUIRE((size_t)RIN_ROWS * W * sizeof(float) <= 48 * 1024, "input row too wide");
  render_input_fwd_kernel<<<blocks_for(n, RIN_ROWS, sm_count() * 8), 256, (size_t)RIN_ROWS * W * sizeof(float), (cudaStream_t)stream>>>(
      feat, ld_feat, rays, ray_id, app_idx, out, n, n_dev, C, viewpe, feape);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_render_input_bwd_compact(const float* feat, int32_t ld_feat, const int32_t* app_idx, const float* g_in, int32_t ld_gin, float* g_app,
                                 int32_t ld_app, int64_t n, const int32_t* n_dev, int32_t C, int32_t viewpe, int32_t feape, void* stream) {
  FFB_REQUIRE(feat && g_in && g_app && C > 0 && ld_app >= C + 1, "bad argument");
  if (n <= 0) return FFB_OK;
  render_input_bwd_kernel<<<blocks_for(n * C, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(feat, ld_feat, app_idx, g_in, ld_gin, nullptr,
                                                                                              n, n_dev, C, viewpe, feape, g_app, ld_app);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_render_input_bwd(const float* feat, int32_t ld_feat, const int32_t* app_idx, const float* g_in, int32_t ld_gin, float* g_feat,
                         int64_t n, const int32_t* n_dev, int32_t C, int32_t viewpe, int32_t feape, void* stream) {
  FFB_REQUIRE(feat && g_in && g_feat && C > 0, "bad argument");
  if (n <= 0) return FFB_OK;
  render_input_bwd_kernel<<<blocks_for(n * C, 256, sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(feat, ld_feat, app_idx, g_in, ld_gin, g_feat,
                                                                                              n, n_dev, C, viewpe, feape);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
