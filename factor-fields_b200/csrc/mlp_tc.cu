// mlp_tc.cu — decoder-MLP layers on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
// Replaces the nn.Linear call sites of MLPMixer / MLPRender_Fea (FactorFields.py:153-156,197-200) and their
// autograd for the hot layer shapes; linear.cu keeps the exact-fp32 SIMT kernels for everything else.
//
// Precision: the parity bar is 1e-4 relative in fp32, which single-pass bf16 / tf32 MMA misses.
// Every fp32 operand is split into bf16 hi + lo (x = hi + lo + O(2^-18 x)) and each product is issued as three
// kind::f16 MMAs (hi*hi + lo*hi + hi*lo) accumulating in fp32 in TMEM  ->  ~5e-6 relative error.
//
// Operands are staged by the CTA's threads (the fp32 -> bf16x2 split needs a register pass, so TMA cannot be
// used for them) into the canonical no-swizzle UMMA shared-memory layouts:
//   K-major  (rows x K, K contiguous):   off(r,k)  = (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2      [LBO = 128]
//   MN-major (MN x K, MN contiguous):    off(mn,k) = (mn/8)*SBO + (k/8)*128 + (k%8)*16 + (mn%8)*2    [LBO = 128]
// One elected thread issues the MMAs and commits them to an mbarrier; the four warps read their 32 TMEM lanes
// back with tcgen05.ld for the fused epilogue (bias + ReLU / sigmoid, or the atomic weight-gradient update).
#include <cuda_bf16.h>
#include "ffb_common.cuh"
#include "ffb_math.h"

namespace ffb {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive columns (fp32) of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | 1<<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, M=128, N=n; major bits: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void split8(const float x[8], uint4& hi, uint4& lo) {
  __nv_bfloat162 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * i]), h1 = __float2bfloat16_rn(x[2 * i + 1]);
    h[i] = __halves2bfloat162(h0, h1);
    l[i] = __halves2bfloat162(__float2bfloat16_rn(x[2 * i] - __bfloat162float(h0)), __float2bfloat16_rn(x[2 * i + 1] - __bfloat162float(h1)));
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]), *reinterpret_cast<uint32_t*>(&h[2]),
                  *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]), *reinterpret_cast<uint32_t*>(&l[2]),
                  *reinterpret_cast<uint32_t*>(&l[3]));
}

__device__ __forceinline__ float tc_act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.0f);
  if (act == 2) return sigmoid_f(v);
  return v;
}
__device__ __forceinline__ float tc_act_mask(float y, int act) {
  if (act == 1) return y > 0.0f ? 1.0f : 0.0f;
  if (act == 2) return y * (1.0f - y);
  return 1.0f;
}

// fp32 -> TERMS bf16 parts (x = p0 + p1 (+ p2) + O(2^-9*TERMS x)), 8 values -> one 16-byte chunk per part
template <int TERMS>
__device__ __forceinline__ void splitN(const float x[8], uint4 out[TERMS]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = x[i];
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat16 a = __float2bfloat16_rn(r[2 * i]), b = __float2bfloat16_rn(r[2 * i + 1]);
      r[2 * i] -= __bfloat162float(a);
      r[2 * i + 1] -= __bfloat162float(b);
      __nv_bfloat162 ab = __halves2bfloat162(a, b);
      w[i] = *reinterpret_cast<uint32_t*>(&ab);
    }
    out[t] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// ---------------------------------------------------------------------------------------------------------
// C[n, N] = act_out( (A .* mask(Y))[n, K] * B^T + bias ),   B(j, k) = Bp[j*sbj + k*sbk]   (j < N, k < K)
// Persistent CTAs; the weight operand is staged once per CTA (TERMS bf16 parts), A is streamed per 128-row
// tile in K chunks of KC.  TERMS = 3 (6 MMAs per product, ~fp32 accuracy) for the forward pass, whose output
// feeds exp() in the compositor; TERMS = 2 (3 MMAs) for gradients.
// ---------------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(128) tc_gemm_rows_kernel(const float* __restrict__ A, const float* __restrict__ Y, int act_in,
                                                           const float* __restrict__ Bp, int64_t sbj, int64_t sbk,
                                                           const float* __restrict__ bias, int act_out, float* __restrict__ C,
                                                           int64_t n, const int32_t* __restrict__ n_dev, int K, int N, int Kp, int Np,
                                                           int KC, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  n = resolve_n(n, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sboB = (uint32_t)(Kp / 8) * 128u;            // bytes between 8-row groups of the weight operand
  const uint32_t sboA = (uint32_t)(KC / 8) * 128u;            // ... of one A chunk
  const size_t szB = (size_t)Np * Kp * 2, szA = (size_t)128 * KC * 2;
  uint8_t* sB = smem;                                         // [TERMS][Np x Kp]
  uint8_t* sA = sB + TERMS * szB;                             // [TERMS][128 x KC]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sA + TERMS * szA);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 0) mbar_init(bar, 1);
  const int kchunks = Kp / 8;
  for (int item = tid; item < Np * kchunks; item += 128) {
    const int j = item % Np, c = item / Np;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c * 8 + i;
      x[i] = (j < N && k < K) ? __ldg(Bp + (int64_t)j * sbj + (int64_t)k * sbk) : 0.0f;
    }
    uint4 parts[TERMS];
    splitN<TERMS>(x, parts);
    const uint32_t off = (uint32_t)(j / 8) * sboB + (uint32_t)c * 128u + (uint32_t)(j % 8) * 16u;
#pragma unroll
    for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint4*>(sB + t * szB + off) = parts[t];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc(Np, 0, 0);
  const int64_t n_tiles = (n + 127) / 128;
  uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    for (int k0 = 0; k0 < Kp; k0 += KC) {
      const int kc = min(KC, Kp - k0);                        // multiple of 16
      // ---- stage the A chunk: item = (row r, 8-wide k chunk c); consecutive threads -> consecutive rows
      for (int item = tid; item < 128 * (kc / 8); item += 128) {
        const int r = item & 127, c = item >> 7;
        const int64_t row = row0 + r;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int k = k0 + c * 8 + i;
          float v = 0.0f;
          if (row < n && k < K) {
            v = A[row * K + k];
            if (Y) v *= tc_act_mask(Y[row * K + k], act_in);
          }
          x[i] = v;
        }
        uint4 parts[TERMS];
        splitN<TERMS>(x, parts);
        const uint32_t off = (uint32_t)(r / 8) * sboA + (uint32_t)c * 128u + (uint32_t)(r % 8) * 16u;
#pragma unroll
        for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint4*>(sA + t * szA + off) = parts[t];
      }
      proxy_fence();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncthreads();
      if (tid == 0) {
        tc_fence_after();
        const uint32_t aBase = smem_u32(sA), bBase = smem_u32(sB);
        for (int ks = 0; ks < kc / 16; ++ks) {
          const uint32_t oa = (uint32_t)ks * 256u, ob = (uint32_t)(k0 / 16 + ks) * 256u;   // two 8-wide k chunks per K=16 slice
          bool first = (k0 == 0 && ks == 0);
#pragma unroll
          for (int ta = 0; ta < TERMS; ++ta)
#pragma unroll
            for (int tb = 0; tb < TERMS; ++tb) {
              if (ta + tb >= TERMS) continue;                 // drop the terms below the fp32 rounding level
              umma_f16(tmem_d, make_desc(aBase + ta * (uint32_t)szA + oa, 128, sboA), make_desc(bBase + tb * (uint32_t)szB + ob, 128, sboB),
                       idesc, first ? 0u : 1u);
              first = false;
            }
        }
        umma_commit(bar);     // implies tcgen05.fence::before_thread_sync
      }
      mbar_wait(bar, phase);  // the chunk buffer may be overwritten only after the MMAs have read it
      phase ^= 1;
    }
    tc_fence_after();
    // ---- epilogue: thread <-> row (TMEM lane), 16 columns at a time
    const int64_t row = row0 + warp * 32 + lane;
    const uint32_t tbase = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < Np; c0 += 16) {
      float v[16];
      tmem_ld16(tbase + (uint32_t)c0, v);
      if (row < n) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = c0 + i;
          if (c < N) {
            float o = v[i];
            if (bias) o += __ldg(bias + c);
            v[i] = tc_act_fwd(o, act_out);
          }
        }
        float* dst = C + row * N + c0;
        if ((N & 3) == 0 && c0 + 16 <= N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
          for (int i = 0; i < 16 && c0 + i < N; ++i) dst[i] = v[i];
        }
      }
    }
    tc_fence_before();
    __syncthreads();        // all TMEM reads complete before the next tile's first MMA overwrites the accumulator
  }
  if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// gW[M, K] += (gy .* mask(y))^T x ;  gb[M] += column sums (through an extra all-ones column of x).
// D[128 (m, zero padded) x Np] accumulates in TMEM over all row tiles of this CTA; one atomic update at the end.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) tc_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ y, int act,
                                                       const float* __restrict__ x, float* __restrict__ gW, float* __restrict__ gb,
                                                       int64_t n, const int32_t* __restrict__ n_dev, int K, int M, int Np, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  n = resolve_n(n, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t SBO = 16 * 128;                          // bytes between 8-wide MN chunks: 128 rows = 16 k-groups of 128 B
  uint8_t* sGhi = smem;                                       // [128 m][128 r] MN-major
  uint8_t* sGlo = sGhi + 128 * 256;
  uint8_t* sXhi = sGlo + 128 * 256;                           // [Np k][128 r]
  uint8_t* sXlo = sXhi + (size_t)Np * 256;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sXlo + (size_t)Np * 256);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 0) mbar_init(bar, 1);
  // rows m >= M of the G operand stay zero for the whole kernel
  for (int i = tid; i < 128 * 256 / 16; i += 128) {
    reinterpret_cast<uint4*>(sGhi)[i] = make_uint4(0, 0, 0, 0);
    reinterpret_cast<uint4*>(sGlo)[i] = make_uint4(0, 0, 0, 0);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc(Np, 1, 1);
  const int64_t n_tiles = (n + 127) / 128;
  const int mchunks = (M + 7) / 8, xchunks = Np / 8;
  uint32_t phase = 0;
  bool any = false;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    for (int item = tid; item < 128 * mchunks; item += 128) {
      const int r = item & 127, mc = item >> 7;
      const int64_t row = row0 + r;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = mc * 8 + i;
        float g = 0.0f;
        if (row < n && m < M) {
          g = gy[row * M + m];
          if (act) g *= tc_act_mask(y[row * M + m], act);
        }
        v[i] = g;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      const uint32_t off = (uint32_t)mc * SBO + (uint32_t)(r / 8) * 128u + (uint32_t)(r % 8) * 16u;
      *reinterpret_cast<uint4*>(sGhi + off) = hi;
      *reinterpret_cast<uint4*>(sGlo + off) = lo;
    }
    for (int item = tid; item < 128 * xchunks; item += 128) {
      const int r = item & 127, kc = item >> 7;
      const int64_t row = row0 + r;
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int k = kc * 8 + i;
        float xv = 0.0f;
        if (row < n) xv = (k < K) ? x[row * K + k] : (k == K ? 1.0f : 0.0f);
        v[i] = xv;
      }
      uint4 hi, lo;
      split8(v, hi, lo);
      const uint32_t off = (uint32_t)kc * SBO + (uint32_t)(r / 8) * 128u + (uint32_t)(r % 8) * 16u;
      *reinterpret_cast<uint4*>(sXhi + off) = hi;
      *reinterpret_cast<uint4*>(sXlo + off) = lo;
    }
    proxy_fence();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t gH = smem_u32(sGhi), gL = smem_u32(sGlo), xH = smem_u32(sXhi), xL = smem_u32(sXlo);
      for (int ks = 0; ks < 8; ++ks) {                        // 128 rows = 8 slices of K = 16
        const uint32_t o = (uint32_t)ks * 256u;
        const uint64_t dGh = make_desc(gH + o, 128, SBO), dGl = make_desc(gL + o, 128, SBO);
        const uint64_t dXh = make_desc(xH + o, 128, SBO), dXl = make_desc(xL + o, 128, SBO);
        umma_f16(tmem_d, dGh, dXh, idesc, (any || ks > 0) ? 1u : 0u);
        umma_f16(tmem_d, dGl, dXh, idesc, 1);
        umma_f16(tmem_d, dGh, dXl, idesc, 1);
      }
      umma_commit(bar);
    }
    any = true;
    mbar_wait(bar, phase);      // operands may be overwritten only after the MMAs have read them
    phase ^= 1;
  }
  tc_fence_after();
  if (any) {
    const int m = warp * 32 + lane;
    const uint32_t tbase = tmem_d + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < Np; c0 += 16) {
      float v[16];
      tmem_ld16(tbase + (uint32_t)c0, v);
      if (m < M) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int k = c0 + i;
          if (v[i] != 0.0f) {
            if (k < K) atomicAdd(gW + (int64_t)m * K + k, v[i]);
            else if (k == K && gb) atomicAdd(gb + m, v[i]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

static int g_tc_enabled = 1;

static int pow2_cols(int n) {
  int c = 32;
  while (c < n) c <<= 1;
  return c;
}
static int max_smem_optin() {
  static int v = 0;
  if (!v) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  }
  return v;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_tensor_cores(int enabled) {
  g_tc_enabled = enabled ? 1 : 0;
  return FFB_OK;
}
int ffb_tensor_cores_enabled(void) { return g_tc_enabled; }

static bool tc_gemm_plan(int K, int N, int terms, int* Kp_, int* Np_, int* KC_, size_t* smem_) {
  const int Kp = (K + 15) / 16 * 16, Np = (N + 15) / 16 * 16;
  if (Np > 256) return false;
  for (int KC = 64; KC >= 16; KC >>= 1) {
    const int kc = KC < Kp ? KC : Kp;
    const size_t smem = (size_t)terms * ((size_t)Np * Kp * 2 + (size_t)128 * kc * 2) + 64;
    if (smem <= (size_t)max_smem_optin()) {
      *Kp_ = Kp; *Np_ = Np; *KC_ = kc; *smem_ = smem;
      return true;
    }
  }
  return false;
}

// returns 1 if the (K, N) layer shape fits the tcgen05 forward (3-term) and input-gradient (2-term) kernels
int ffb_linear_tc_eligible(int32_t K, int32_t N) {
  if (!g_tc_enabled || K < 1 || N < 1) return 0;
  int Kp, Np, KC;
  size_t smem;
  return tc_gemm_plan(K, N, 3, &Kp, &Np, &KC, &smem) ? 1 : 0;
}

static int tc_gemm_launch(int terms, const float* A, const float* Y, int act_in, const float* Bp, int64_t sbj, int64_t sbk,
                          const float* bias, int act_out, float* C, int64_t n, const int32_t* n_dev, int K, int N, cudaStream_t s) {
  int Kp, Np, KC;
  size_t smem;
  FFB_REQUIRE(tc_gemm_plan(K, N, terms, &Kp, &Np, &KC, &smem), "layer does not fit in shared memory");
  const int cols = pow2_cols(Np);
  int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  if (per_sm > 512 / cols) per_sm = 512 / cols;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = (n + 127) / 128;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > tiles) grid = tiles;
  if (terms == 3) {
    FFB_CUDA(cudaFuncSetAttribute(tc_gemm_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
    tc_gemm_rows_kernel<3><<<(unsigned)grid, 128, smem, s>>>(A, Y, act_in, Bp, sbj, sbk, bias, act_out, C, n, n_dev, K, N, Kp, Np, KC, cols);
  } else {
    FFB_CUDA(cudaFuncSetAttribute(tc_gemm_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
    tc_gemm_rows_kernel<2><<<(unsigned)grid, 128, smem, s>>>(A, Y, act_in, Bp, sbj, sbk, bias, act_out, C, n, n_dev, K, N, Kp, Np, KC, cols);
  }
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_linear_tc_fwd(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                      int32_t act, void* stream) {
  FFB_REQUIRE(x && W && y, "null argument");
  FFB_REQUIRE(ffb_linear_tc_eligible(K, M), "layer shape not eligible for the tensor-core path");
  if (n <= 0) return FFB_OK;
  return tc_gemm_launch(3, x, nullptr, 0, W, K, 1, b, act, y, n, n_dev, K, M, (cudaStream_t)stream);
}

int ffb_linear_tc_bwd_input(const float* gy, const float* y, const float* W, float* gx, int64_t n, const int32_t* n_dev, int32_t K,
                            int32_t M, int32_t act, void* stream) {
  FFB_REQUIRE(gy && W && gx && (act == 0 || y), "bad argument");
  FFB_REQUIRE(ffb_linear_tc_eligible(M, K), "layer shape not eligible for the tensor-core path");
  if (n <= 0) return FFB_OK;
  // gx[n, K] = (gy .* mask)[n, M] * W[M, K]:  inner dim = M, B(j = k_in, k = m) = W[m*K + j]
  return tc_gemm_launch(2, gy, act ? y : nullptr, act, W, 1, K, nullptr, 0, gx, n, n_dev, M, K, (cudaStream_t)stream);
}

int ffb_linear_tc_wgrad_eligible(int32_t K, int32_t M) {
  if (!g_tc_enabled || K < 1 || M < 1 || M > 128) return 0;
  const int Np = (K + 1 + 15) / 16 * 16;
  if (Np > 256) return 0;
  const size_t smem = (size_t)2 * 128 * 256 + (size_t)2 * Np * 256 + 64;
  return smem <= (size_t)max_smem_optin() ? 1 : 0;
}

int ffb_linear_tc_bwd_weight(const float* gy, const float* y, int32_t act, const float* x, float* gW, float* gb, int64_t n,
                             const int32_t* n_dev, int32_t K, int32_t M, void* stream) {
  FFB_REQUIRE(gy && x && gW && (act == 0 || y), "bad argument");
  FFB_REQUIRE(ffb_linear_tc_wgrad_eligible(K, M), "layer shape not eligible for the tensor-core weight-gradient path");
  if (n <= 0) return FFB_OK;
  const int Np = (K + 1 + 15) / 16 * 16;
  const size_t smem = (size_t)2 * 128 * 256 + (size_t)2 * Np * 256 + 64;
  const int cols = pow2_cols(Np);
  FFB_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
  int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  if (per_sm > 512 / cols) per_sm = 512 / cols;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int64_t tiles = (n + 127) / 128;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > tiles) grid = tiles;
  tc_wgrad_kernel<<<(unsigned)grid, 128, smem, (cudaStream_t)stream>>>(gy, y, act, x, gW, gb, n, n_dev, K, M, Np, cols);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
