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
#include "tc_common.cuh"
#include "ffb_math.h"

namespace ffb {

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

// Stage a [128 rows x nchunks*8 cols] fp32 block as TERMS bf16 operand copies.  Warp-cooperative: a warp owns an
// 8-row group at a time; lane = (rr = lane>>2 : row in the group, q = lane&3 : column pair 2q,2q+1 of an 8-wide
// chunk), so one load instruction touches 8 fully-used 32-byte sectors and one store instruction writes 128
// contiguous bytes (bank-conflict free).  Element (r, k) lands at
//     (r/8)*stride_g + (k/8)*stride_c + (r%8)*16 + (k%8)*2        (see the layout table at the top of the file)
template <int TERMS, class Load>
__device__ __forceinline__ void stage_block(uint8_t* dst, uint32_t part_bytes, uint32_t stride_g, uint32_t stride_c, int nchunks,
                                            int warp, int nwarps, int lane, Load load) {
  const int rr = lane >> 2, q = lane & 3;
  constexpr int BATCH = 8;     // independent loads in flight per lane
  for (int g = warp; g < 16; g += nwarps) {
    const int r = g * 8 + rr;
    uint8_t* base = dst + (uint32_t)g * stride_g + (uint32_t)rr * 16u + (uint32_t)q * 4u;
    for (int cb = 0; cb < nchunks; cb += BATCH) {
      float2 v[BATCH];
#pragma unroll
      for (int i = 0; i < BATCH; ++i) v[i] = (cb + i < nchunks) ? load(r, (cb + i) * 8 + 2 * q) : make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < BATCH; ++i) {
        if (cb + i < nchunks) {
#pragma unroll
          for (int t = 0; t < TERMS; ++t) {
            const __nv_bfloat16 a = __float2bfloat16_rn(v[i].x), b = __float2bfloat16_rn(v[i].y);
            v[i].x -= __bfloat162float(a);
            v[i].y -= __bfloat162float(b);
            __nv_bfloat162 ab = __halves2bfloat162(a, b);
            *reinterpret_cast<uint32_t*>(base + (uint32_t)t * part_bytes + (uint32_t)(cb + i) * stride_c) = *reinterpret_cast<uint32_t*>(&ab);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// C[n, N] = act_out( (A .* mask(Y))[n, K] * B^T + bias ),   B(j, k) = Bp[j*sbj + k*sbk]   (j < N, k < K)
// Persistent CTAs; the weight operand is staged once per CTA (TERMS bf16 parts), A is streamed per 128-row
// tile in K chunks of KC.  TERMS = 3 (6 MMAs per product, ~fp32 accuracy) for the forward pass, whose output
// feeds exp() in the compositor; TERMS = 2 (3 MMAs) for gradients.  The output tile goes TMEM -> registers ->
// shared memory -> coalesced 16-byte global stores.
// ---------------------------------------------------------------------------------------------------------
template <int TERMS>
__global__ void __launch_bounds__(512) tc_gemm_rows_kernel(const float* __restrict__ A, const float* __restrict__ Y, int act_in,
                                                           const float* __restrict__ Bp, int64_t sbj, int64_t sbk,
                                                           const float* __restrict__ bias, int act_out, float* __restrict__ C,
                                                           int64_t n, const int32_t* __restrict__ n_dev, int K, int N, int Kp, int Np,
                                                           int KC, int CB, uint32_t szU, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  n = resolve_n(n, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NT = blockDim.x, nwarps = NT >> 5;
  const uint32_t sboB = (uint32_t)(Kp / 8) * 128u;            // bytes between 8-row groups of the weight operand
  const uint32_t sboA = (uint32_t)(KC / 8) * 128u;            // ... of one A chunk
  const uint32_t szB = (uint32_t)Np * Kp * 2, szA = (uint32_t)128 * KC * 2;
  uint8_t* sB = smem;                                         // [TERMS][Np x Kp]
  uint8_t* sA = sB + TERMS * szB;                             // [TERMS][128 x KC], re-used as the fp32 output staging tile
  float* sE = reinterpret_cast<float*>(sA);
  uint64_t* bar = reinterpret_cast<uint64_t*>(sA + szU);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 0) mbar_init(bar, 1);
  const int kchunks = Kp / 8;
  for (int item = tid; item < Np * kchunks; item += NT) {
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
  const bool vec2 = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 7) == 0) && (!Y || (reinterpret_cast<uintptr_t>(Y) & 7) == 0);
  const int pitchE = CB + 4;
  uint32_t phase = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    for (int k0 = 0; k0 < Kp; k0 += KC) {
      const int kc = min(KC, Kp - k0);                        // multiple of 16
      stage_block<TERMS>(sA, szA, sboA, 128u, kc / 8, warp, nwarps, lane, [&](int r, int k) {
        const int64_t row = row0 + r;
        const int kk = k0 + k;
        float2 v = make_float2(0.0f, 0.0f);
        if (row < n && kk < K) {
          const float* p = A + row * K + kk;
          if (vec2) {
            v = *reinterpret_cast<const float2*>(p);
            if (Y) {
              const float2 yy = *reinterpret_cast<const float2*>(Y + row * K + kk);
              v.x *= tc_act_mask(yy.x, act_in);
              v.y *= tc_act_mask(yy.y, act_in);
            }
          } else {
            v.x = p[0];
            if (kk + 1 < K) v.y = p[1];
            if (Y) {
              v.x *= tc_act_mask(Y[row * K + kk], act_in);
              if (kk + 1 < K) v.y *= tc_act_mask(Y[row * K + kk + 1], act_in);
            }
          }
        }
        return v;
      });
      proxy_fence();          // generic-proxy smem writes -> visible to the tensor-core (async) proxy
      __syncthreads();
      if (warp == 0) {        // whole warp enters; lane 0 issues, the others wait at __syncwarp (not inside try_wait, which
                              // would suspend the warp and delay the issuing lane)
        if (elect_one()) {
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
              umma_f16(tmem_d, make_desc(aBase + ta * szA + oa, 128, sboA), make_desc(bBase + tb * szB + ob, 128, sboB), idesc,
                       first ? 0u : 1u);
              first = false;
            }
        }
        umma_commit(bar);     // implies tcgen05.fence::before_thread_sync
        }
        __syncwarp();
      }
      mbar_wait(bar, phase);  // the chunk buffer may be overwritten only after the MMAs have read it
      phase ^= 1;
    }
    tc_fence_after();
    // ---- epilogue: thread <-> row (TMEM lane quarter = warp % 4); warps sharing a quarter split the columns;
    //      CB columns at a time through shared memory, then coalesced 16-byte stores
    const int rloc = (warp & 3) * 32 + lane, wsplit = warp >> 2, nsplit = nwarps >> 2;
    const uint32_t tbase = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    const int cbshift = CB == 64 ? 4 : (CB == 32 ? 3 : 2);    // log2(CB / 4)
    for (int cb0 = 0; cb0 < Np; cb0 += CB) {
      for (int c0 = wsplit * 16; c0 < CB && cb0 + c0 < Np; c0 += nsplit * 16) {
        float v[16];
        tmem_ld16(tbase + (uint32_t)(cb0 + c0), v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int c = cb0 + c0 + i;
          float o = v[i];
          if (bias && c < N) o += __ldg(bias + c);
          v[i] = tc_act_fwd(o, act_out);
        }
#pragma unroll
        for (int i = 0; i < 16; i += 4)
          *reinterpret_cast<float4*>(sE + rloc * pitchE + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      }
      __syncthreads();
      const int cb = min(CB, N - cb0);                        // valid columns in this block (may be <= 0 for padding)
      if (cb > 0) {
        if ((N & 3) == 0) {
          for (int idx = tid; idx < (128 << cbshift); idx += NT) {
            const int r = idx >> cbshift, c4 = idx & ((1 << cbshift) - 1);
            if (c4 * 4 < cb && row0 + r < n)
              *reinterpret_cast<float4*>(C + (row0 + r) * N + cb0 + c4 * 4) = *reinterpret_cast<const float4*>(sE + r * pitchE + c4 * 4);
          }
        } else {
          for (int idx = tid; idx < 128 * CB; idx += NT) {
            const int r = idx >> (cbshift + 2), c = idx & (CB - 1);
            if (c < cb && row0 + r < n) C[(row0 + r) * N + cb0 + c] = sE[r * pitchE + c];
          }
        }
      }
      __syncthreads();        // sE (= the A chunk buffer) is free again
    }
    tc_fence_before();
    __syncthreads();          // all TMEM reads complete before the next tile's first MMA overwrites the accumulator
  }
  if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

// ---------------------------------------------------------------------------------------------------------
// gW[M, K] += (gy .* mask(y))^T x ;  gb[M] += column sums (through an extra all-ones column of x).
// D[128 (m, zero padded) x Np] accumulates in TMEM over all row tiles of this CTA; one atomic update at the end.
// Both operands are MN-major (the reduction index = the sample row is the MMA K dimension).
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(512) tc_wgrad_kernel(const float* __restrict__ gy, const float* __restrict__ y, int act,
                                                       const float* __restrict__ x, float* __restrict__ gW, float* __restrict__ gb,
                                                       int64_t n, const int32_t* __restrict__ n_dev, int K, int M, int Np, int tmem_cols) {
  extern __shared__ __align__(128) uint8_t smem[];
  n = resolve_n(n, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, NT = blockDim.x, nwarps = NT >> 5;
  constexpr uint32_t SBO = 16 * 128;                          // bytes between 8-wide MN chunks: 128 rows = 16 k-groups of 128 B
  constexpr uint32_t szG = 128 * 256;
  const uint32_t szX = (uint32_t)Np * 256;
  uint8_t* sG = smem;                                         // [2][128 m x 128 r]
  uint8_t* sX = sG + 2 * szG;                                 // [2][Np k x 128 r]
  uint64_t* bar = reinterpret_cast<uint64_t*>(sX + 2 * szX);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  if (warp == 0) tmem_alloc(tmem_slot, (uint32_t)tmem_cols);
  if (tid == 0) mbar_init(bar, 1);
  // rows m >= M of the G operand stay zero for the whole kernel
  for (int i = tid; i < (int)(2 * szG / 16); i += NT) reinterpret_cast<uint4*>(sG)[i] = make_uint4(0, 0, 0, 0);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot;
  const uint32_t idesc = make_idesc(Np, 1, 1);
  const int64_t n_tiles = (n + 127) / 128;
  const int mchunks = (M + 7) / 8, xchunks = Np / 8;
  const bool vecG = ((M & 1) == 0) && ((reinterpret_cast<uintptr_t>(gy) & 7) == 0) && (!act || (reinterpret_cast<uintptr_t>(y) & 7) == 0);
  const bool vecX = ((K & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
  uint32_t phase = 0;
  bool any = false;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row0 = tile * 128;
    // element (row r, column m) -> (m/8)*SBO + (r/8)*128 + (r%8)*16 ... : here the *column* index is the MN index,
    // so stage_block is called with (stride_g = 128 for the row group, stride_c = SBO for the 8-wide column chunk)
    // and lanes split as (rr = row in group, q = column pair)  ->  offset rr*16 + q*4 is (k%8)*16 + (mn%8)*2.
    stage_block<2>(sG, szG, 128u, SBO, mchunks, warp, nwarps, lane, [&](int r, int m) {
      const int64_t row = row0 + r;
      float2 v = make_float2(0.0f, 0.0f);
      if (row < n && m < M) {
        const float* p = gy + row * M + m;
        if (vecG) {
          v = *reinterpret_cast<const float2*>(p);
          if (act) {
            const float2 yy = *reinterpret_cast<const float2*>(y + row * M + m);
            v.x *= tc_act_mask(yy.x, act);
            v.y *= tc_act_mask(yy.y, act);
          }
        } else {
          v.x = p[0];
          if (act) v.x *= tc_act_mask(y[row * M + m], act);
          if (m + 1 < M) {
            v.y = p[1];
            if (act) v.y *= tc_act_mask(y[row * M + m + 1], act);
          }
        }
      }
      return v;
    });
    stage_block<2>(sX, szX, 128u, SBO, xchunks, warp, nwarps, lane, [&](int r, int k) {
      const int64_t row = row0 + r;
      float2 v = make_float2(0.0f, 0.0f);
      if (row < n) {
        if (vecX && k + 1 < K) {
          v = *reinterpret_cast<const float2*>(x + row * K + k);
        } else {
          v.x = k < K ? x[row * K + k] : (k == K ? 1.0f : 0.0f);
          v.y = k + 1 < K ? x[row * K + k + 1] : (k + 1 == K ? 1.0f : 0.0f);
        }
      }
      return v;
    });
    proxy_fence();
    __syncthreads();
    if (warp == 0) {
      if (elect_one()) {
      tc_fence_after();
      const uint32_t gB = smem_u32(sG), xB = smem_u32(sX);
      for (int ks = 0; ks < 8; ++ks) {                        // 128 rows = 8 slices of K = 16
        const uint32_t o = (uint32_t)ks * 256u;
        const uint64_t dGh = make_desc(gB + o, 128, SBO), dGl = make_desc(gB + szG + o, 128, SBO);
        const uint64_t dXh = make_desc(xB + o, 128, SBO), dXl = make_desc(xB + szX + o, 128, SBO);
        umma_f16(tmem_d, dGh, dXh, idesc, (any || ks > 0) ? 1u : 0u);
        umma_f16(tmem_d, dGl, dXh, idesc, 1);
        umma_f16(tmem_d, dGh, dXl, idesc, 1);
      }
      umma_commit(bar);
      }
      __syncwarp();
    }
    any = true;
    mbar_wait(bar, phase);      // operands may be overwritten only after the MMAs have read them
    phase ^= 1;
  }
  tc_fence_after();
  if (any) {
    const int m = (warp & 3) * 32 + lane;
    const uint32_t tbase = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    for (int c0 = (warp >> 2) * 16; c0 < Np; c0 += (nwarps >> 2) * 16) {
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
static int max_smem_optin() { return smem_optin_bytes(); }

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_tensor_cores(int enabled) {
  g_tc_enabled = enabled ? 1 : 0;
  return FFB_OK;
}
int ffb_tensor_cores_enabled(void) { return g_tc_enabled; }

static bool tc_gemm_plan(int K, int N, int terms, int* Kp_, int* Np_, int* KC_, int* CB_, uint32_t* szU_, size_t* smem_) {
  const int Kp = (K + 15) / 16 * 16, Np = (N + 15) / 16 * 16;
  if (Np > 256) return false;
  const int CB = Np >= 64 ? 64 : (Np >= 32 ? 32 : 16);   // output column block staged through shared memory (last block may be partial)
  const size_t szE = (size_t)128 * (CB + 4) * 4;
  for (int KC = 64; KC >= 16; KC >>= 1) {
    const int kc = KC < Kp ? KC : Kp;
    size_t szU = (size_t)terms * 128 * kc * 2;
    if (szU < szE) szU = szE;
    const size_t smem = (size_t)terms * Np * Kp * 2 + szU + 64;
    if (smem <= (size_t)max_smem_optin()) {
      *Kp_ = Kp; *Np_ = Np; *KC_ = kc; *CB_ = CB; *szU_ = (uint32_t)szU; *smem_ = smem;
      return true;
    }
  }
  return false;
}

// returns 1 if the (K, N) layer shape fits the tcgen05 forward (3-term) and input-gradient (2-term) kernels
int ffb_linear_tc_eligible(int32_t K, int32_t N) {
  if (!g_tc_enabled || K < 1 || N < 1) return 0;
  int Kp, Np, KC, CB;
  uint32_t szU;
  size_t smem;
  return tc_gemm_plan(K, N, 3, &Kp, &Np, &KC, &CB, &szU, &smem) ? 1 : 0;
}

static int tc_gemm_launch(int terms, const float* A, const float* Y, int act_in, const float* Bp, int64_t sbj, int64_t sbk,
                          const float* bias, int act_out, float* C, int64_t n, const int32_t* n_dev, int K, int N, cudaStream_t s) {
  int Kp, Np, KC, CB;
  uint32_t szU;
  size_t smem;
  FFB_REQUIRE(tc_gemm_plan(K, N, terms, &Kp, &Np, &KC, &CB, &szU, &smem), "layer does not fit in shared memory");
  const int cols = pow2_cols(Np);
  int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  if (per_sm > 512 / cols) per_sm = 512 / cols;
  if (per_sm > 8) per_sm = 8;
  if (per_sm < 1) per_sm = 1;
  const int NT = per_sm >= 4 ? 256 : 512;
  if (per_sm * NT > 2048) per_sm = 2048 / NT;
  const int64_t tiles = (n + 127) / 128;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > tiles) grid = tiles;
  static PerDeviceOnce attr_done;
  if (attr_done.first()) {
    FFB_CUDA(cudaFuncSetAttribute(tc_gemm_rows_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
    FFB_CUDA(cudaFuncSetAttribute(tc_gemm_rows_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
  }
  if (terms == 3) {
    tc_gemm_rows_kernel<3><<<(unsigned)grid, NT, smem, s>>>(A, Y, act_in, Bp, sbj, sbk, bias, act_out, C, n, n_dev, K, N, Kp, Np, KC, CB, szU, cols);
  } else {
    tc_gemm_rows_kernel<2><<<(unsigned)grid, NT, smem, s>>>(A, Y, act_in, Bp, sbj, sbk, bias, act_out, C, n, n_dev, K, N, Kp, Np, KC, CB, szU, cols);
  }
  FFB_LAUNCHED();
  return FFB_OK;
}

// split_terms: 3 = fp32-class (6 MMAs per product), 2 = ~5e-6 relative (3 MMAs): enough wherever the output does not
// feed exp() — the appearance MLP.
int ffb_linear_tc_fwd_ex(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                         int32_t act, int32_t split_terms, void* stream) {
  FFB_REQUIRE(x && W && y, "null argument");
  FFB_REQUIRE(split_terms == 2 || split_terms == 3, "split_terms must be 2 or 3");
  FFB_REQUIRE(ffb_linear_tc_eligible(K, M), "layer shape not eligible for the tensor-core path");
  if (n <= 0) return FFB_OK;
  return tc_gemm_launch(split_terms, x, nullptr, 0, W, K, 1, b, act, y, n, n_dev, K, M, (cudaStream_t)stream);
}

int ffb_linear_tc_fwd(const float* x, const float* W, const float* b, float* y, int64_t n, const int32_t* n_dev, int32_t K, int32_t M,
                      int32_t act, void* stream) {
  return ffb_linear_tc_fwd_ex(x, W, b, y, n, n_dev, K, M, act, 3, stream);
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
  static PerDeviceOnce wattr_done;
  if (wattr_done.first()) {
    FFB_CUDA(cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem_optin()));
  }
  int per_sm = (int)((size_t)(220 * 1024) / (smem + 1024));
  if (per_sm > 512 / cols) per_sm = 512 / cols;
  if (per_sm > 4) per_sm = 4;
  if (per_sm < 1) per_sm = 1;
  const int NT = per_sm >= 4 ? 256 : 512;
  const int64_t tiles = (n + 127) / 128;
  int64_t grid = (int64_t)sm_count() * per_sm;
  if (grid > tiles) grid = tiles;
  tc_wgrad_kernel<<<(unsigned)grid, NT, smem, (cudaStream_t)stream>>>(gy, y, act, x, gW, gb, n, n_dev, K, M, Np, cols);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
