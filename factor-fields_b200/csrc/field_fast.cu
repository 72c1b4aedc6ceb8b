// field_fast.cu — specialised field-query kernels for the grid-coefficient x grid-basis fields
// (nerf.yaml / sdf.yaml / image.yaml / image_set.yaml shapes): the headline HBM-roofline kernels K1 / K2.
//
// One thread per query, consecutive threads = consecutive samples of a ray (the compacted order), so that
// neighbouring lanes hit the same or neighbouring texels.  Channels-last texels are read as float2/float4
// vectors; the x-neighbour corners of a linear tap are adjacent in memory and fetched together.  The
// backward pass re-gathers (no saved activations) and scatters with vector reductions
// (red.global.add.v2/v4.f32).  Coordinate arithmetic is the shared unfused fp32 sequence of ffb_math.h, so
// tap indices are identical to the generic path and to the reference.
#include "field_fast.cuh"

namespace ffb {

// NT threads per CTA, at least MINB CTAs resident per SM (caps the register count: the kernel is bound by the latency
// of L2-resident gathers, so more resident warps = more loads in flight).  basis (optional): the concatenated basis
// row, saved for the gather-free backward pass.
// STAGE (knob "field_fwd_stage", on for W <= 32): a warp parks its 32 coefficient rows and its 32 feature rows in shared
// memory and streams the two contiguous 32*W-float chunks out as coalesced 16-byte pieces (one row per lane writes 8-byte
// pieces 4 W bytes apart: ~9x the LSU wavefronts).  A small win (278 -> 273 us) — an earlier version that streamed 8-byte
// pieces and formed the product on the way out was 6 % slower than the direct stores.
// LPAR (small batches: the regression drivers' 40-100 k points leave most of the 148 SMs without work at one thread per
// query): one thread per (query, level) — consecutive lanes take the levels of one query, so a query's row segments are
// still written by neighbouring lanes; the coefficient taps are recomputed per level (ALU only).
// The whole WC-channel coefficient row of a query in one pass over its 2^DC corners, each texel fetched as ALIGNED 16-byte
// pieces: a texel of WC = 18 floats (72 B) starts 16-byte aligned for even texel indices and 8 bytes past an aligned address for
// odd ones, so a window of 5 float4 starting at the aligned address below it covers the texel either way (2 floats of the
// neighbouring texel ride along: always inside the tensor when the texel count is even).  40 loads per query instead of the
// 72 8-byte loads of the per-level gathers; same corner order and weights, so the sums are bit-identical.
template <int DC, int WC>
__device__ __forceinline__ void coeff_row_windows(const float* __restrict__ cdata, const TapSet<DC, false>& t, float acc[WC]) {
  static_assert((WC & 1) == 0 && (WC % 4) == 2, "window scheme written for rows of 4k + 2 floats (8-byte aligned texels)");
  constexpr int ROWS = 1 << (DC - 1), NQ = (WC + 2) / 4;
#pragma unroll
  for (int c = 0; c < WC; ++c) acc[c] = 0.0f;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (!t.row_ok[r]) continue;
#pragma unroll
    for (int xs = 0; xs < 2; ++xs) {
      const float wx = xs ? t.wx1 : t.wx0;
      if (xs ? !t.x1_ok : (wx == 0.0f)) continue;
      const float w = t.wrow[r] * wx;
      const size_t e0 = (size_t)(t.base[r] + xs) * WC;
      const bool odd = (e0 & 2) != 0;
      const float4* p = reinterpret_cast<const float4*>(cdata + (e0 - (odd ? 2 : 0)));
      float win[NQ * 4];
#pragma unroll
      for (int k = 0; k < NQ; ++k) {
        const float4 q = __ldg(p + k);
        win[4 * k] = q.x; win[4 * k + 1] = q.y; win[4 * k + 2] = q.z; win[4 * k + 3] = q.w;
      }
#pragma unroll
      for (int c = 0; c < WC; ++c) acc[c] += (odd ? win[c + 2] : win[c]) * w;
    }
  }
}

template <int DB, int DC, bool NEAR_B, bool NEAR_C, int NT, int MINB, bool STAGE, bool LPAR = false, int CROW = 0>
__global__ void __launch_bounds__(NT, MINB) fast_fwd_kernel(const FastParams P, const float* __restrict__ x, int64_t n,
                                                            const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                            float* __restrict__ coeff, float* __restrict__ basis) {
  extern __shared__ float2 fwd_stage[];
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  const int lane = threadIdx.x & 31, W = P.W;
  float* sC = nullptr;
  float* sB = nullptr;
  if (STAGE) {
    sC = reinterpret_cast<float*>(fwd_stage) + (size_t)(threadIdx.x >> 5) * 64 * W;
    sB = sC + 32 * W;
  }
  const int64_t n_items = LPAR ? n * P.n_levels : n;
  const int64_t n_chunks = (n_items + 31) / 32;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_chunks; k += nwarps) {
    const int64_t item = k * 32 + lane;
    const int64_t i = LPAR ? item / P.n_levels : item;
    const int l_begin = LPAR ? (int)(item % P.n_levels) : 0, l_end = LPAR ? l_begin + 1 : P.n_levels;
    if (item < n_items) {
      float xr[3];
      for (int d = 0; d < P.xdim; ++d) xr[d] = x[i * P.xdim + d];
      TapSet<DC, NEAR_C> tc;
      coeff_taps<DC, NEAR_C>(P, xr, tc);
      // STAGE: the coefficient and feature rows are parked in shared memory (sC, sB) and leave as coalesced 16-byte pieces
      float* frow = STAGE ? sB + lane * W : (feats ? feats + i * W : nullptr);
      float* crow = STAGE ? sC + lane * W : (coeff ? coeff + i * W : nullptr);
      float* brow = nullptr;                              // the saved basis row is written blocked, below
      if constexpr (CROW > 0) {                           // CROW = W: the coefficient row first, by aligned 16-byte windows
        float acc[CROW > 0 ? CROW : 2];
        coeff_row_windows<DC, (CROW > 0 ? CROW : 2)>(P.cdata, tc, acc);
#pragma unroll
        for (int c = 0; c < CROW; c += 2) *reinterpret_cast<float2*>(crow + c) = make_float2(acc[c], acc[c + 1]);
      }
      for (int l = l_begin; l < l_end; ++l) {
        const FastLevel L = P.lv[l];
        TapSet<DB, NEAR_B> tb;
        basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
        if ((L.C & 3) == 0) {
          for (int c0 = 0; c0 < L.C; c0 += 4) {
            float b[4], ca[2], cb[2];
            gather_vec<DB, NEAR_B, 4>(L.data, L.C, c0, tb, b);
            const int o = L.col + c0;
            if (CROW > 0) {
              const float2 t0 = *reinterpret_cast<const float2*>(crow + o), t1 = *reinterpret_cast<const float2*>(crow + o + 2);
              ca[0] = t0.x; ca[1] = t0.y; cb[0] = t1.x; cb[1] = t1.y;
            } else {
              gather_vec<DC, NEAR_C, 2>(P.cdata, W, L.col + c0, tc, ca);
              gather_vec<DC, NEAR_C, 2>(P.cdata, W, L.col + c0 + 2, tc, cb);
            }
            if (frow) {
              *reinterpret_cast<float2*>(frow + o) = make_float2(b[0] * ca[0], b[1] * ca[1]);
              *reinterpret_cast<float2*>(frow + o + 2) = make_float2(b[2] * cb[0], b[3] * cb[1]);
            }
            if (crow && CROW == 0) {
              *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
              *reinterpret_cast<float2*>(crow + o + 2) = make_float2(cb[0], cb[1]);
            }
            if (brow) {
              *reinterpret_cast<float2*>(brow + o) = make_float2(b[0], b[1]);
              *reinterpret_cast<float2*>(brow + o + 2) = make_float2(b[2], b[3]);
            } else if (basis) {
#pragma unroll
              for (int j = 0; j < 4; ++j) basis[blk_idx(i, o + j, W)] = b[j];
            }
          }
        } else {
          for (int c0 = 0; c0 < L.C; c0 += 2) {
            float b[2], ca[2];
            gather_vec<DB, NEAR_B, 2>(L.data, L.C, c0, tb, b);
            const int o = L.col + c0;
            if (CROW > 0) {
              const float2 t0 = *reinterpret_cast<const float2*>(crow + o);
              ca[0] = t0.x; ca[1] = t0.y;
            } else {
              gather_vec<DC, NEAR_C, 2>(P.cdata, W, L.col + c0, tc, ca);
            }
            if (frow) *reinterpret_cast<float2*>(frow + o) = make_float2(b[0] * ca[0], b[1] * ca[1]);
            if (crow && CROW == 0) *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
            if (brow) *reinterpret_cast<float2*>(brow + o) = make_float2(b[0], b[1]);
            else if (basis) { basis[blk_idx(i, o, W)] = b[0]; basis[blk_idx(i, o + 1, W)] = b[1]; }
          }
        }
      }
    }
    if (STAGE) {
      __syncwarp();
      const int64_t rows = (n - k * 32) < 32 ? (n - k * 32) : 32;
      const int total = (int)(rows * W) >> 2;                 // 16-byte pieces of the chunk (32 W floats: a multiple of 4)
      const int64_t base = k * 8 * W;                         // = k * 32 * W / 4
      const float4* c4 = reinterpret_cast<const float4*>(sC);
      const float4* f4 = reinterpret_cast<const float4*>(sB);
      for (int t = lane; t < total; t += 32) {
        if (coeff) reinterpret_cast<float4*>(coeff)[base + t] = c4[t];
        if (feats) reinterpret_cast<float4*>(feats)[base + t] = f4[t];
      }
      for (int t = (total << 2) + lane; t < (int)(rows * W); t += 32) {     // tail of a partial last chunk
        if (coeff) coeff[k * 32 * W + t] = sC[t];
        if (feats) feats[k * 32 * W + t] = sB[t];
      }
      __syncwarp();
    }
  }
}

template <int DB, int DC, bool NEAR_B, bool NEAR_C>
__global__ void __launch_bounds__(128) fast_bwd_kernel(const FastParams P, const FastGrads G, const float* __restrict__ x, int64_t n,
                                                       const int32_t* __restrict__ n_dev, const float* __restrict__ g_feats,
                                                       const float* __restrict__ g_coeff) {
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
    for (int k = 0; k < P.xdim; ++k) xr[k] = x[i * P.xdim + k];
    TapSet<DC, NEAR_C> tc;
    coeff_taps<DC, NEAR_C>(P, xr, tc);
    const float* gf = g_feats ? g_feats + i * P.W : nullptr;
    const float* gcf = g_coeff ? g_coeff + i * P.W : nullptr;
    for (int l = 0; l < P.n_levels; ++l) {
      const FastLevel L = P.lv[l];
      TapSet<DB, NEAR_B> tb;
      basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
      for (int c0 = 0; c0 < L.C; c0 += 2) {
        const int o = L.col + c0;
        float b[2], ca[2];
        gather_vec<DB, NEAR_B, 2>(L.data, L.C, c0, tb, b);
        gather_vec<DC, NEAR_C, 2>(P.cdata, P.W, o, tc, ca);
        float2 g = gf ? *reinterpret_cast<const float2*>(gf + o) : make_float2(0.f, 0.f);
        float gc[2] = {g.x * b[0], g.y * b[1]};
        if (gcf) {
          const float2 g2 = *reinterpret_cast<const float2*>(gcf + o);
          gc[0] += g2.x;
          gc[1] += g2.y;
        }
        const float gb[2] = {g.x * ca[0], g.y * ca[1]};
        if (G.c) scatter_vec<DC, NEAR_C, 2>(G.c, P.W, o, tc, gc);
        if (G.b[l]) scatter_vec<DB, NEAR_B, 2>(G.b[l], L.C, c0, tb, gb);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Gather-free backward: the forward pass saved the coefficient row and the basis row of every query, so the
// backward pass only recomputes the tap indices / weights (ALU) and scatters.  Per query this removes the 120
// scattered loads of the re-gathering kernel; what remains is the reduction traffic, issued as 16-byte
// red.global.add.v4.f32 wherever the target is 16-byte aligned (basis texels with 4 channels; the coefficient
// texel is W floats = 8-byte aligned, so even/odd texels use {v4.., v2} / {v2, v4..} splits).
// AGGW > 0: the coefficient gradient row (W <= AGGW) is accumulated in registers over the levels and scattered
// once per corner with that split; AGGW == 0: wide rows (image presets) scatter level by level.
// ---------------------------------------------------------------------------------------------------------
template <int D, bool NEAREST, int wmax>
__device__ __forceinline__ void scatter_row(float* __restrict__ grad, int W, const TapSet<D, NEAREST>& t, const float* g /*[wmax]*/) {
  constexpr int ROWS = NEAREST ? 1 : (1 << (D - 1));
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (!t.row_ok[r]) continue;
#pragma unroll
    for (int xs = 0; xs < (NEAREST ? 1 : 2); ++xs) {
      const float wx = xs ? t.wx1 : t.wx0;
      if (xs ? !t.x1_ok : (!NEAREST && wx == 0.0f)) continue;
      const float w = t.wrow[r] * wx;
      const size_t e0 = (size_t)(t.base[r] + xs) * W;      // first float of the texel
      float* p = grad + e0;
      if ((e0 & 3) == 0) {                                 // 16-byte aligned texel: v4 v4 ... [v2]
#pragma unroll
        for (int c = 0; c < wmax; c += 4) {
          if (c + 4 <= wmax && c + 4 <= W) red_add_v4(p + c, g[c] * w, g[c + 1] * w, g[c + 2] * w, g[c + 3] * w);
          else if (c + 2 <= wmax && c + 2 <= W) red_add_v2(p + c, g[c] * w, g[c + 1] * w);
        }
      } else {                                             // 8-byte aligned only: v2 v4 v4 ... [v2]
        red_add_v2(p, g[0] * w, g[1] * w);
#pragma unroll
        for (int c = 2; c < wmax; c += 4) {
          if (c + 4 <= wmax && c + 4 <= W) red_add_v4(p + c, g[c] * w, g[c + 1] * w, g[c + 2] * w, g[c + 3] * w);
          else if (c + 2 <= wmax && c + 2 <= W) red_add_v2(p + c, g[c] * w, g[c + 1] * w);
        }
      }
    }
  }
}

template <int DB, int DC, bool NEAR_B, bool NEAR_C, int AGGW, int NT, int MINB, bool LPAR = false>
__global__ void __launch_bounds__(NT, MINB) fast_bwd_saved_kernel(const FastParams P, const FastGrads G, const float* __restrict__ x,
                                                                  int64_t n, const int32_t* __restrict__ n_dev,
                                                                  const float* __restrict__ g_feats, const float* __restrict__ g_coeff,
                                                                  const float* __restrict__ coeff, const float* __restrict__ basis) {
  static_assert(!(LPAR && AGGW > 0), "the level-parallel variant scatters the coefficient gradient level by level");
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  const int64_t n_items = LPAR ? n * P.n_levels : n;
  for (int64_t item = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; item < n_items; item += (int64_t)gridDim.x * blockDim.x) {
    const int64_t i = LPAR ? item / P.n_levels : item;
    const int l_begin = LPAR ? (int)(item % P.n_levels) : 0, l_end = LPAR ? l_begin + 1 : P.n_levels;
    float xr[3];
    for (int k = 0; k < P.xdim; ++k) xr[k] = x[i * P.xdim + k];
    const float* gf = g_feats ? g_feats + i * P.W : nullptr;
    const float* gcf = g_coeff ? g_coeff + i * P.W : nullptr;
    const float* crow = coeff + i * P.W;
    TapSet<DC, NEAR_C> tc;
    if (G.c) coeff_taps<DC, NEAR_C>(P, xr, tc);
    if (AGGW > 0 && G.c) {
      // d/d coeff row = g_feats * basis (+ g_coeff): a flat pass over the W columns, no level structure needed
      float gacc[AGGW > 0 ? AGGW : 2];
#pragma unroll
      for (int c = 0; c < AGGW; c += 2) {
        gacc[c] = gacc[c + 1] = 0.0f;
        if (c < P.W) {
          const float2 g = gf ? *reinterpret_cast<const float2*>(gf + c) : make_float2(0.f, 0.f);
          const float2 b = make_float2(basis[blk_idx(i, c, P.W)], basis[blk_idx(i, c + 1, P.W)]);
          gacc[c] = g.x * b.x;
          gacc[c + 1] = g.y * b.y;
          if (gcf) {
            const float2 g2 = *reinterpret_cast<const float2*>(gcf + c);
            gacc[c] += g2.x;
            gacc[c + 1] += g2.y;
          }
        }
      }
      scatter_row<DC, NEAR_C, (AGGW > 0 ? AGGW : 2)>(G.c, P.W, tc, gacc);
    }
#pragma unroll 1
    for (int l = l_begin; l < l_end; ++l) {
      const FastLevel L = P.lv[l];
      if (!G.b[l] && (AGGW > 0 || !G.c)) continue;
      TapSet<DB, NEAR_B> tb;
      if (G.b[l]) basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
      for (int c0 = 0; c0 < L.C; c0 += 2) {
        const int o = L.col + c0;
        const float2 g = gf ? *reinterpret_cast<const float2*>(gf + o) : make_float2(0.f, 0.f);
        const float2 ca = *reinterpret_cast<const float2*>(crow + o);
        if (AGGW == 0 && G.c) {
          const float2 b = make_float2(basis[blk_idx(i, o, P.W)], basis[blk_idx(i, o + 1, P.W)]);
          float gc[2] = {g.x * b.x, g.y * b.y};
          if (gcf) {
            const float2 g2 = *reinterpret_cast<const float2*>(gcf + o);
            gc[0] += g2.x;
            gc[1] += g2.y;
          }
          scatter_vec<DC, NEAR_C, 2>(G.c, P.W, o, tc, gc);
        }
        if (G.b[l]) {
          float gb[4];
          gb[0] = g.x * ca.x;
          gb[1] = g.y * ca.y;
          if ((L.C & 3) == 0) {          // 16-byte texels: two channel pairs -> one v4 reduction per corner
            if ((c0 & 2) == 0) {
              const float2 g_n = gf ? *reinterpret_cast<const float2*>(gf + o + 2) : make_float2(0.f, 0.f);
              const float2 ca_n = *reinterpret_cast<const float2*>(crow + o + 2);
              gb[2] = g_n.x * ca_n.x;
              gb[3] = g_n.y * ca_n.y;
              scatter_vec<DB, NEAR_B, 4>(G.b[l], L.C, c0, tb, gb);
            }
          } else {
            scatter_vec<DB, NEAR_B, 2>(G.b[l], L.C, c0, tb, gb);
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Run-aggregated coefficient scatter.  Consecutive lanes are consecutive samples of a ray, and a ray takes ~8 steps
// through one cell of the (coarse) coefficient grid, so most lanes of a warp add into the same 2^DC texels.  Each lane
// parks its coefficient-gradient row, its corner weights and its cell in shared memory; the warp then splits the
// (run of equal cells) x (corner) x (16-byte vector) items among its lanes: an item sums its run's contributions and
// issues ONE vector reduction — ~6x fewer L2 reductions for the coefficient grid (40 -> ~7 per query at nerf.yaml).
// The basis levels (fine grids, short runs) scatter per lane as in fast_bwd_saved_kernel.
// ---------------------------------------------------------------------------------------------------------
constexpr int AGG_G = 24;                 // gradient row (W <= 24)
constexpr int AGG_STRIDE = AGG_G + 8 + 4 + 1;   // + corner weights + row bases + flags; odd: conflict-free column reads across rows

template <int DB, int DC, bool NEAR_B, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) fast_bwd_saved_agg_kernel(const FastParams P, const FastGrads G, const float* __restrict__ x,
                                                                      int64_t n, const int32_t* __restrict__ n_dev,
                                                                      const float* __restrict__ g_feats, const float* __restrict__ g_coeff,
                                                                      const float* __restrict__ coeff, const float* __restrict__ basis,
                                                                      int agg_levels) {
  constexpr int ROWS = 1 << (DC - 1), NCORN = 2 * ROWS;
  __shared__ float rec_all[NT / 32][32 * AGG_STRIDE];
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  const int lane = threadIdx.x & 31, W = P.W;
  float* rec = rec_all[threadIdx.x >> 5];
  const int nops = (W + 3) >> 2;            // vector reductions per texel (16-byte ones + at most one 8-byte one)
  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_chunks; k += nwarps) {
    const int64_t i = k * 32 + lane;
    const bool active = i < n;
    float xr[3] = {0.f, 0.f, 0.f};
    const float* gf = nullptr;
    const float* crow = nullptr;
    TapSet<DC, false> tc;
    float* my = rec + lane * AGG_STRIDE;
    int key = -1 - lane;
    if (active) {
      for (int d = 0; d < P.xdim; ++d) xr[d] = x[i * P.xdim + d];
      gf = g_feats ? g_feats + i * W : nullptr;
      crow = coeff + i * W;
    }
    if (G.c) {
      int flags = 0;
      if (active) {
        const float* gcf = g_coeff ? g_coeff + i * W : nullptr;
        coeff_taps<DC, false>(P, xr, tc);
#pragma unroll
        for (int c = 0; c < AGG_G; c += 2) {
          float2 v = make_float2(0.f, 0.f);
          if (c < W) {
            const float2 g = gf ? *reinterpret_cast<const float2*>(gf + c) : make_float2(0.f, 0.f);
            const float2 b = make_float2(basis[blk_idx(i, c, W)], basis[blk_idx(i, c + 1, W)]);
            v = make_float2(g.x * b.x, g.y * b.y);
            if (gcf) {
              const float2 g2 = *reinterpret_cast<const float2*>(gcf + c);
              v.x += g2.x;
              v.y += g2.y;
            }
          }
          my[c] = v.x;
          my[c + 1] = v.y;
        }
#pragma unroll
        for (int r = 0; r < ROWS; ++r) {
          my[AGG_G + 2 * r] = tc.wrow[r] * tc.wx0;
          my[AGG_G + 2 * r + 1] = tc.wrow[r] * tc.wx1;
          my[AGG_G + 8 + r] = __int_as_float(tc.base[r]);
          flags |= (tc.row_ok[r] ? 1 : 0) << r;
        }
        flags |= (tc.x1_ok ? 1 : 0) << 4;
        key = tc.base[0];
      }
      my[AGG_G + 12] = __int_as_float(flags);
      const int prev = __shfl_up_sync(0xffffffffu, key, 1);
      const unsigned hm = __ballot_sync(0xffffffffu, lane == 0 || key != prev);     // run heads
      const unsigned am = __ballot_sync(0xffffffffu, active);
      __syncwarp();
      const int total = __popc(hm) * NCORN * nops;
      for (int t = lane; t < total; t += 32) {
        const int v = t % nops, c = (t / nops) % NCORN, rho = t / (nops * NCORN);
        const int s = (int)__fns(hm, 0, rho + 1);
        if (!((am >> s) & 1u)) continue;
        const unsigned rest = s < 31 ? (hm >> (s + 1)) : 0u;
        const int e = rest ? s + __ffs(rest) : 32;                                    // one past the run's last lane
        const float* L = rec + s * AGG_STRIDE;
        const int fl = __float_as_int(L[AGG_G + 12]), r = c >> 1, xs = c & 1;
        if (!((fl >> r) & 1) || (xs && !((fl >> 4) & 1))) continue;                   // corner outside the grid
        const size_t e0 = (size_t)(__float_as_int(L[AGG_G + 8 + r]) + xs) * W;         // first float of the texel
        int ch0, wd;
        if ((e0 & 3) == 0) { ch0 = 4 * v; wd = (W - ch0) >= 4 ? 4 : 2; }              // 16-byte aligned texel: v4 ... [v2]
        else { ch0 = v == 0 ? 0 : 4 * v - 2; wd = v == 0 ? 2 : 4; }                   // 8-byte aligned only: v2 v4 ...
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int m = s; m < e; ++m) {
          const float* M = rec + m * AGG_STRIDE;
          const float w = M[AGG_G + c];
          a0 += w * M[ch0];
          a1 += w * M[ch0 + 1];
          if (wd == 4) {
            a2 += w * M[ch0 + 2];
            a3 += w * M[ch0 + 3];
          }
        }
        if (wd == 4) red_add_v4(G.c + e0 + ch0, a0, a1, a2, a3);
        else red_add_v2(G.c + e0 + ch0, a0, a1);
      }
      __syncwarp();
    }
    // ---- the coarsest basis levels (4-channel texels: one 16-byte reduction per corner) the same way
    constexpr int BROWS = NEAR_B ? 1 : (1 << (DB - 1)), BCORN = NEAR_B ? 1 : 2 * BROWS;
    int l_first = 0;
    if (!NEAR_B) {
#pragma unroll 1
      for (; l_first < P.n_levels && l_first < agg_levels; ++l_first) {
        const FastLevel L = P.lv[l_first];
        if (L.C != 4) break;
        if (!G.b[l_first]) continue;
        int bkey = -1 - lane, flags = 0;
        if (active) {
          TapSet<DB, NEAR_B> tb;
          basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
          // rows are W floats apart (8-byte aligned only): float2 loads
          const float2 ga = gf ? *reinterpret_cast<const float2*>(gf + L.col) : make_float2(0.f, 0.f);
          const float2 gb2 = gf ? *reinterpret_cast<const float2*>(gf + L.col + 2) : make_float2(0.f, 0.f);
          const float2 ca = *reinterpret_cast<const float2*>(crow + L.col), cb = *reinterpret_cast<const float2*>(crow + L.col + 2);
          my[0] = ga.x * ca.x; my[1] = ga.y * ca.y; my[2] = gb2.x * cb.x; my[3] = gb2.y * cb.y;
#pragma unroll
          for (int r = 0; r < BROWS; ++r) {
            my[AGG_G + 2 * r] = tb.wrow[r] * tb.wx0;
            my[AGG_G + 2 * r + 1] = tb.wrow[r] * tb.wx1;
            my[AGG_G + 8 + r] = __int_as_float(tb.base[r]);
            flags |= (tb.row_ok[r] ? 1 : 0) << r;
          }
          flags |= (tb.x1_ok ? 1 : 0) << 4;
          flags |= (tb.x0_ok ? 1 : 0) << 5;
          bkey = tb.base[0];
        }
        my[AGG_G + 12] = __int_as_float(flags);
        const int prev = __shfl_up_sync(0xffffffffu, bkey, 1);
        const unsigned hm = __ballot_sync(0xffffffffu, lane == 0 || bkey != prev);
        const unsigned am = __ballot_sync(0xffffffffu, active);
        __syncwarp();
        const int total = __popc(hm) * BCORN;
        for (int t = lane; t < total; t += 32) {
          const int c = t % BCORN, rho = t / BCORN;
          const int s = (int)__fns(hm, 0, rho + 1);
          if (!((am >> s) & 1u)) continue;
          const unsigned rest = s < 31 ? (hm >> (s + 1)) : 0u;
          const int e = rest ? s + __ffs(rest) : 32;
          const float* Ld = rec + s * AGG_STRIDE;
          const int fl = __float_as_int(Ld[AGG_G + 12]), r = c >> 1, xs = c & 1;
          if (!((fl >> r) & 1) || !((fl >> (xs ? 4 : 5)) & 1)) continue;               // corner outside the grid (zeros padding)
          float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
          for (int m = s; m < e; ++m) {
            const float* M = rec + m * AGG_STRIDE;
            const float w = M[AGG_G + c];
            a0 += w * M[0]; a1 += w * M[1]; a2 += w * M[2]; a3 += w * M[3];
          }
          red_add_v4(G.b[l_first] + (size_t)(__float_as_int(Ld[AGG_G + 8 + r]) + xs) * 4, a0, a1, a2, a3);
        }
        __syncwarp();
      }
    }
    if (!active) continue;
#pragma unroll 1
    for (int l = l_first; l < P.n_levels; ++l) {
      const FastLevel L = P.lv[l];
      if (!G.b[l]) continue;
      TapSet<DB, NEAR_B> tb;
      basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
      for (int c0 = 0; c0 < L.C; c0 += 2) {
        const int o = L.col + c0;
        const float2 g = gf ? *reinterpret_cast<const float2*>(gf + o) : make_float2(0.f, 0.f);
        const float2 ca = *reinterpret_cast<const float2*>(crow + o);
        float gb[4];
        gb[0] = g.x * ca.x;
        gb[1] = g.y * ca.y;
        if ((L.C & 3) == 0) {
          if ((c0 & 2) == 0) {
            const float2 g_n = gf ? *reinterpret_cast<const float2*>(gf + o + 2) : make_float2(0.f, 0.f);
            const float2 ca_n = *reinterpret_cast<const float2*>(crow + o + 2);
            gb[2] = g_n.x * ca_n.x;
            gb[3] = g_n.y * ca_n.y;
            scatter_vec<DB, NEAR_B, 4>(G.b[l], L.C, c0, tb, gb);
          }
        } else {
          scatter_vec<DB, NEAR_B, 2>(G.b[l], L.C, c0, tb, gb);
        }
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Wide rows (image.yaml / image_set.yaml: W = 144 channels, levels of 32 / 16): one thread per (query, 16-byte column group).
// Consecutive lanes are consecutive column groups of one query, so every access of a warp is coalesced: the texel of a query
// (W floats of the coefficient grid, C_l floats of a basis level) is read as contiguous float4 pieces by neighbouring lanes and
// the feats / coeff / basis rows leave as contiguous 512-byte pieces — one thread per query (or per (query, level)) walks its
// 576-byte row alone, 32 lanes 576 bytes apart (image.yaml forward 157 us for 102 400 points = 0.17 of the HBM model).  The
// coordinate arithmetic is repeated by the W / 4 lanes of a query (ALU only; same ffb_math.h sequence, identical tap indices).
// The saved basis row of this kernel pair is ROW-MAJOR [n, W] (the narrow kernels store theirs blocked by 32 queries).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int wide_level(const FastParams& P, int c) {
  int l = 0;
  while (l + 1 < P.n_levels && P.lv[l + 1].col <= c) ++l;
  return l;
}

// V = 16-byte pieces per thread (2 when every level's channel count and offset is a multiple of 8: half the repeated
// coordinate arithmetic per byte moved; a thread's two pieces are adjacent, so a warp still covers contiguous memory)
template <int DB, int DC, bool NEAR_B, bool NEAR_C, int V>
__global__ void __launch_bounds__(256) wide_fwd_kernel(const FastParams P, const float* __restrict__ x, int64_t n,
                                                       const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                       float* __restrict__ coeff, float* __restrict__ basis) {
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  const unsigned WG = (unsigned)P.W / (4 * V);
  const unsigned n_items = (unsigned)n * WG;                       // < 2^31 (checked by the launcher)
  for (unsigned item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
    const unsigned i = item / WG;
    const int c = (int)(item - i * WG) * 4 * V;
    const int l = wide_level(P, c);
    const FastLevel& L = P.lv[l];
    float xr[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < DC; ++d) xr[d] = x[(size_t)i * DC + d];
    TapSet<DC, NEAR_C> tc;
    coeff_taps<DC, NEAR_C>(P, xr, tc);
    TapSet<DB, NEAR_B> tb;
    basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
    float b[V][4], ca[V][4];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      gather_vec<DB, NEAR_B, 4>(L.data, L.C, c + 4 * v - L.col, tb, b[v]);
      gather_vec<DC, NEAR_C, 4>(P.cdata, P.W, c + 4 * v, tc, ca[v]);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const size_t o = (size_t)i * P.W + c + 4 * v;
      if (feats) *reinterpret_cast<float4*>(feats + o) = make_float4(b[v][0] * ca[v][0], b[v][1] * ca[v][1], b[v][2] * ca[v][2], b[v][3] * ca[v][3]);
      if (coeff) *reinterpret_cast<float4*>(coeff + o) = make_float4(ca[v][0], ca[v][1], ca[v][2], ca[v][3]);
      if (basis) *reinterpret_cast<float4*>(basis + o) = make_float4(b[v][0], b[v][1], b[v][2], b[v][3]);
    }
  }
}

// coeff / basis (both or neither): the rows saved by wide_fwd_kernel; otherwise the texels are re-gathered (coalesced as well)
template <int DB, int DC, bool NEAR_B, bool NEAR_C, int V>
__global__ void __launch_bounds__(256) wide_bwd_kernel(const FastParams P, const FastGrads G, const float* __restrict__ x, int64_t n,
                                                       const int32_t* __restrict__ n_dev, const float* __restrict__ g_feats,
                                                       const float* __restrict__ g_coeff, const float* __restrict__ coeff,
                                                       const float* __restrict__ basis) {
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  const unsigned WG = (unsigned)P.W / (4 * V);
  const unsigned n_items = (unsigned)n * WG;
  const bool saved = coeff && basis;
  for (unsigned item = blockIdx.x * blockDim.x + threadIdx.x; item < n_items; item += gridDim.x * blockDim.x) {
    const unsigned i = item / WG;
    const int c = (int)(item - i * WG) * 4 * V;
    const int l = wide_level(P, c);
    const FastLevel& L = P.lv[l];
    float xr[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int d = 0; d < DC; ++d) xr[d] = x[(size_t)i * DC + d];
    TapSet<DC, NEAR_C> tc;
    coeff_taps<DC, NEAR_C>(P, xr, tc);
    TapSet<DB, NEAR_B> tb;
    basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
    float* gbl = nullptr;                       // G.b[l] without indexing the parameter struct dynamically (local-memory copy)
#pragma unroll
    for (int k = 0; k < FAST_MAX_LEVELS; ++k)
      if (k == l) gbl = G.b[k];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const int cv = c + 4 * v;
      const size_t o = (size_t)i * P.W + cv;
      float b[4], ca[4];
      if (saved) {
        const float4 bv = *reinterpret_cast<const float4*>(basis + o), cf = *reinterpret_cast<const float4*>(coeff + o);
        b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
        ca[0] = cf.x; ca[1] = cf.y; ca[2] = cf.z; ca[3] = cf.w;
      } else {
        gather_vec<DB, NEAR_B, 4>(L.data, L.C, cv - L.col, tb, b);
        gather_vec<DC, NEAR_C, 4>(P.cdata, P.W, cv, tc, ca);
      }
      const float4 g = g_feats ? *reinterpret_cast<const float4*>(g_feats + o) : make_float4(0.f, 0.f, 0.f, 0.f);
      float gc[4] = {g.x * b[0], g.y * b[1], g.z * b[2], g.w * b[3]};
      if (g_coeff) {
        const float4 g2 = *reinterpret_cast<const float4*>(g_coeff + o);
        gc[0] += g2.x; gc[1] += g2.y; gc[2] += g2.z; gc[3] += g2.w;
      }
      const float gb[4] = {g.x * ca[0], g.y * ca[1], g.z * ca[2], g.w * ca[3]};
      if (G.c) scatter_vec<DC, NEAR_C, 4>(G.c, P.W, cv, tc, gc);
      if (gbl) scatter_vec<DB, NEAR_B, 4>(gbl, L.C, cv - L.col, tb, gb);
    }
  }
}

}  // namespace ffb

using namespace ffb;

// ---- launch configuration (tunable at run time for experiments; defaults are the measured best) -------------
static int g_fwd_cfg = 1;   // 0: 128 threads, compiler-chosen registers   1: 128 x >=8 CTAs/SM   2: 128 x >=6   3: 256 x >=4
static int g_fwd_crow = 1;  // 1: W == 18 rows take the coefficient row by aligned 16-byte windows first (knob "field_fwd_crow")
static int g_fwd_stage = 1; // 1: narrow rows (W <= 32) leave through shared memory as coalesced 16-byte pieces (278 -> 273 us at nerf.yaml)
static int g_bwd_cfg = 2;   // 0: re-gathering kernel   1: saved-activation kernel (when coeff/basis rows are supplied)   2: 1 + run-aggregated coefficient scatter
static int g_agg_levels = 0;   // leading 4-channel basis levels whose scatter is run-aggregated too (knob "field_bwd_agg_levels"): measured
                               // SLOWER at nerf.yaml (2 levels: 403 vs 342 us) — their runs are 2-5 samples, the extra staging pass costs more
static int g_fwd_lpar_all = 0;  // experiment knob "field_fwd_lpar_all": level-parallel forward at every batch size
static int g_lpar = 1;      // 1: batches of at most LPAR_MAX_ITEMS (query, level) pairs use the level-parallel kernels
constexpr int64_t LPAR_MAX_ITEMS = 148 * 2048 * 3;   // ~3 full waves of resident threads; above that one thread per query wins
static int g_wide = 1;      // knob "field_wide": rows of >= 64 channels in 16-byte texel pieces take the column-parallel kernels

// wide rows: every level's channels and offset a multiple of 4 floats (16-byte pieces never straddle a level or a texel)
// (a function of the field and the batch size only: the forward and the backward launch must agree on the saved-row layout)
static bool wide_eligible(const FastParams& P, int64_t n) {
  if (!g_wide || P.W < 64 || (P.W & 3) || n * (P.W >> 2) >= ((int64_t)1 << 31)) return false;
  for (int l = 0; l < P.n_levels; ++l)
    if ((P.lv[l].C & 3) || (P.lv[l].col & 3)) return false;
  return true;
}
static int g_wide_pairs = 1;    // knob "field_wide_pairs": two adjacent 16-byte pieces per thread where the level widths allow
static bool wide_pairs(const FastParams& P) {
  if (!g_wide_pairs || (P.W & 7)) return false;
  for (int l = 0; l < P.n_levels; ++l)
    if ((P.lv[l].C & 7) || (P.lv[l].col & 7)) return false;
  return true;
}
static bool aligned16(const void* a, const void* b, const void* c, const void* d) {
  return (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c | (uintptr_t)d) & 15) == 0;
}

template <int DB, int DC, bool NB, bool NC, int NT, int MINB>
static void launch_fwd_cfg(const FastParams& P, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis,
                           cudaStream_t s) {
  if (wide_eligible(P, n)) {
    if (NB && NC && wide_pairs(P))        // nearest taps only (measured: image.yaml 62 + 72 -> 58 + 57 us; with linear taps — image_set — 82 + 100 -> 86 + 133 us)
      wide_fwd_kernel<DB, DC, NB, NC, 2><<<blocks_for(n * (P.W >> 3), 256, (int64_t)sm_count() * 32), 256, 0, s>>>(P, x, n, n_dev, feats, coeff, basis);
    else
      wide_fwd_kernel<DB, DC, NB, NC, 1><<<blocks_for(n * (P.W >> 2), 256, (int64_t)sm_count() * 32), 256, 0, s>>>(P, x, n, n_dev, feats, coeff, basis);
    return;
  }
  if (n * P.n_levels <= (g_fwd_lpar_all ? (int64_t)1 << 40 : LPAR_MAX_ITEMS) && g_lpar) {      // small batch: one thread per (query, level)
    fast_fwd_kernel<DB, DC, NB, NC, NT, MINB, false, true><<<blocks_for(n * P.n_levels, NT, (int64_t)sm_count() * 64), NT, 0, s>>>(
        P, x, n, n_dev, feats, coeff, basis);
    return;
  }
  const unsigned grid = blocks_for(n, NT, (int64_t)sm_count() * 64);
  if (P.W <= 32 && g_fwd_stage) {
    const size_t smem = (size_t)(NT / 32) * 64 * P.W * sizeof(float);
    const int64_t texels = (int64_t)P.csize[0] * P.csize[1] * P.csize[2];
    if (!NC && P.W == 18 && g_fwd_crow && (texels & 1) == 0) {
      fast_fwd_kernel<DB, DC, NB, false, NT, MINB, true, false, 18><<<grid, NT, smem, s>>>(P, x, n, n_dev, feats, coeff, basis);
      return;
    }
    fast_fwd_kernel<DB, DC, NB, NC, NT, MINB, true><<<grid, NT, smem, s>>>(P, x, n, n_dev, feats, coeff, basis);
  } else {
    fast_fwd_kernel<DB, DC, NB, NC, NT, MINB, false><<<grid, NT, 0, s>>>(P, x, n, n_dev, feats, coeff, basis);
  }
}

template <int DB, int DC, bool NB, bool NC>
static void launch_fwd(const FastParams& P, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis,
                       cudaStream_t s) {
  switch (g_fwd_cfg) {
    case 0: launch_fwd_cfg<DB, DC, NB, NC, 128, 1>(P, x, n, n_dev, feats, coeff, basis, s); break;
    case 2: launch_fwd_cfg<DB, DC, NB, NC, 128, 6>(P, x, n, n_dev, feats, coeff, basis, s); break;
    case 3: launch_fwd_cfg<DB, DC, NB, NC, 256, 4>(P, x, n, n_dev, feats, coeff, basis, s); break;
    default: launch_fwd_cfg<DB, DC, NB, NC, 128, 8>(P, x, n, n_dev, feats, coeff, basis, s); break;
  }
}

template <int DB, int DC, bool NB, bool NC>
static void launch_bwd(const FastParams& P, const FastGrads& G, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                       const float* g_coeff, const float* coeff, const float* basis, cudaStream_t s) {
  const int64_t cap = (int64_t)sm_count() * 64;
  if (wide_eligible(P, n)) {
    const bool saved = coeff && basis;
    if (NB && NC && wide_pairs(P))        // nearest taps only (measured: image.yaml 62 + 72 -> 58 + 57 us; with linear taps — image_set — 82 + 100 -> 86 + 133 us)
      wide_bwd_kernel<DB, DC, NB, NC, 2><<<blocks_for(n * (P.W >> 3), 256, (int64_t)sm_count() * 32), 256, 0, s>>>(
          P, G, x, n, n_dev, g_feats, g_coeff, saved ? coeff : nullptr, saved ? basis : nullptr);
    else
      wide_bwd_kernel<DB, DC, NB, NC, 1><<<blocks_for(n * (P.W >> 2), 256, (int64_t)sm_count() * 32), 256, 0, s>>>(
          P, G, x, n, n_dev, g_feats, g_coeff, saved ? coeff : nullptr, saved ? basis : nullptr);
    return;
  }
  if (coeff && basis && g_bwd_cfg != 0 && n * P.n_levels <= LPAR_MAX_ITEMS && g_lpar) {
    fast_bwd_saved_kernel<DB, DC, NB, NC, 0, 128, 6, true><<<blocks_for(n * P.n_levels, 128, cap), 128, 0, s>>>(P, G, x, n, n_dev, g_feats, g_coeff,
                                                                                                            coeff, basis);
    return;
  }
  if (coeff && basis && g_bwd_cfg == 2 && P.W <= AGG_G && !NC && (P.W & 1) == 0) {      // run-aggregated coefficient scatter
    fast_bwd_saved_agg_kernel<DB, DC, NB, 128, 6><<<blocks_for(n, 128, cap), 128, 0, s>>>(P, G, x, n, n_dev, g_feats, g_coeff, coeff, basis, g_agg_levels);
    return;
  }
  if (coeff && basis && g_bwd_cfg != 0) {
    if (P.W <= 24)
      fast_bwd_saved_kernel<DB, DC, NB, NC, 24, 128, 6><<<blocks_for(n, 128, cap), 128, 0, s>>>(P, G, x, n, n_dev, g_feats, g_coeff, coeff, basis);
    else
      fast_bwd_saved_kernel<DB, DC, NB, NC, 0, 128, 6><<<blocks_for(n, 128, cap), 128, 0, s>>>(P, G, x, n, n_dev, g_feats, g_coeff, coeff, basis);
  } else {
    fast_bwd_kernel<DB, DC, NB, NC><<<blocks_for(n, 128, cap), 128, 0, s>>>(P, G, x, n, n_dev, g_feats, g_coeff);
  }
}

#define FAST_DISPATCH(FN, ...)                                                                                  \
  do {                                                                                                          \
    const bool nb = f->h.ops[f->h.bterms[0].op[0]].nearest, nc = f->h.ops[f->h.cterms[0].op[0]].nearest;        \
    const int db = P.in_dim, dc = P.xdim;                                                                       \
    if (db == 3 && dc == 3 && !nb && !nc) FN<3, 3, false, false>(__VA_ARGS__);                                  \
    else if (db == 3 && dc == 3 && nb && nc) FN<3, 3, true, true>(__VA_ARGS__);                                 \
    else if (db == 2 && dc == 2 && !nb && !nc) FN<2, 2, false, false>(__VA_ARGS__);                             \
    else if (db == 2 && dc == 2 && nb && nc) FN<2, 2, true, true>(__VA_ARGS__);                                 \
    else if (db == 2 && dc == 3 && !nb && !nc) FN<2, 3, false, false>(__VA_ARGS__);                             \
    else { set_error("fast path: unsupported dim/mode combination"); return FFB_EINVAL; }                       \
  } while (0)

static bool mode_supported(ffb_field_t f, const FastParams& P) {
  const bool nb = f->h.ops[f->h.bterms[0].op[0]].nearest, nc = f->h.ops[f->h.cterms[0].op[0]].nearest;
  const int db = P.in_dim, dc = P.xdim;
  if (db == 3 && dc == 3) return nb == nc;
  if (db == 2 && dc == 2) return nb == nc;
  if (db == 2 && dc == 3) return !nb && !nc;
  return false;
}

extern "C" int ffb_set_field_planes_tuning(int which, int value);
extern "C" int ffb_set_field_lines_walk(int value);

extern "C" {

int ffb_set_tuning(const char* key, int value) {
  FFB_REQUIRE(key, "null key");
  if (!strcmp(key, "field_fwd_cfg")) g_fwd_cfg = value;
  else if (!strcmp(key, "field_bwd_cfg")) g_bwd_cfg = value;
  else if (!strcmp(key, "field_fwd_stage")) g_fwd_stage = value;
  else if (!strcmp(key, "field_fwd_crow")) g_fwd_crow = value;
  else if (!strcmp(key, "field_level_parallel")) g_lpar = value;
  else if (!strcmp(key, "field_bwd_agg_levels")) g_agg_levels = value;
  else if (!strcmp(key, "field_fwd_lpar_all")) g_fwd_lpar_all = value;
  else if (!strcmp(key, "field_wide")) g_wide = value;
  else if (!strcmp(key, "field_wide_pairs")) g_wide_pairs = value;
  else if (!strcmp(key, "field_deterministic")) ffb::g_deterministic = value;
  else if (!strcmp(key, "field_lines_walk")) return ffb_set_field_lines_walk(value);
  else if (!strcmp(key, "field_planes_v2")) return ffb_set_field_planes_tuning(0, value);
  else if (!strcmp(key, "field_planes_unroll")) return ffb_set_field_planes_tuning(1, value);
  else { set_error("ffb_set_tuning: unknown key %s", key); return FFB_EINVAL; }
  return FFB_OK;
}

int ffb_field_fast_eligible(ffb_field_t f) {
  if (!f) return 0;
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  return (build_params(f->h, P, idx) && mode_supported(f, P)) ? 1 : 0;
}

int ffb_field_saved_basis_layout(ffb_field_t f, int64_t n) {
  if (!f) return -1;
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  if (!(build_params(f->h, P, idx) && mode_supported(f, P))) return -1;
  return wide_eligible(P, n) ? 1 : 0;
}

int ffb_field_fast_fwd_train(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis,
                             void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  FFB_REQUIRE(build_params(f->h, P, idx) && mode_supported(f, P), "descriptor is not eligible for the fast path");
  if (n <= 0) return FFB_OK;
  FFB_REQUIRE(!wide_eligible(P, n) || aligned16(feats, coeff, basis, nullptr), "feats / coeff / basis rows must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  FAST_DISPATCH(launch_fwd, P, x, n, n_dev, feats, coeff, basis, s);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_fast_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, void* stream) {
  return ffb_field_fast_fwd_train(f, x, n, n_dev, feats, coeff, nullptr, stream);
}

int ffb_field_fast_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                             const float* coeff, const float* basis, float* const* h_grads, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  FFB_REQUIRE(build_params(f->h, P, idx) && mode_supported(f, P), "descriptor is not eligible for the fast path");
  if (n <= 0) return FFB_OK;
  FastGrads G;
  G.c = h_grads ? h_grads[idx[0]] : f->h.ops[idx[0]].grad;
  bool aligned = ((uintptr_t)G.c & 15) == 0;
  for (int l = 0; l < FAST_MAX_LEVELS; ++l) {
    G.b[l] = l < P.n_levels ? (h_grads ? h_grads[idx[l + 1]] : f->h.ops[idx[l + 1]].grad) : nullptr;
    aligned = aligned && ((uintptr_t)G.b[l] & 15) == 0;
  }
  FFB_REQUIRE(aligned, "gradient tensors must be 16-byte aligned");
  FFB_REQUIRE(!wide_eligible(P, n) || aligned16(g_feats, g_coeff, coeff, basis), "gradient / saved rows must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  FAST_DISPATCH(launch_bwd, P, G, x, n, n_dev, g_feats, g_coeff, coeff, basis, s);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_fast_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                       float* const* h_grads, void* stream) {
  return ffb_field_fast_bwd_saved(f, x, n, n_dev, g_feats, g_coeff, nullptr, nullptr, h_grads, stream);
}

}  // extern "C"
