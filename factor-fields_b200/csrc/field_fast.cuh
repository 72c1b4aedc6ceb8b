// field_fast.cuh — device building blocks of the specialised grid x grid field kernels, shared by field_fast.cu (gather /
// scatter kernels) and field_mlp.cu (the same gather fused with linear_mat on the tensor cores): tap sets, vector
// gathers / reductions, the blocked row layout, and the descriptor -> FastParams translation.
#pragma once
#include "ffb_common.cuh"
#include "ffb_math.h"
#include <string.h>

struct ffb_field {
  ffb_field_desc h;
  ffb_field_desc* d;
};

namespace ffb {

constexpr int FAST_MAX_LEVELS = 8;

struct FastLevel {
  const float* data;
  int C, R, col;
  float freq;
};

struct FastParams {
  int xdim, in_dim, mapping, n_levels, W;
  float lo[3], hi[3];
  const float* cdata;
  int csize[3];  // W, H, D of the coefficient grid
  FastLevel lv[FAST_MAX_LEVELS];
};

struct FastGrads {
  float* c;
  float* b[FAST_MAX_LEVELS];
};

template <int D, bool NEAREST>
struct TapSet {
  // linear: 2^(D-1) row bases (offset of the x-low corner, in texels) + per-row weight + x weights
  int base[NEAREST ? 1 : (1 << (D - 1))];
  float wrow[NEAREST ? 1 : (1 << (D - 1))];
  float wx0, wx1;
  bool x0_ok;   // x-low corner inside the grid (its weight is also folded to 0 when not)
  bool x1_ok;   // x-high corner inside the grid
  bool row_ok[NEAREST ? 1 : (1 << (D - 1))];
};

template <int D, bool NEAREST>
__device__ __forceinline__ void make_tapset(const float c[3], const int size[3], TapSet<D, NEAREST>& t) {
  if (NEAREST) {
    int idx = 0, stride = 1;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < D; ++k) {
      int i = nearest_index(c[k]);
      ok = ok && i >= 0 && i < size[k];
      idx += i * stride;
      stride *= size[k];
    }
    t.base[0] = idx;
    t.row_ok[0] = ok;
    t.wrow[0] = 1.0f;
    t.wx0 = 1.0f;
    t.wx1 = 0.0f;
    t.x0_ok = ok;
    t.x1_ok = false;
    return;
  }
  Axis ax[3];
#pragma unroll
  for (int k = 0; k < D; ++k) ax[k] = linear_axis(c[k]);
  t.wx0 = ax[0].w0;
  t.wx1 = ax[0].w1;
  const bool x0_ok = ax[0].i0 >= 0 && ax[0].i0 < size[0];
  t.x1_ok = ax[0].i0 + 1 >= 0 && ax[0].i0 + 1 < size[0];
#pragma unroll
  for (int r = 0; r < (1 << (D - 1)); ++r) {
    int idx = ax[0].i0, stride = size[0];
    bool ok = true;
    float w = 1.0f;
#pragma unroll
    for (int k = 1; k < D; ++k) {
      const int b = (r >> (k - 1)) & 1;
      const int i = ax[k].i0 + b;
      ok = ok && i >= 0 && i < size[k];
      idx += i * stride;
      stride *= size[k];
      const float wk = b ? ax[k].w1 : ax[k].w0;
      w = (k == 1) ? wk : FFB_MUL(w, wk);
    }
    t.base[r] = idx;
    t.row_ok[r] = ok;
    t.wrow[r] = w;
  }
  // fold "x-low corner out of bounds" (zeros padding, only possible when the coordinate is outside [-1,1]) into the weight
  t.x0_ok = x0_ok;
  if (!x0_ok) t.wx0 = 0.0f;
  if (!t.x1_ok) t.wx1 = 0.0f;
  // NB: weights are products (wx*wy)*wz in ATen; we apply wx * (wy*wz) — equal up to one rounding (inside 1e-4 bar).
}

// v[j] = sum over taps of w * texel[c0 + j], j < NV (NV = 2 or 4); texel stride C floats.
template <int D, bool NEAREST, int NV>
__device__ __forceinline__ void gather_vec(const float* __restrict__ data, int C, int c0, const TapSet<D, NEAREST>& t, float v[NV]) {
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = 0.0f;
  constexpr int ROWS = NEAREST ? 1 : (1 << (D - 1));
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (!t.row_ok[r]) continue;
    const float* p = data + (size_t)t.base[r] * C + c0;
    const float w0 = t.wrow[r] * t.wx0;
    if (NV == 4) {
      if (NEAREST || t.wx0 != 0.0f) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(p));
        v[0] += a.x * w0; v[1] += a.y * w0; v[2] += a.z * w0; v[3] += a.w * w0;
      }
      if (!NEAREST && t.x1_ok) {
        const float w1 = t.wrow[r] * t.wx1;
        const float4 b = __ldg(reinterpret_cast<const float4*>(p + C));
        v[0] += b.x * w1; v[1] += b.y * w1; v[2] += b.z * w1; v[3] += b.w * w1;
      }
    } else {
      if (NEAREST || t.wx0 != 0.0f) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(p));
        v[0] += a.x * w0; v[1] += a.y * w0;
      }
      if (!NEAREST && t.x1_ok) {
        const float w1 = t.wrow[r] * t.wx1;
        const float2 b = __ldg(reinterpret_cast<const float2*>(p + C));
        v[0] += b.x * w1; v[1] += b.y * w1;
      }
    }
  }
}

__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int D, bool NEAREST, int NV>
__device__ __forceinline__ void scatter_vec(float* __restrict__ grad, int C, int c0, const TapSet<D, NEAREST>& t, const float g[NV]) {
  constexpr int ROWS = NEAREST ? 1 : (1 << (D - 1));
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (!t.row_ok[r]) continue;
    float* p = grad + (size_t)t.base[r] * C + c0;
    const float w0 = t.wrow[r] * t.wx0;
    if (!NEAREST && NV == 2 && C == 2 && t.x0_ok && t.x1_ok && (t.base[r] & 1) == 0) {
      // 2-channel texels: both x corners in one 16-byte reduction (-5 % on the scatter; the same merge on the forward's
      // gathers made that kernel 15 % SLOWER — a divergent branch in its latency-bound inner loop — and is not used)
      const float w1 = t.wrow[r] * t.wx1;
      red_add_v4(p, g[0] * w0, g[1] * w0, g[0] * w1, g[1] * w1);
      continue;
    }
    if (NEAREST || t.wx0 != 0.0f) {
      if (NV == 4) red_add_v4(p, g[0] * w0, g[1] * w0, g[2] * w0, g[3] * w0);
      else red_add_v2(p, g[0] * w0, g[1] * w0);
    }
    if (!NEAREST && t.x1_ok) {
      const float w1 = t.wrow[r] * t.wx1;
      if (NV == 4) red_add_v4(p + C, g[0] * w1, g[1] * w1, g[2] * w1, g[3] * w1);
      else red_add_v2(p + C, g[0] * w1, g[1] * w1);
    }
  }
}

template <int DC, bool NEAR_C>
__device__ __forceinline__ void coeff_taps(const FastParams& P, const float* xr, TapSet<DC, NEAR_C>& t) {
  float c[3];
#pragma unroll
  for (int k = 0; k < DC; ++k) c[k] = source_index(normalize_coord(xr[k], P.lo[k], P.hi[k]), P.csize[k], 0, 1);
  make_tapset<DC, NEAR_C>(c, P.csize, t);
}

template <int DB, bool NEAR_B>
__device__ __forceinline__ void basis_taps(const FastParams& P, const FastLevel& L, const float* xr, float msize, TapSet<DB, NEAR_B>& t) {
  float c[3];
  const int size[3] = {L.R, L.R, L.R};
  const float scale = FFB_DIV(msize, L.freq);
#pragma unroll
  for (int k = 0; k < DB; ++k) c[k] = source_index(map_coord(xr[k], P.lo[k], scale, P.mapping, nullptr), L.R, 1, 0);
  make_tapset<DB, NEAR_B>(c, size, t);
}

// The saved basis row is private to the forward / backward kernel pair, so it is stored BLOCKED: element (query i, column c)
// at (i / 32) * 32 W + c * 32 + i % 32.  A warp's 32 queries then write / read 128 contiguous bytes per column instead of 32
// pieces 4 W bytes apart (one row per lane costs ~9x the LSU wavefronts; the row alone was 51 of the forward's 300 us).
// Buffers hold ceil(n / 32) * 32 rows.
__device__ __forceinline__ size_t blk_idx(int64_t i, int c, int W) { return (size_t)(i >> 5) * (size_t)(32 * W) + (size_t)c * 32 + (size_t)(i & 31); }

__device__ __forceinline__ float fast_msize(const FastParams& P) {
  float m = FFB_SUB(P.hi[0], P.lo[0]);
  for (int k = 1; k < P.in_dim; ++k) m = fmaxf(m, FFB_SUB(P.hi[k], P.lo[k]));
  return m;
}

inline bool build_params(const ffb_field_desc& d, FastParams& P, int op_index[FAST_MAX_LEVELS + 1]) {
  if (d.coeff_width <= 0 || d.basis_width != d.coeff_width || d.basis_is_x || d.basis_perm) return false;
  if (d.n_cterms != 1 || d.cterms[0].n_ops != 1 || d.cterms[0].col != 0) return false;
  if (d.n_bterms < 1 || d.n_bterms > FAST_MAX_LEVELS) return false;
  if (d.mapping == FFB_MAP_TRIG) return false;
  const ffb_gather_op& c = d.ops[d.cterms[0].op[0]];
  if (c.nd != d.xdim || (c.nd != 2 && c.nd != 3) || c.space != 0 || c.align_corners || !c.border) return false;
  if (c.C != d.coeff_width || (c.C & 1) || ((uintptr_t)c.data & 15)) return false;
  for (int k = 0; k < c.nd; ++k)
    if (c.src[k] != k) return false;
  P.xdim = d.xdim;
  P.in_dim = d.in_dim;
  P.mapping = d.mapping;
  P.n_levels = d.n_bterms;
  P.W = d.coeff_width;
  for (int k = 0; k < 3; ++k) {
    P.lo[k] = d.aabb_min[k];
    P.hi[k] = d.aabb_max[k];
    P.csize[k] = k < c.nd ? c.size[k] : 1;
  }
  P.cdata = c.data;
  op_index[0] = d.cterms[0].op[0];
  int col = 0;
  for (int l = 0; l < d.n_bterms; ++l) {
    const ffb_term& T = d.bterms[l];
    if (T.n_ops != 1 || T.col != col) return false;
    const ffb_gather_op& b = d.ops[T.op[0]];
    if (b.nd != d.in_dim || b.space != 1 || b.level != l || !b.align_corners || b.border) return false;
    if ((b.C & 1) || ((uintptr_t)b.data & 15)) return false;
    if (b.nearest != d.ops[d.bterms[0].op[0]].nearest) return false;
    for (int k = 0; k < b.nd; ++k)
      if (b.src[k] != k || b.size[k] != b.size[0]) return false;
    P.lv[l].data = b.data;
    P.lv[l].C = b.C;
    P.lv[l].R = b.size[0];
    P.lv[l].col = col;
    P.lv[l].freq = d.freq[l];
    op_index[l + 1] = T.op[0];
    col += b.C;
  }
  return col == d.coeff_width && (d.in_dim == 2 || d.in_dim == 3);
}

}  // namespace ffb
