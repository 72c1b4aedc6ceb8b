// field_fast.cu — specialised field-query kernels for the grid-coefficient x grid-basis fields
// (nerf.yaml / sdf.yaml / image.yaml / image_set.yaml shapes): the headline HBM-roofline kernels K1 / K2.
//
// One thread per query, consecutive threads = consecutive samples of a ray (the compacted order), so that
// neighbouring lanes hit the same or neighbouring texels.  Channels-last texels are read as float2/float4
// vectors; the x-neighbour corners of a linear tap are adjacent in memory and fetched together.  The
// backward pass re-gathers (no saved activations) and scatters with vector reductions
// (red.global.add.v2/v4.f32).  Coordinate arithmetic is the shared unfused fp32 sequence of ffb_math.h, so
// tap indices are identical to the generic path and to the reference.
#include "ffb_common.cuh"
#include "ffb_math.h"

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

__device__ __forceinline__ float fast_msize(const FastParams& P) {
  float m = FFB_SUB(P.hi[0], P.lo[0]);
  for (int k = 1; k < P.in_dim; ++k) m = fmaxf(m, FFB_SUB(P.hi[k], P.lo[k]));
  return m;
}

template <int DB, int DC, bool NEAR_B, bool NEAR_C>
__global__ void __launch_bounds__(128) fast_fwd_kernel(const FastParams P, const float* __restrict__ x, int64_t n,
                                                       const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                       float* __restrict__ coeff) {
  n = resolve_n(n, n_dev);
  const float msize = fast_msize(P);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
    for (int k = 0; k < P.xdim; ++k) xr[k] = x[i * P.xdim + k];
    TapSet<DC, NEAR_C> tc;
    coeff_taps<DC, NEAR_C>(P, xr, tc);
    float* frow = feats ? feats + i * P.W : nullptr;
    float* crow = coeff ? coeff + i * P.W : nullptr;
    for (int l = 0; l < P.n_levels; ++l) {
      const FastLevel L = P.lv[l];
      TapSet<DB, NEAR_B> tb;
      basis_taps<DB, NEAR_B>(P, L, xr, msize, tb);
      if ((L.C & 3) == 0) {
        for (int c0 = 0; c0 < L.C; c0 += 4) {
          float b[4], ca[2], cb[2];
          gather_vec<DB, NEAR_B, 4>(L.data, L.C, c0, tb, b);
          gather_vec<DC, NEAR_C, 2>(P.cdata, P.W, L.col + c0, tc, ca);
          gather_vec<DC, NEAR_C, 2>(P.cdata, P.W, L.col + c0 + 2, tc, cb);
          const int o = L.col + c0;
          if (frow) {
            *reinterpret_cast<float2*>(frow + o) = make_float2(b[0] * ca[0], b[1] * ca[1]);
            *reinterpret_cast<float2*>(frow + o + 2) = make_float2(b[2] * cb[0], b[3] * cb[1]);
          }
          if (crow) {
            *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
            *reinterpret_cast<float2*>(crow + o + 2) = make_float2(cb[0], cb[1]);
          }
        }
      } else {
        for (int c0 = 0; c0 < L.C; c0 += 2) {
          float b[2], ca[2];
          gather_vec<DB, NEAR_B, 2>(L.data, L.C, c0, tb, b);
          gather_vec<DC, NEAR_C, 2>(P.cdata, P.W, L.col + c0, tc, ca);
          const int o = L.col + c0;
          if (frow) *reinterpret_cast<float2*>(frow + o) = make_float2(b[0] * ca[0], b[1] * ca[1]);
          if (crow) *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
        }
      }
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

static bool build_params(const ffb_field_desc& d, FastParams& P, int op_index[FAST_MAX_LEVELS + 1]) {
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

using namespace ffb;

#define FAST_DISPATCH(KERNEL, ...)                                                                              \
  do {                                                                                                          \
    const bool nb = f->h.ops[f->h.bterms[0].op[0]].nearest, nc = f->h.ops[f->h.cterms[0].op[0]].nearest;        \
    const int db = P.in_dim, dc = P.xdim;                                                                       \
    if (db == 3 && dc == 3 && !nb && !nc) KERNEL<3, 3, false, false><<<grid, 128, 0, s>>>(__VA_ARGS__);         \
    else if (db == 3 && dc == 3 && nb && nc) KERNEL<3, 3, true, true><<<grid, 128, 0, s>>>(__VA_ARGS__);        \
    else if (db == 2 && dc == 2 && !nb && !nc) KERNEL<2, 2, false, false><<<grid, 128, 0, s>>>(__VA_ARGS__);    \
    else if (db == 2 && dc == 2 && nb && nc) KERNEL<2, 2, true, true><<<grid, 128, 0, s>>>(__VA_ARGS__);        \
    else if (db == 2 && dc == 3 && !nb && !nc) KERNEL<2, 3, false, false><<<grid, 128, 0, s>>>(__VA_ARGS__);    \
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

extern "C" {

int ffb_field_fast_eligible(ffb_field_t f) {
  if (!f) return 0;
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  return (build_params(f->h, P, idx) && mode_supported(f, P)) ? 1 : 0;
}

int ffb_field_fast_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  FFB_REQUIRE(build_params(f->h, P, idx) && mode_supported(f, P), "descriptor is not eligible for the fast path");
  if (n <= 0) return FFB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = blocks_for(n, 128, sm_count() * 64);
  FAST_DISPATCH(fast_fwd_kernel, P, x, n, n_dev, feats, coeff);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_fast_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                       float* const* h_grads, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  FFB_REQUIRE(build_params(f->h, P, idx) && mode_supported(f, P), "descriptor is not eligible for the fast path");
  if (n <= 0) return FFB_OK;
  FastGrads G;
  G.c = h_grads ? h_grads[idx[0]] : f->h.ops[idx[0]].grad;
  for (int l = 0; l < FAST_MAX_LEVELS; ++l) G.b[l] = l < P.n_levels ? (h_grads ? h_grads[idx[l + 1]] : f->h.ops[idx[l + 1]].grad) : nullptr;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = blocks_for(n, 128, sm_count() * 64);
  FAST_DISPATCH(fast_bwd_kernel, P, G, x, n, n_dev, g_feats, g_coeff);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
