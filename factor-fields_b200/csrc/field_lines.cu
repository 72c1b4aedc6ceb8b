// field_lines.cu — specialised field kernels for the CP-factorised presets (README_FactorField.md `-CP`: coefficient VECTOR
// [1, W, Hc, 1] over coordinate 0 x per-level products of (d-1) basis LINES [1, C, H, 1] over the mapped coordinates 1..d-1;
// FactorFields.py:437-441 (vec coefficient), :497-509 (cp basis)).  The factors are tiny (nerf -CP: 12 lines of 64 KB + one
// 196 KB vector), so unlike the grid fields they FIT in shared memory one level at a time:
//   forward : one persistent CTA per SM walks the levels; per level the two lines arrive by TMA bulk copies (cp.async.bulk ->
//             mbarrier) and the level's channel slice of the coefficient vector by cooperative loads; every query of the CTA
//             then interpolates from shared memory (C/4 lanes per query, 16-byte conflict-free reads) and writes its C
//             features as one coalesced 16-byte piece per lane.
//   backward: the same walk with SHARED-MEMORY-PRIVATISED accumulation — the gradient of a level's lines / coefficient slice is
//             summed in shared memory with red.shared and flushed ONCE per CTA and pass with 16-byte red.global.add.v4.f32
//             (the generic kernel's per-element global atomics collide on a few hundred rows: 7.8 ms at the -CP bench shape).
// Tap arithmetic is the shared unfused fp32 sequence of ffb_math.h (same indices / weights as the generic kernels and the
// reference's grid_sample on an [H, 1] image: the x corner outside the single column drops out, leaving a 1-D interpolation).
#include "ffb_common.cuh"
#include "ffb_math.h"
#include "tc_common.cuh"

struct ffb_field {
  ffb_field_desc h;
  ffb_field_desc* d;
};

namespace ffb {

constexpr int LN_MAX_LEVELS = 8;
constexpr int LN_THREADS = 1024;

struct LineRef {
  const float* data;
  int H, axis, space, align, border, level, op;
};

struct LineParams {
  int xdim, in_dim, mapping, n_levels, C, W;      // C channels per level, W = n_levels * C = coefficient width
  float lo[3], hi[3], freq[LN_MAX_LEVELS];
  LineRef coeff;
  LineRef line[LN_MAX_LEVELS][2];
  int n_lines;                                     // lines per level: in_dim - 1 (1 or 2)
};

struct Tap1 {
  int i0;
  float w0, w1;
  bool ok0, ok1;
};

__device__ __forceinline__ Tap1 line_tap(const LineParams& P, const LineRef& L, const float* xr, float msize) {
  float u;
  if (L.space == 0) u = normalize_coord(xr[L.axis], P.lo[L.axis], P.hi[L.axis]);
  else u = map_coord(xr[L.axis], P.lo[L.axis], FFB_DIV(msize, P.freq[L.level]), P.mapping, nullptr);
  const Axis ax = linear_axis(source_index(u, L.H, L.align, L.border));
  Tap1 t;
  t.i0 = ax.i0;
  t.w0 = ax.w0;
  t.w1 = ax.w1;
  t.ok0 = ax.i0 >= 0 && ax.i0 < L.H;
  t.ok1 = ax.i0 + 1 >= 0 && ax.i0 + 1 < L.H;
  return t;
}

__device__ __forceinline__ float4 lerp4(const float* s, int stride, int c0, const Tap1& t) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t.ok0) {
    const float4 a = *reinterpret_cast<const float4*>(s + (size_t)t.i0 * stride + c0);
    v.x += a.x * t.w0; v.y += a.y * t.w0; v.z += a.z * t.w0; v.w += a.w * t.w0;
  }
  if (t.ok1) {
    const float4 b = *reinterpret_cast<const float4*>(s + (size_t)(t.i0 + 1) * stride + c0);
    v.x += b.x * t.w1; v.y += b.y * t.w1; v.z += b.z * t.w1; v.w += b.w * t.w1;
  }
  return v;
}

__device__ __forceinline__ float ln_msize(const LineParams& P) {
  float m = FFB_SUB(P.hi[0], P.lo[0]);
  for (int k = 1; k < P.in_dim; ++k) m = fmaxf(m, FFB_SUB(P.hi[k], P.lo[k]));
  return m;
}

// ---------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LN_THREADS, 1) lines_fwd_kernel(const LineParams P, const float* __restrict__ x, int64_t n_cap,
                                                                  const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                                  float* __restrict__ coeff_out, float* __restrict__ basis_out) {
  extern __shared__ __align__(128) float ln_smem[];
  __shared__ uint64_t bar;
  const int64_t n = resolve_n(n_cap, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, W = P.W, lpq = C >> 2, qpw = 32 / lpq;            // lanes per query, queries per warp
  const int H = P.line[0][0].H, Hc = P.coeff.H;
  float* sA = ln_smem;
  float* sB = sA + (size_t)H * C;
  float* sC = sB + (size_t)(P.n_lines > 1 ? H * C : 0);
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t q0 = per * blockIdx.x, q1 = q0 + per < n ? q0 + per : n;
  const float msize = ln_msize(P);
  if (tid == 0) mbar_init(&bar, 1);
  __syncthreads();
  uint32_t phase = 0;
  for (int l = 0; l < P.n_levels; ++l) {
    __syncthreads();                               // every read of the previous level's tables is done
    if (warp == 0 && elect_one()) {
      const uint32_t bytes = (uint32_t)(H * C * sizeof(float));
      mbar_expect_tx(&bar, bytes * (uint32_t)P.n_lines);
      bulk_g2s(sA, P.line[l][0].data, bytes, &bar);
      if (P.n_lines > 1) bulk_g2s(sB, P.line[l][1].data, bytes, &bar);
    }
    for (int idx = tid; idx < Hc * lpq; idx += LN_THREADS) {   // the level's channel slice of the coefficient vector
      const int row = idx / lpq, c4 = idx % lpq;
      reinterpret_cast<float4*>(sC)[idx] = __ldg(reinterpret_cast<const float4*>(P.coeff.data + (size_t)row * W + l * C) + c4);
    }
    mbar_wait(&bar, phase);
    phase ^= 1;
    __syncthreads();
    const int c0 = (lane % lpq) * 4;
    for (int64_t q = q0 + warp * qpw + lane / lpq; q < q1; q += (LN_THREADS / 32) * qpw) {
      float xr[3];
      for (int k = 0; k < P.xdim; ++k) xr[k] = x[q * P.xdim + k];
      const Tap1 tc = line_tap(P, P.coeff, xr, msize), ta = line_tap(P, P.line[l][0], xr, msize);
      float4 b = lerp4(sA, C, c0, ta);
      if (P.n_lines > 1) {
        const Tap1 tb = line_tap(P, P.line[l][1], xr, msize);
        const float4 b2 = lerp4(sB, C, c0, tb);
        b.x *= b2.x; b.y *= b2.y; b.z *= b2.z; b.w *= b2.w;
      }
      const float4 c = lerp4(sC, C, c0, tc);
      const size_t o = (size_t)q * W + (size_t)l * C + c0;
      if (feats) *reinterpret_cast<float4*>(feats + o) = make_float4(b.x * c.x, b.y * c.y, b.z * c.z, b.w * c.w);
      if (coeff_out) *reinterpret_cast<float4*>(coeff_out + o) = c;
      if (basis_out) *reinterpret_cast<float4*>(basis_out + o) = b;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward: passes over (level, CB-channel block); values and accumulators of the pass in shared memory
// ---------------------------------------------------------------------------------------------------------
// 4 floats += g in shared memory.  There is no native fp32 add on shared memory: atomicAdd(float*) compiles to one
// ATOMS.CAST.SPIN loop per float.  (Measured alternative: one 128-bit compare-and-swap loop per 4 floats, ATOMS.CAS.128 —
// 1.9x SLOWER at the -CP bench shape, 2.45 vs 1.29 ms: the wide CAS retries far more often.)
__device__ __forceinline__ void red_shared4(float* s, int stride, int c0, const Tap1& t, const float4 g) {
  if (t.ok0) {
    float* p = s + (size_t)t.i0 * stride + c0;
    atomicAdd(p, g.x * t.w0); atomicAdd(p + 1, g.y * t.w0); atomicAdd(p + 2, g.z * t.w0); atomicAdd(p + 3, g.w * t.w0);
  }
  if (t.ok1) {
    float* p = s + (size_t)(t.i0 + 1) * stride + c0;
    atomicAdd(p, g.x * t.w1); atomicAdd(p + 1, g.y * t.w1); atomicAdd(p + 2, g.z * t.w1); atomicAdd(p + 3, g.w * t.w1);
  }
}

struct LineGrads {
  float* coeff;
  float* line[LN_MAX_LEVELS][2];
};

__global__ void __launch_bounds__(LN_THREADS, 1) lines_bwd_kernel(const LineParams P, const LineGrads G, const float* __restrict__ x,
                                                                  int64_t n_cap, const int32_t* __restrict__ n_dev,
                                                                  const float* __restrict__ g_feats, const float* __restrict__ g_coeff,
                                                                  int CB, int walk) {
  extern __shared__ __align__(128) float ln_smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int C = P.C, W = P.W, lpq = CB >> 2, qpw = 32 / lpq;
  const int H = P.line[0][0].H, Hc = P.coeff.H;
  const int nl = P.n_lines;
  // values: A, B, Cf      accumulators: gA, gB, gCf     (row stride CB floats)
  float* vA = ln_smem;
  float* vB = vA + (size_t)H * CB;
  float* vC = vB + (size_t)(nl > 1 ? H * CB : 0);
  float* gA = vC + (size_t)Hc * CB;
  float* gB = gA + (size_t)H * CB;
  float* gC = gB + (size_t)(nl > 1 ? H * CB : 0);
  const int total = (nl * H + Hc) * CB;                     // floats of the value block == floats of the accumulator block
  const int64_t per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t q0 = per * blockIdx.x, q1 = q0 + per < n ? q0 + per : n;
  const float msize = ln_msize(P);
  const int blocks_per_level = C / CB;
  for (int pass = 0; pass < P.n_levels * blocks_per_level; ++pass) {
    const int l = pass / blocks_per_level, cb0 = (pass % blocks_per_level) * CB, col0 = l * C + cb0;
    __syncthreads();
    for (int idx = tid; idx < H * lpq; idx += LN_THREADS) {
      const int row = idx / lpq, c4 = idx % lpq;
      reinterpret_cast<float4*>(vA)[idx] = __ldg(reinterpret_cast<const float4*>(P.line[l][0].data + (size_t)row * C + cb0) + c4);
      if (nl > 1) reinterpret_cast<float4*>(vB)[idx] = __ldg(reinterpret_cast<const float4*>(P.line[l][1].data + (size_t)row * C + cb0) + c4);
    }
    for (int idx = tid; idx < Hc * lpq; idx += LN_THREADS) {
      const int row = idx / lpq, c4 = idx % lpq;
      reinterpret_cast<float4*>(vC)[idx] = __ldg(reinterpret_cast<const float4*>(P.coeff.data + (size_t)row * W + col0) + c4);
    }
    for (int idx = tid; idx < total / 4; idx += LN_THREADS) reinterpret_cast<float4*>(gA)[idx] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    const int c0 = (lane % lpq) * 4;
    // walk = 1: every query slot of every warp ("walker") owns one contiguous piece of the CTA's range, so the queries in flight
    // at any moment are far apart; walk = 0: a warp takes qpw CONSECUTIVE queries per step — neighbouring samples of a ray, which
    // add into the same rows, and the compare-and-swap loops of the shared-memory float adds then retry against each other
    const int walkers = (LN_THREADS / 32) * qpw, wid = warp * qpw + lane / lpq;
    const int64_t sub = (q1 - q0 + walkers - 1) / walkers;
    int64_t qb = q0 + wid, qe = q1, qs = walkers;
    if (walk) { qb = q0 + wid * sub; qe = qb + sub < q1 ? qb + sub : q1; qs = 1; }
    for (int64_t q = qb; q < qe; q += qs) {
      float xr[3];
      for (int k = 0; k < P.xdim; ++k) xr[k] = x[q * P.xdim + k];
      const Tap1 tc = line_tap(P, P.coeff, xr, msize), ta = line_tap(P, P.line[l][0], xr, msize);
      Tap1 tb = ta;
      const float4 a = lerp4(vA, CB, c0, ta);
      float4 b = make_float4(1.f, 1.f, 1.f, 1.f);
      if (nl > 1) {
        tb = line_tap(P, P.line[l][1], xr, msize);
        b = lerp4(vB, CB, c0, tb);
      }
      const float4 c = lerp4(vC, CB, c0, tc);
      const size_t o = (size_t)q * W + col0 + c0;
      const float4 g = g_feats ? __ldg(reinterpret_cast<const float4*>(g_feats + o)) : make_float4(0.f, 0.f, 0.f, 0.f);
      // feats = a * b * c  (coeff row = c):  d/da = g c b,  d/db = g c a,  d/dc = g a b (+ g_coeff)
      const float4 gc_ = make_float4(g.x * c.x, g.y * c.y, g.z * c.z, g.w * c.w);
      if (G.line[l][0]) red_shared4(gA, CB, c0, ta, make_float4(gc_.x * b.x, gc_.y * b.y, gc_.z * b.z, gc_.w * b.w));
      if (nl > 1 && G.line[l][1]) red_shared4(gB, CB, c0, tb, make_float4(gc_.x * a.x, gc_.y * a.y, gc_.z * a.z, gc_.w * a.w));
      if (G.coeff) {
        float4 gcf = make_float4(g.x * a.x * b.x, g.y * a.y * b.y, g.z * a.z * b.z, g.w * a.w * b.w);
        if (g_coeff) {
          const float4 g2 = __ldg(reinterpret_cast<const float4*>(g_coeff + o));
          gcf.x += g2.x; gcf.y += g2.y; gcf.z += g2.z; gcf.w += g2.w;
        }
        red_shared4(gC, CB, c0, tc, gcf);
      }
    }
    __syncthreads();
    // flush: one 16-byte reduction per touched group
    for (int idx = tid; idx < H * lpq; idx += LN_THREADS) {
      const int row = idx / lpq, c4 = idx % lpq;
      if (G.line[l][0]) {
        const float4 v = reinterpret_cast<const float4*>(gA)[idx];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
          float* p = G.line[l][0] + (size_t)row * C + cb0 + c4 * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
      if (nl > 1 && G.line[l][1]) {
        const float4 v = reinterpret_cast<const float4*>(gB)[idx];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
          float* p = G.line[l][1] + (size_t)row * C + cb0 + c4 * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
    }
    if (G.coeff)
      for (int idx = tid; idx < Hc * lpq; idx += LN_THREADS) {
        const int row = idx / lpq, c4 = idx % lpq;
        const float4 v = reinterpret_cast<const float4*>(gC)[idx];
        if (v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f) {
          float* p = G.coeff + (size_t)row * W + col0 + c4 * 4;
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
        }
      }
  }
}

// descriptor -> LineParams; false when the field is not "vector coefficient x products of lines"
static bool build_line_params(const ffb_field_desc& d, LineParams& P) {
  if (d.coeff_width <= 0 || d.basis_width != d.coeff_width || d.basis_is_x || d.basis_perm) return false;
  if (d.n_cterms != 1 || d.cterms[0].n_ops != 1 || d.cterms[0].col != 0) return false;
  if (d.n_bterms < 1 || d.n_bterms > LN_MAX_LEVELS || d.mapping == FFB_MAP_TRIG) return false;
  if (d.xdim != d.in_dim || d.in_dim < 2 || d.in_dim > 3) return false;
  auto line_of = [&](int oi, LineRef& L) -> bool {
    const ffb_gather_op& o = d.ops[oi];
    if (o.nd != 2 || o.size[0] != 1 || o.src[0] >= 0 || o.src[1] < 0 || o.nearest || ((uintptr_t)o.data & 15)) return false;
    // the constant x coordinate must select column 0 with weight 1 (single scene): cst == 0 for both unnormalisations
    if (o.cst[0] != 0.0f) return false;
    L.data = o.data; L.H = o.size[1]; L.axis = o.src[1]; L.space = o.space; L.align = o.align_corners; L.border = o.border;
    L.level = o.level; L.op = oi;
    return true;
  };
  if (!line_of(d.cterms[0].op[0], P.coeff) || P.coeff.space != 0) return false;
  const int nl = d.bterms[0].n_ops;
  if (nl < 1 || nl > 2) return false;
  const int C = d.ops[d.bterms[0].op[0]].C;
  if (C < 4 || (C & 3) || 32 % (C / 4) != 0) return false;
  int col = 0;
  for (int l = 0; l < d.n_bterms; ++l) {
    const ffb_term& T = d.bterms[l];
    if (T.n_ops != nl || T.col != col) return false;
    for (int k = 0; k < nl; ++k) {
      if (!line_of(T.op[k], P.line[l][k]) || P.line[l][k].space != 1 || P.line[l][k].level != l) return false;
      if (d.ops[T.op[k]].C != C || P.line[l][k].H != P.line[0][0].H) return false;
    }
    P.freq[l] = d.freq[l];
    col += C;
  }
  if (col != d.coeff_width || d.ops[d.cterms[0].op[0]].C != col) return false;
  P.xdim = d.xdim; P.in_dim = d.in_dim; P.mapping = d.mapping; P.n_levels = d.n_bterms; P.C = C; P.W = col; P.n_lines = nl;
  for (int k = 0; k < 3; ++k) { P.lo[k] = d.aabb_min[k]; P.hi[k] = d.aabb_max[k]; }
  return true;
}

static size_t lines_fwd_smem(const LineParams& P) { return ((size_t)P.n_lines * P.line[0][0].H + P.coeff.H) * P.C * sizeof(float); }
static int lines_bwd_cb(const LineParams& P) {      // widest channel block whose values + accumulators fit
  for (int cb = P.C; cb >= 4; cb >>= 1) {
    if (P.C % cb || 32 % (cb / 4)) continue;
    if (2 * ((size_t)P.n_lines * P.line[0][0].H + P.coeff.H) * cb * sizeof(float) <= (size_t)smem_optin_bytes() - 1024) return cb;
  }
  return 0;
}

static int g_lines_enabled = 1;
static int g_lines_walk = 0;       // knob "field_lines_walk" (see lines_bwd_kernel): measured SLOWER at the -CP bench shape (1.55 vs 1.29 ms), left off

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_field_lines(int enabled) {
  g_lines_enabled = enabled ? 1 : 0;
  return FFB_OK;
}

int ffb_set_field_lines_walk(int value) {        // reached through ffb_set_tuning("field_lines_walk")
  g_lines_walk = value;
  return FFB_OK;
}

int ffb_field_lines_eligible(ffb_field_t f) {
  if (!f || !g_lines_enabled) return 0;
  LineParams P;
  if (!build_line_params(f->h, P)) return 0;
  if (lines_fwd_smem(P) + 64 > (size_t)smem_optin_bytes() || lines_bwd_cb(P) == 0) return 0;
  if (((size_t)P.line[0][0].H * P.C * sizeof(float)) % 16) return 0;
  return 1;
}

int ffb_field_lines_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  LineParams P;
  FFB_REQUIRE(ffb_field_lines_eligible(f) == 1 && build_line_params(f->h, P), "descriptor is not a vector x lines (CP) field");
  if (n <= 0) return FFB_OK;
  const size_t smem = lines_fwd_smem(P);
  static PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(lines_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_bytes() - 1024));
  const int64_t min_per_cta = 512;
  int64_t grid = (n + min_per_cta - 1) / min_per_cta;
  if (grid > sm_count()) grid = sm_count();
  lines_fwd_kernel<<<(unsigned)grid, LN_THREADS, smem, (cudaStream_t)stream>>>(P, x, n, n_dev, feats, coeff, basis);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_lines_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                        float* const* h_grads, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  LineParams P;
  FFB_REQUIRE(ffb_field_lines_eligible(f) == 1 && build_line_params(f->h, P), "descriptor is not a vector x lines (CP) field");
  if (n <= 0) return FFB_OK;
  LineGrads G;
  auto grad_of = [&](int op) { return h_grads ? h_grads[op] : f->h.ops[op].grad; };
  G.coeff = grad_of(P.coeff.op);
  bool aligned = ((uintptr_t)G.coeff & 15) == 0;
  for (int l = 0; l < LN_MAX_LEVELS; ++l)
    for (int k = 0; k < 2; ++k) {
      G.line[l][k] = (l < P.n_levels && k < P.n_lines) ? grad_of(P.line[l][k].op) : nullptr;
      aligned = aligned && ((uintptr_t)G.line[l][k] & 15) == 0;
    }
  FFB_REQUIRE(aligned, "gradient tensors must be 16-byte aligned");
  const int cb = lines_bwd_cb(P);
  const size_t smem = 2 * ((size_t)P.n_lines * P.line[0][0].H + P.coeff.H) * cb * sizeof(float);
  static PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(lines_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin_bytes() - 1024));
  const int64_t min_per_cta = 512;
  int64_t grid = (n + min_per_cta - 1) / min_per_cta;
  if (grid > sm_count()) grid = sm_count();
  lines_bwd_kernel<<<(unsigned)grid, LN_THREADS, smem, (cudaStream_t)stream>>>(P, G, x, n, n_dev, g_feats, g_coeff, cb, g_lines_walk);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
