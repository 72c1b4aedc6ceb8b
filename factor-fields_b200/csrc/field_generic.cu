// field_generic.cu — descriptor-driven field query (any coeff_type x basis_type x mapping x mode the
// reference's presets use, README_FactorField.md:12-32) and its scatter-add backward.
// Replaces FactorFields.get_coeff / get_basis / get_coding (FactorFields.py:425-533) and the ATen
// grid_sampler_{2d,3d}(_backward) kernels underneath them.  One thread per query; the hot grid x grid
// shapes have their own specialised kernels in field_fast.cu.
#include "ffb_common.cuh"
#include "ffb_math.h"

struct ffb_field {
  ffb_field_desc h;    // host copy
  ffb_field_desc* d;   // device copy
};

namespace ffb {

constexpr int CH = 8;  // channels processed per register chunk

struct Taps {
  int n;
  int off[8];   // texel offset (already multiplied by C); -1 = out of bounds (zeros padding)
  float w[8];
};

__device__ __forceinline__ float max_size(const ffb_field_desc& D) {
  float m = FFB_SUB(D.aabb_max[0], D.aabb_min[0]);
  for (int k = 1; k < D.in_dim; ++k) m = fmaxf(m, FFB_SUB(D.aabb_max[k], D.aabb_min[k]));
  return m;
}

__device__ __forceinline__ void make_taps(const ffb_field_desc& D, const ffb_gather_op& op, const float* xr,
                                          float msize, Taps& t) {
  float c[3];
  for (int k = 0; k < op.nd; ++k) {
    float u;
    int col = op.src[k];
    if (col < 0) {
      u = op.cst[k];
    } else if (op.space == 0) {
      u = normalize_coord(xr[col], D.aabb_min[col], D.aabb_max[col]);
    } else {
      float scale = FFB_DIV(msize, D.freq[op.level]);
      u = map_coord(xr[col], D.aabb_min[col], scale, D.mapping, nullptr);
    }
    c[k] = source_index(u, op.size[k], op.align_corners, op.border);
  }
  if (op.nearest) {
    t.n = 1;
    int idx = 0, stride = 1;
    bool ok = true;
    for (int k = 0; k < op.nd; ++k) {
      int i = nearest_index(c[k]);
      ok = ok && i >= 0 && i < op.size[k];
      idx += i * stride;
      stride *= op.size[k];
    }
    t.off[0] = ok ? idx * op.C : -1;
    t.w[0] = 1.0f;
    return;
  }
  Axis ax[3];
  for (int k = 0; k < op.nd; ++k) ax[k] = linear_axis(c[k]);
  t.n = 1 << op.nd;
  for (int cn = 0; cn < t.n; ++cn) {
    int idx = 0, stride = 1;
    bool ok = true;
    float w = 1.0f;
    for (int k = 0; k < op.nd; ++k) {
      int b = (cn >> k) & 1;
      int i = ax[k].i0 + b;
      ok = ok && i >= 0 && i < op.size[k];
      idx += i * stride;
      stride *= op.size[k];
      float wk = b ? ax[k].w1 : ax[k].w0;
      w = (k == 0) ? wk : FFB_MUL(w, wk);
    }
    t.off[cn] = ok ? idx * op.C : -1;
    t.w[cn] = w;
  }
}

// acc[j] = sum_corners w * data[off + c0 + j]   (ATen accumulation order).  Texels are read with the widest vector their
// alignment allows (16 bytes when C % 4 == 0, 8 bytes when C % 2 == 0: c0 is a multiple of CH = 8 and the tensors are 16-byte
// aligned) — the vm planes (C = 4 / 2) and the 18-channel coefficient lines used to go through scalar loads.
__device__ __forceinline__ void gather_chunk(const ffb_gather_op& op, const Taps& t, int c0, int nc, float acc[CH]) {
#pragma unroll
  for (int j = 0; j < CH; ++j) acc[j] = 0.0f;
  const int vw = ((op.C & 3) == 0 && (nc & 3) == 0) ? 4 : (((op.C & 1) == 0 && (nc & 1) == 0) ? 2 : 1);
  for (int cn = 0; cn < t.n; ++cn) {
    if (t.off[cn] < 0) continue;
    const float* p = op.data + t.off[cn] + c0;
    const float w = t.w[cn];
    if (vw == 4) {
#pragma unroll
      for (int j = 0; j < CH; j += 4)
        if (j < nc) {
          const float4 a = __ldg(reinterpret_cast<const float4*>(p + j));
          acc[j] += a.x * w; acc[j + 1] += a.y * w; acc[j + 2] += a.z * w; acc[j + 3] += a.w * w;
        }
    } else if (vw == 2) {
#pragma unroll
      for (int j = 0; j < CH; j += 2)
        if (j < nc) {
          const float2 a = __ldg(reinterpret_cast<const float2*>(p + j));
          acc[j] += a.x * w; acc[j + 1] += a.y * w;
        }
    } else {
      for (int j = 0; j < nc; ++j) acc[j] += __ldg(p + j) * w;
    }
  }
}

__device__ __forceinline__ int term_channels(const ffb_field_desc& D, const ffb_term& T) { return D.ops[T.op[0]].C; }

// value of column t of the 'x' basis row (FactorFields.py:481-482,510-511)
__device__ __forceinline__ float basis_x_value(const ffb_field_desc& D, const float* xr, float msize, int col) {
  const int d = D.in_dim, F = D.n_freq;
  if (D.mapping == FFB_MAP_TRIG) {
    int i = col / (2 * d), r = col % (2 * d);
    int h = r / d, dd = r % d;
    int t = h * d * F + dd * F + i;
    int a = t / (2 * F), j = t % (2 * F);
    int fj = j < F ? j : j - F;
    float cs;
    float sn = map_coord(xr[a], D.aabb_min[a], FFB_DIV(msize, D.freq[fj]), FFB_MAP_TRIG, &cs);
    return j < F ? sn : cs;
  }
  int i = col / d, dd = col % d;
  return map_coord(xr[dd], D.aabb_min[dd], FFB_DIV(msize, D.freq[i]), D.mapping, nullptr);
}

// Forward.  coeff: [n, Wc] (required when both factors exist), basis_out: optional [n, W] copy of the
// (permuted) basis row, saved for the backward.
__global__ void __launch_bounds__(128) field_generic_fwd(const ffb_field_desc* __restrict__ Dp, const float* __restrict__ x,
                                                         int64_t n, const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                         float* __restrict__ coeff, float* __restrict__ basis_out) {
  const ffb_field_desc& D = *Dp;
  n = resolve_n(n, n_dev);
  const int Wc = D.coeff_width, Wb = D.basis_width;
  const int W = Wb > 0 ? Wb : Wc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
    for (int k = 0; k < D.xdim; ++k) xr[k] = x[i * D.xdim + k];
    const float msize = max_size(D);
    float* crow = coeff ? coeff + i * W : nullptr;
    float* frow = feats ? feats + i * W : nullptr;
    float* brow = basis_out ? basis_out + i * W : nullptr;
    // ---- coefficient terms
    for (int ti = 0; ti < D.n_cterms; ++ti) {
      const ffb_term& T = D.cterms[ti];
      Taps taps[3];
      for (int o = 0; o < T.n_ops; ++o) make_taps(D, D.ops[T.op[o]], xr, msize, taps[o]);
      const int C = term_channels(D, T);
      for (int c0 = 0; c0 < C; c0 += CH) {
        const int nc = min(CH, C - c0);
        float prod[CH], acc[CH];
        gather_chunk(D.ops[T.op[0]], taps[0], c0, nc, prod);
        for (int o = 1; o < T.n_ops; ++o) {
          gather_chunk(D.ops[T.op[o]], taps[o], c0, nc, acc);
#pragma unroll
          for (int j = 0; j < CH; ++j) prod[j] *= acc[j];
        }
        for (int j = 0; j < nc; ++j) {
          if (crow) crow[T.col + c0 + j] = prod[j];
          if (Wb == 0 && frow) frow[T.col + c0 + j] = prod[j];
        }
      }
    }
    // ---- basis terms
    if (D.basis_is_x) {
      for (int q = 0; q < Wb; ++q) {
        float b = basis_x_value(D, xr, msize, q);
        float c = Wc > 0 ? crow[q] : 1.0f;
        if (frow) frow[q] = b * c;
        if (brow) brow[q] = b;
        if (Wc == 0 && crow) crow[q] = b;
      }
      continue;
    }
    for (int ti = 0; ti < D.n_bterms; ++ti) {
      const ffb_term& T = D.bterms[ti];
      Taps taps[3];
      for (int o = 0; o < T.n_ops; ++o) make_taps(D, D.ops[T.op[o]], xr, msize, taps[o]);
      const int C = term_channels(D, T);
      for (int c0 = 0; c0 < C; c0 += CH) {
        const int nc = min(CH, C - c0);
        float prod[CH], acc[CH];
        gather_chunk(D.ops[T.op[0]], taps[0], c0, nc, prod);
        for (int o = 1; o < T.n_ops; ++o) {
          gather_chunk(D.ops[T.op[o]], taps[o], c0, nc, acc);
#pragma unroll
          for (int j = 0; j < CH; ++j) prod[j] *= acc[j];
        }
        for (int j = 0; j < nc; ++j) {
          const int q = T.col + c0 + j;
          const int p = D.basis_perm ? D.basis_perm[q] : q;
          const float c = Wc > 0 ? crow[p] : 1.0f;
          if (frow) frow[p] = prod[j] * c;
          if (brow) brow[p] = prod[j];
          if (Wc == 0 && crow) crow[p] = prod[j];
        }
      }
    }
  }
}

struct GradPtrs {
  float* p[FFB_MAX_OPS];
};

// grad[off + c0 + j] += w * g[j] with vector reductions (red.global.add.v4 / v2.f32) wherever the texel alignment allows:
// one 16-byte reduction per corner of a 4-channel plane texel instead of four scalar atomics.
__device__ __forceinline__ void scatter_chunk(float* grad, int C, const Taps& t, int c0, int nc, const float g[CH]) {
  if (!grad) return;
  const int vw = ((C & 3) == 0 && (nc & 3) == 0) ? 4 : (((C & 1) == 0 && (nc & 1) == 0) ? 2 : 1);
  for (int cn = 0; cn < t.n; ++cn) {
    if (t.off[cn] < 0) continue;
    float* p = grad + t.off[cn] + c0;
    const float w = t.w[cn];
    if (vw == 4) {
#pragma unroll
      for (int j = 0; j < CH; j += 4)
        if (j < nc) {
          const float a = g[j] * w, b = g[j + 1] * w, c = g[j + 2] * w, d = g[j + 3] * w;
          if (a != 0.0f || b != 0.0f || c != 0.0f || d != 0.0f)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + j), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
        }
    } else if (vw == 2) {
#pragma unroll
      for (int j = 0; j < CH; j += 2)
        if (j < nc) {
          const float a = g[j] * w, b = g[j + 1] * w;
          if (a != 0.0f || b != 0.0f) asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p + j), "f"(a), "f"(b) : "memory");
        }
    } else {
      for (int j = 0; j < nc; ++j) {
        float v = g[j] * w;
        if (v != 0.0f) atomicAdd(p + j, v);
      }
    }
  }
}

__device__ __forceinline__ void bwd_term(const ffb_field_desc& D, const ffb_term& T, const float* xr, float msize,
                                         const float* grow /* gradient w.r.t. the term's columns, indexed by column */,
                                         const int32_t* perm, const GradPtrs& G) {
  Taps taps[3];
  for (int o = 0; o < T.n_ops; ++o) make_taps(D, D.ops[T.op[o]], xr, msize, taps[o]);
  const int C = term_channels(D, T);
  for (int c0 = 0; c0 < C; c0 += CH) {
    const int nc = min(CH, C - c0);
    float g[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) g[j] = 0.0f;
    for (int j = 0; j < nc; ++j) {
      const int q = T.col + c0 + j;
      g[j] = grow[perm ? perm[q] : q];
    }
    if (T.n_ops == 1) {
      scatter_chunk(G.p[T.op[0]], D.ops[T.op[0]].C, taps[0], c0, nc, g);
    } else {
      float vals[3][CH];
      for (int o = 0; o < T.n_ops; ++o) gather_chunk(D.ops[T.op[o]], taps[o], c0, nc, vals[o]);
      for (int o = 0; o < T.n_ops; ++o) {
        float go[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          float v = g[j];
          for (int oo = 0; oo < T.n_ops; ++oo)
            if (oo != o) v *= vals[oo][j];
          go[j] = v;
        }
        scatter_chunk(G.p[T.op[o]], D.ops[T.op[o]].C, taps[o], c0, nc, go);
      }
    }
  }
}

// Backward.  gc_row / gb_row are per-query scratch rows [n, W] prepared here from g_feats, the saved
// coeff and basis rows: gc = g_feats*basis + g_coeff ; gb = g_feats*coeff (both indexed by OUTPUT column).
__global__ void __launch_bounds__(128) field_generic_bwd(const ffb_field_desc* __restrict__ Dp, const float* __restrict__ x,
                                                         int64_t n, const int32_t* __restrict__ n_dev,
                                                         const float* __restrict__ g_feats, const float* __restrict__ g_coeff,
                                                         float* __restrict__ coeff /* in: coeff rows, out: gc */,
                                                         float* __restrict__ basis /* in: basis rows, out: gb */,
                                                         const GradPtrs G) {
  const ffb_field_desc& D = *Dp;
  n = resolve_n(n, n_dev);
  const int Wc = D.coeff_width, Wb = D.basis_width;
  const int W = Wb > 0 ? Wb : Wc;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
    for (int k = 0; k < D.xdim; ++k) xr[k] = x[i * D.xdim + k];
    const float msize = max_size(D);
    float* crow = coeff + i * W;
    float* brow = basis + i * W;
    const float* gf = g_feats ? g_feats + i * W : nullptr;
    const float* gcf = g_coeff ? g_coeff + i * W : nullptr;
    for (int p = 0; p < W; ++p) {
      const float g = gf ? gf[p] : 0.0f;
      const float gcc = gcf ? gcf[p] : 0.0f;
      if (Wc > 0 && Wb > 0) {
        const float c = crow[p], b = brow[p];
        crow[p] = g * b + gcc;
        brow[p] = g * c;
      } else if (Wc > 0) {
        crow[p] = g + gcc;   // get_coding returns (coeff, coeff)
      } else {
        brow[p] = g + gcc;   // (basis, basis)
      }
    }
    for (int ti = 0; ti < D.n_cterms; ++ti) bwd_term(D, D.cterms[ti], xr, msize, crow, nullptr, G);
    if (!D.basis_is_x)
      for (int ti = 0; ti < D.n_bterms; ++ti) bwd_term(D, D.bterms[ti], xr, msize, brow, D.basis_perm, G);
  }
}

struct FreqTable {
  float f[FFB_MAX_FREQ];     // by value in the kernel parameters: no allocation, no host->device copy, capturable in a CUDA graph
};

__global__ void grid_mapping_kernel(const float* __restrict__ x, int64_t n, int in_dim, float3 lo, float msize, const FreqTable freq,
                                    int F, int mapping, float* __restrict__ out) {
  const int per = in_dim * F;
  const float los[3] = {lo.x, lo.y, lo.z};
  for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n * per; t += (int64_t)gridDim.x * blockDim.x) {
    int64_t i = t / per;
    int r = (int)(t % per);
    int d = r / F, f = r % F;
    float cs;
    float v = map_coord(x[i * in_dim + d], los[d], FFB_DIV(msize, freq.f[f]), mapping, &cs);
    if (mapping == FFB_MAP_TRIG) {
      out[(i * in_dim + d) * 2 * F + f] = v;
      out[(i * in_dim + d) * 2 * F + F + f] = cs;
    } else {
      out[t] = v;
    }
  }
}

}  // namespace ffb

using namespace ffb;

static int validate_desc(const ffb_field_desc* d) {
  FFB_REQUIRE(d->xdim >= 1 && d->xdim <= 3 && d->in_dim >= 1 && d->in_dim <= d->xdim, "bad xdim/in_dim");
  FFB_REQUIRE(d->n_freq >= 0 && d->n_freq <= FFB_MAX_FREQ, "too many frequency bands");
  FFB_REQUIRE(d->n_ops >= 0 && d->n_ops <= FFB_MAX_OPS, "too many gather ops");
  FFB_REQUIRE(d->n_cterms >= 0 && d->n_cterms <= FFB_MAX_TERMS && d->n_bterms >= 0 && d->n_bterms <= FFB_MAX_TERMS, "too many terms");
  FFB_REQUIRE(d->coeff_width > 0 || d->basis_width > 0, "coeff_type and basis_type are both 'none'");
  FFB_REQUIRE(d->coeff_width == 0 || d->basis_width == 0 || d->coeff_width == d->basis_width,
              "coefficient and basis rows must have the same width");
  for (int i = 0; i < d->n_ops; ++i) {
    const ffb_gather_op& o = d->ops[i];
    FFB_REQUIRE(o.data != nullptr && o.C > 0 && o.nd >= 1 && o.nd <= 3, "bad gather op");
    FFB_REQUIRE(o.space == 0 || (o.level >= 0 && o.level < d->n_freq), "gather op level out of range");
    FFB_REQUIRE(o.space == 0 || d->mapping != FFB_MAP_TRIG, "trigonometric mapping is only defined for the 'x' basis");
    for (int k = 0; k < o.nd; ++k) FFB_REQUIRE(o.size[k] >= 1 && o.src[k] < d->xdim, "bad gather op size/src");
  }
  for (int pass = 0; pass < 2; ++pass) {
    const ffb_term* T = pass ? d->bterms : d->cterms;
    const int nT = pass ? d->n_bterms : d->n_cterms;
    const int W = pass ? d->basis_width : d->coeff_width;
    for (int i = 0; i < nT; ++i) {
      FFB_REQUIRE(T[i].n_ops >= 1 && T[i].n_ops <= 3, "term must have 1..3 ops");
      int C = -1;
      for (int o = 0; o < T[i].n_ops; ++o) {
        FFB_REQUIRE(T[i].op[o] >= 0 && T[i].op[o] < d->n_ops, "term op index out of range");
        int Co = d->ops[T[i].op[o]].C;
        FFB_REQUIRE(C < 0 || C == Co, "ops of one term must have equal channel counts");
        C = Co;
      }
      FFB_REQUIRE(T[i].col >= 0 && T[i].col + C <= W, "term columns out of range");
    }
  }
  return FFB_OK;
}

extern "C" {

int ffb_field_create(const ffb_field_desc* h_desc, ffb_field_t* out) {
  FFB_REQUIRE(h_desc && out, "null argument");
  int rc = validate_desc(h_desc);
  if (rc != FFB_OK) return rc;
  ffb_field* f = new ffb_field();
  f->h = *h_desc;
  f->d = nullptr;
  cudaError_t e = cudaMalloc(&f->d, sizeof(ffb_field_desc));
  if (e == cudaSuccess) e = cudaMemcpy(f->d, h_desc, sizeof(ffb_field_desc), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (f->d) cudaFree(f->d);
    delete f;
    return check_cuda(e, "ffb_field_create");
  }
  *out = f;
  return FFB_OK;
}

int ffb_field_destroy(ffb_field_t f) {
  if (!f) return FFB_OK;
  if (f->d) cudaFree(f->d);
  delete f;
  return FFB_OK;
}

int ffb_field_generic_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff,
                          float* basis_out, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  if (n <= 0) return FFB_OK;
  FFB_REQUIRE(!(f->h.coeff_width > 0 && f->h.basis_width > 0) || coeff, "coeff buffer required when both factors exist");
  for (int i = 0; i < f->h.n_ops; ++i) FFB_REQUIRE(((uintptr_t)f->h.ops[i].data & 15) == 0, "factor tensors must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  field_generic_fwd<<<blocks_for(n, 128, sm_count() * 32), 128, 0, s>>>(f->d, x, n, n_dev, feats, coeff, basis_out);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_generic_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats,
                          const float* g_coeff, float* const* h_grads, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  if (n <= 0) return FFB_OK;
  GradPtrs G;
  for (int i = 0; i < FFB_MAX_OPS; ++i) G.p[i] = i < f->h.n_ops ? (h_grads ? h_grads[i] : f->h.ops[i].grad) : nullptr;
  for (int i = 0; i < f->h.n_ops; ++i)
    FFB_REQUIRE(((uintptr_t)f->h.ops[i].data & 15) == 0 && ((uintptr_t)G.p[i] & 15) == 0, "factor and gradient tensors must be 16-byte aligned");
  cudaStream_t s = (cudaStream_t)stream;
  const int W = f->h.basis_width > 0 ? f->h.basis_width : f->h.coeff_width;
  float *c = nullptr, *b = nullptr;
  keep_pool_cached();
  FFB_CUDA(cudaMallocAsync(&c, sizeof(float) * n * W, s));
  cudaError_t e = cudaMallocAsync(&b, sizeof(float) * n * W, s);
  if (e != cudaSuccess) {
    cudaFreeAsync(c, s);
    return check_cuda(e, "cudaMallocAsync");
  }
  const unsigned blocks = blocks_for(n, 128, sm_count() * 32);
  field_generic_fwd<<<blocks, 128, 0, s>>>(f->d, x, n, n_dev, nullptr, c, b);
  g_launches.fetch_add(1);
  // "field_deterministic": ONE thread walks the queries in order, so every gradient element is summed in a fixed order and
  // the result is bit-reproducible run to run (the reference's CUDA backward is not; its CPU backward is).  For debugging and
  // gradient tests at small sizes — it is serial.
  if (g_deterministic) field_generic_bwd<<<1, 1, 0, s>>>(f->d, x, n, n_dev, g_feats, g_coeff, c, b, G);
  else field_generic_bwd<<<blocks, 128, 0, s>>>(f->d, x, n, n_dev, g_feats, g_coeff, c, b, G);
  g_launches.fetch_add(1);
  cudaError_t le = cudaGetLastError();
  cudaFreeAsync(c, s);
  cudaFreeAsync(b, s);
  return check_cuda(le, "field_generic_bwd");
}

int ffb_grid_mapping(const float* x, int64_t n, int32_t in_dim, const float* h_aabb_min, const float* h_aabb_max,
                     const float* h_freq, int32_t n_freq, int32_t mapping, float* out, void* stream) {
  FFB_REQUIRE(x && out && h_aabb_min && h_aabb_max && h_freq, "null argument");
  FFB_REQUIRE(in_dim >= 1 && in_dim <= 3 && n_freq >= 1 && n_freq <= FFB_MAX_FREQ, "bad in_dim / n_freq");
  if (n <= 0) return FFB_OK;
  cudaStream_t s = (cudaStream_t)stream;
  float msize = h_aabb_max[0] - h_aabb_min[0];
  for (int k = 1; k < in_dim; ++k) msize = fmaxf(msize, h_aabb_max[k] - h_aabb_min[k]);
  FreqTable dfreq;
  for (int i = 0; i < FFB_MAX_FREQ; ++i) dfreq.f[i] = i < n_freq ? h_freq[i] : 1.0f;
  float3 lo = make_float3(h_aabb_min[0], in_dim > 1 ? h_aabb_min[1] : 0.f, in_dim > 2 ? h_aabb_min[2] : 0.f);
  grid_mapping_kernel<<<blocks_for(n * in_dim * n_freq, 256, sm_count() * 16), 256, 0, s>>>(x, n, in_dim, lo, msize, dfreq, n_freq,
                                                                                         mapping, out);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
