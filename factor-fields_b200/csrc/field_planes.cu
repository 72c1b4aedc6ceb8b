// field_planes.cu — specialised field kernels for the vector-matrix (VM, TensoRF-style) preset of README_FactorField.md (`-vm`:
// coeff_type 'vm' = three coefficient LINES [1, Cc, H, 1] along the axes vecMode = [2, 1, 0], concatenated (FactorFields.py:
// 443-450); basis_type 'vm' = three PLANES [1, C_l, H, W] per level over the axis pairs matMode = [[0,1],[0,2],[1,2]], followed by
// the global column re-ordering view(N, F, -1).permute(0, 2, 1) (:491-496, :514-515)).
//
// Same design as field_fast.cu (one thread per query, consecutive threads = consecutive samples of a ray, channels-last texels
// read as float4 / float2, x-neighbour corners adjacent, scatter with red.global.add.v4/v2.f32), adapted to the factorisation:
// the 3 x Cc coefficient row of a query is interpolated first and parked in shared memory, because the re-ordering scatters the
// basis columns across it; every plane sample is then combined with its coefficient, and in the backward pass REPLACED by the
// coefficient gradient in place.  Two generations: vm_fwd_kernel / vm_bwd_kernel (one column of a [W][threads] tile per thread,
// per-lane row traffic to global memory) and — the default — vm_fwd2_kernel / vm_bwd2_kernel, which move whole 32-query WARP
// TILES to / from global memory as coalesced 16-byte pieces (see below).  The plane texels are always re-gathered by the
// backward pass (L2 resident); the coefficient row comes from the forward's output.
#include "field_fast.cuh"

namespace ffb {

constexpr int VM_MAX_LEVELS = 8;
constexpr int VM_NT = 128;

struct VmLevel {
  const float* plane[3];
  int pw[3], ph[3];      // plane m: [ph, pw] texels, x (W) <- axis ax0[m], y (H) <- axis ax1[m]
  int C, col;            // channels per plane; basis column of plane 0 (plane m starts at col + m * C)
  float freq;
  int op[3];
};

struct VmParams {
  int mapping, n_levels, W, Cc, Hc;
  float lo[3], hi[3];
  const float* cline[3];
  int caxis[3], cop[3];
  int ax0[3], ax1[3];
  const int32_t* perm;   // device: output column of basis column q
  VmLevel lv[VM_MAX_LEVELS];
};

struct VmGrads {
  float* cline[3];
  float* plane[VM_MAX_LEVELS][3];
};

struct LTap {
  int i0;
  float w0, w1;
  bool ok0, ok1;
};

// v[a] for a run-time axis a without indexing a register array dynamically (that would put it in local memory)
__device__ __forceinline__ float pick3(const float* v, int a) { return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]); }

__device__ __forceinline__ LTap vm_line_tap(const VmParams& P, int m, const float* xr) {
  const int a = P.caxis[m];
  const Axis ax = linear_axis(source_index(normalize_coord(pick3(xr, a), P.lo[a], P.hi[a]), P.Hc, 0, 1));
  LTap t;
  t.i0 = ax.i0; t.w0 = ax.w0; t.w1 = ax.w1;
  t.ok0 = ax.i0 >= 0 && ax.i0 < P.Hc;
  t.ok1 = ax.i0 + 1 >= 0 && ax.i0 + 1 < P.Hc;
  return t;
}

__device__ __forceinline__ float vm_msize(const VmParams& P) {
  float m = FFB_SUB(P.hi[0], P.lo[0]);
  for (int k = 1; k < 3; ++k) m = fmaxf(m, FFB_SUB(P.hi[k], P.lo[k]));
  return m;
}

// one texel of WC = 4k + 2 floats through aligned 16-byte loads (see coeff_row_windows in field_fast.cu): out[c] = texel[c]
template <int WC>
__device__ __forceinline__ void texel_window(const float* __restrict__ data, size_t texel, float out[WC]) {
  constexpr int NQ = (WC + 2) / 4;
  const size_t e0 = texel * WC;
  const bool odd = (e0 & 2) != 0;
  const float4* p = reinterpret_cast<const float4*>(data + (e0 - (odd ? 2 : 0)));
  float win[NQ * 4];
#pragma unroll
  for (int k = 0; k < NQ; ++k) {
    const float4 q = __ldg(p + k);
    win[4 * k] = q.x; win[4 * k + 1] = q.y; win[4 * k + 2] = q.z; win[4 * k + 3] = q.w;
  }
#pragma unroll
  for (int c = 0; c < WC; ++c) out[c] = odd ? win[c + 2] : win[c];
}

// coefficient row of this thread's query -> srow[c * CS] (c < 3 Cc); CS = VM_NT: one column of a [W][threads] tile per thread,
// CS = 1: one row of a [32][W + 1] warp tile per lane
template <int CS = VM_NT>
__device__ __forceinline__ void vm_coeff_row(const VmParams& P, const float* xr, float* srow) {
#pragma unroll 1
  for (int m = 0; m < 3; ++m) {
    const LTap t = vm_line_tap(P, m, xr);
    const float* base = P.cline[m];
    if (P.Cc == 18 && (P.Hc & 1) == 0) {          // 72-byte texels: 5 aligned 16-byte loads per tap instead of 9 8-byte ones
      float a[18], b[18];
#pragma unroll
      for (int c = 0; c < 18; ++c) a[c] = b[c] = 0.0f;
      if (t.ok0) texel_window<18>(base, (size_t)t.i0, a);
      if (t.ok1) texel_window<18>(base, (size_t)(t.i0 + 1), b);
#pragma unroll
      for (int c = 0; c < 18; ++c) {
        float v = 0.0f;
        if (t.ok0) v += a[c] * t.w0;
        if (t.ok1) v += b[c] * t.w1;
        srow[(m * 18 + c) * CS] = v;
      }
      continue;
    }
    for (int c = 0; c < P.Cc; c += 2) {
      float2 v = make_float2(0.f, 0.f);
      if (t.ok0) {
        const float2 a = __ldg(reinterpret_cast<const float2*>(base + (size_t)t.i0 * P.Cc + c));
        v.x += a.x * t.w0; v.y += a.y * t.w0;
      }
      if (t.ok1) {
        const float2 b = __ldg(reinterpret_cast<const float2*>(base + (size_t)(t.i0 + 1) * P.Cc + c));
        v.x += b.x * t.w1; v.y += b.y * t.w1;
      }
      srow[(m * P.Cc + c) * CS] = v.x;
      srow[(m * P.Cc + c + 1) * CS] = v.y;
    }
  }
}

__device__ __forceinline__ void vm_plane_taps(const VmParams& P, const VmLevel& L, int m, const float u[3], TapSet<2, false>& t) {
  float c[3];
  const int size[3] = {L.pw[m], L.ph[m], 1};
  c[0] = source_index(pick3(u, P.ax0[m]), L.pw[m], 1, 0);
  c[1] = source_index(pick3(u, P.ax1[m]), L.ph[m], 1, 0);
  c[2] = 0.0f;
  make_tapset<2, false>(c, size, t);
}

template <int MINB>
__global__ void __launch_bounds__(VM_NT, MINB) vm_fwd_kernel(const VmParams P, const float* __restrict__ x, int64_t n_cap,
                                                             const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                             float* __restrict__ coeff_out, float* __restrict__ basis_out) {
  extern __shared__ float vm_smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const float msize = vm_msize(P);
  float* srow = vm_smem + threadIdx.x;
  const int W = P.W;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xr[k] = x[i * 3 + k];
    vm_coeff_row(P, xr, srow);
    if (coeff_out)
      for (int c = 0; c < W; c += 2) *reinterpret_cast<float2*>(coeff_out + i * W + c) = make_float2(srow[c * VM_NT], srow[(c + 1) * VM_NT]);
#pragma unroll 1
    for (int l = 0; l < P.n_levels; ++l) {
      const VmLevel& L = P.lv[l];
      const float scale = FFB_DIV(msize, L.freq);
      float u[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) u[k] = map_coord(xr[k], P.lo[k], scale, P.mapping, nullptr);
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        TapSet<2, false> tb;
        vm_plane_taps(P, L, m, u, tb);
        float b[4];
        if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb, b);
        else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb, b);
        const int q0 = L.col + m * L.C;
        for (int j = 0; j < L.C; ++j) {
          const int p = __ldg(P.perm + q0 + j);
          if (basis_out) basis_out[i * W + p] = b[j];
          srow[p * VM_NT] *= b[j];                   // coefficient -> feature, in place
        }
      }
    }
    if (feats)
      for (int c = 0; c < W; c += 2) *reinterpret_cast<float2*>(feats + i * W + c) = make_float2(srow[c * VM_NT], srow[(c + 1) * VM_NT]);
  }
}

template <int MINB>
__global__ void __launch_bounds__(VM_NT, MINB) vm_bwd_kernel(const VmParams P, const VmGrads G, const float* __restrict__ x, int64_t n_cap,
                                                             const int32_t* __restrict__ n_dev, const float* __restrict__ g_feats,
                                                             const float* __restrict__ g_coeff, const float* __restrict__ coeff_saved,
                                                             const float* __restrict__ basis_saved) {
  // coeff_saved / basis_saved (both or neither): the [n, W] rows the training forward wrote — the backward then only
  // recomputes tap indices / weights and scatters, instead of re-gathering 126 texel vectors per query
  extern __shared__ float vm_smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const float msize = vm_msize(P);
  float* srow = vm_smem + threadIdx.x;
  const int W = P.W;
  const bool saved = coeff_saved && basis_saved;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float xr[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) xr[k] = x[i * 3 + k];
    if (saved) {
      for (int c = 0; c < W; c += 2) {
        const float2 v = *reinterpret_cast<const float2*>(coeff_saved + i * W + c);
        srow[c * VM_NT] = v.x;
        srow[(c + 1) * VM_NT] = v.y;
      }
    } else {
      vm_coeff_row(P, xr, srow);
    }
    const float* gf = g_feats ? g_feats + i * W : nullptr;
    const float* gcf = g_coeff ? g_coeff + i * W : nullptr;
#pragma unroll 1
    for (int l = 0; l < P.n_levels; ++l) {
      const VmLevel& L = P.lv[l];
      const float scale = FFB_DIV(msize, L.freq);
      float u[3];
#pragma unroll
      for (int k = 0; k < 3; ++k) u[k] = map_coord(xr[k], P.lo[k], scale, P.mapping, nullptr);
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        TapSet<2, false> tb;
        vm_plane_taps(P, L, m, u, tb);
        float b[4], gb[4];
        if (!saved) {
          if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb, b);
          else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb, b);
        }
        const int q0 = L.col + m * L.C;
        for (int j = 0; j < 4; ++j) {
          gb[j] = 0.0f;
          if (j < L.C) {
            const int p = __ldg(P.perm + q0 + j);
            if (saved) b[j] = basis_saved[i * W + p];
            const float g = gf ? gf[p] : 0.0f;
            gb[j] = g * srow[p * VM_NT];                                   // d/d basis = g * coefficient
            srow[p * VM_NT] = g * b[j] + (gcf ? gcf[p] : 0.0f);            // d/d coefficient, in place
          }
        }
        if (G.plane[l][m]) {
          if (L.C == 4) scatter_vec<2, false, 4>(G.plane[l][m], 4, 0, tb, gb);
          else scatter_vec<2, false, 2>(G.plane[l][m], 2, 0, tb, gb);
        }
      }
    }
    // coefficient lines: 1-D scatter of the 3 x Cc gradient row
#pragma unroll 1
    for (int m = 0; m < 3; ++m) {
      float* gl = G.cline[m];
      if (!gl) continue;
      const LTap t = vm_line_tap(P, m, xr);
      for (int c = 0; c < P.Cc; c += 2) {
        const float g0 = srow[(m * P.Cc + c) * VM_NT], g1 = srow[(m * P.Cc + c + 1) * VM_NT];
        if (t.ok0 && t.w0 != 0.0f) red_add_v2(gl + (size_t)t.i0 * P.Cc + c, g0 * t.w0, g1 * t.w0);
        if (t.ok1 && t.w1 != 0.0f) red_add_v2(gl + (size_t)(t.i0 + 1) * P.Cc + c, g0 * t.w1, g1 * t.w1);
      }
    }
  }
}


// ---------------------------------------------------------------------------------------------------------
// Second generation (default; knob "field_planes_v2"): rows through WARP TILES.  A warp owns 32 consecutive queries; their
// [32, W] rows are one contiguous piece of every [n, W] tensor.  Each lane keeps its row in a shared-memory tile of 32 rows x
// (W + 1) floats (odd stride: a column read across the lanes and a row walk by one lane are both conflict-free), and the warp
// moves whole tiles to / from global memory as coalesced 16-byte pieces — one row per lane reads / writes 8-byte pieces 4 W
// bytes apart (forward: 26 M L2 write sectors for 5.7 M of payload; backward: the permuted scalar reads of the upstream
// gradient row hit a different sector per lane, L1 hit rate 16 %, 46 warps stalled on long_scoreboard).
// The backward pass reads the coefficient row the forward returned anyway (`coeff`, public output) instead of re-gathering
// the three lines, and scatters the line gradient as 16-byte reductions (72-byte texels: {v4 x4, v2} / {v2, v4 x4}).
// ---------------------------------------------------------------------------------------------------------
// global [rows, W] chunk (contiguous, 16-byte aligned) <-> warp tile [32][W + 1]
template <bool STORE>
__device__ __forceinline__ void tile_rows_io(float* __restrict__ g, float* __restrict__ tile, int rows, int W, int lane) {
  const int total = rows * W, nvec = total >> 2;
  int q = (lane * 4) / W, c = lane * 4 - q * W;
  for (int t = lane; t < nvec; t += 32) {
    int idx[4];
    int qq = q, cc = c;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      idx[j] = qq * (W + 1) + cc;
      if (++cc == W) { cc = 0; ++qq; }
    }
    if (STORE) {
      *reinterpret_cast<float4*>(g + 4 * t) = make_float4(tile[idx[0]], tile[idx[1]], tile[idx[2]], tile[idx[3]]);
    } else {
      const float4 v = *reinterpret_cast<const float4*>(g + 4 * t);
      tile[idx[0]] = v.x; tile[idx[1]] = v.y; tile[idx[2]] = v.z; tile[idx[3]] = v.w;
    }
    c += 128;
    while (c >= W) { c -= W; ++q; }
  }
  for (int e = (nvec << 2) + lane; e < total; e += 32) {          // tail of a partial last chunk
    const int qq = e / W, cc = e - qq * W;
    if (STORE) g[e] = tile[qq * (W + 1) + cc];
    else tile[qq * (W + 1) + cc] = g[e];
  }
}

template <int MINB, bool UNROLL>
__global__ void __launch_bounds__(VM_NT, MINB) vm_fwd2_kernel(const VmParams P, const float* __restrict__ x, int64_t n_cap,
                                                              const int32_t* __restrict__ n_dev, float* __restrict__ feats,
                                                              float* __restrict__ coeff_out) {
  extern __shared__ float vm_smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const float msize = vm_msize(P);
  const int lane = threadIdx.x & 31, W = P.W, TS = W + 1;
  float* tile = vm_smem + (size_t)(threadIdx.x >> 5) * 32 * TS;
  float* row = tile + lane * TS;
  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_chunks; k += nwarps) {
    const int64_t i = k * 32 + lane;
    const bool active = i < n;
    const int rows = (n - k * 32) < 32 ? (int)(n - k * 32) : 32;
    float xr[3] = {0.f, 0.f, 0.f};
    if (active) {
#pragma unroll
      for (int d = 0; d < 3; ++d) xr[d] = x[i * 3 + d];
      vm_coeff_row<1>(P, xr, row);
    }
    __syncwarp();
    if (coeff_out) tile_rows_io<true>(coeff_out + k * 32 * W, tile, rows, W, lane);
    __syncwarp();
    if (active) {
#pragma unroll 1
      for (int l = 0; l < P.n_levels; ++l) {
        const VmLevel& L = P.lv[l];
        const float scale = FFB_DIV(msize, L.freq);
        float u[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = map_coord(xr[d], P.lo[d], scale, P.mapping, nullptr);
        if (UNROLL) {                       // the three planes of a level in flight together (12 vector loads per thread)
          float b[3][4];
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            TapSet<2, false> tb;
            vm_plane_taps(P, L, m, u, tb);
            if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb, b[m]);
            else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb, b[m]);
          }
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const int q0 = L.col + m * L.C;
#pragma unroll
            for (int j = 0; j < 4; ++j)
              if (j < L.C) row[__ldg(P.perm + q0 + j)] *= b[m][j];
          }
        } else {
#pragma unroll 1
          for (int m = 0; m < 3; ++m) {
            TapSet<2, false> tb;
            vm_plane_taps(P, L, m, u, tb);
            float b[4];
            if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb, b);
            else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb, b);
            const int q0 = L.col + m * L.C;
            for (int j = 0; j < L.C; ++j) row[__ldg(P.perm + q0 + j)] *= b[j];       // coefficient -> feature, in place
          }
        }
      }
    }
    __syncwarp();
    if (feats) tile_rows_io<true>(feats + k * 32 * W, tile, rows, W, lane);
    __syncwarp();
  }
}

// one 18-float (72-byte) line texel: += w * g[0..17] as 16-byte reductions wherever the address allows
__device__ __forceinline__ void red_texel18(float* __restrict__ p, size_t texel, const float* g, float w) {
  float* q = p + texel * 18;
  if ((texel & 1) == 0) {
#pragma unroll
    for (int c = 0; c < 16; c += 4) red_add_v4(q + c, g[c] * w, g[c + 1] * w, g[c + 2] * w, g[c + 3] * w);
    red_add_v2(q + 16, g[16] * w, g[17] * w);
  } else {
    red_add_v2(q, g[0] * w, g[1] * w);
#pragma unroll
    for (int c = 2; c < 18; c += 4) red_add_v4(q + c, g[c] * w, g[c + 1] * w, g[c + 2] * w, g[c + 3] * w);
  }
}

template <int MINB, bool UNROLL>
__global__ void __launch_bounds__(VM_NT, MINB) vm_bwd2_kernel(const VmParams P, const VmGrads G, const float* __restrict__ x, int64_t n_cap,
                                                              const int32_t* __restrict__ n_dev, const float* __restrict__ g_feats,
                                                              const float* __restrict__ coeff_saved) {
  extern __shared__ float vm_smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const float msize = vm_msize(P);
  const int lane = threadIdx.x & 31, W = P.W, TS = W + 1;
  float* tileA = vm_smem + (size_t)(threadIdx.x >> 5) * 64 * TS;      // coefficient row, replaced by its gradient in place
  float* tileG = tileA + 32 * TS;                                      // upstream gradient row
  float* rowA = tileA + lane * TS;
  const float* rowG = tileG + lane * TS;
  const int64_t n_chunks = (n + 31) / 32;
  const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t k = warp0; k < n_chunks; k += nwarps) {
    const int64_t i = k * 32 + lane;
    const bool active = i < n;
    const int rows = (n - k * 32) < 32 ? (int)(n - k * 32) : 32;
    float xr[3] = {0.f, 0.f, 0.f};
    if (active)
#pragma unroll
      for (int d = 0; d < 3; ++d) xr[d] = x[i * 3 + d];
    tile_rows_io<false>(const_cast<float*>(g_feats) + k * 32 * W, tileG, rows, W, lane);
    if (coeff_saved) tile_rows_io<false>(const_cast<float*>(coeff_saved) + k * 32 * W, tileA, rows, W, lane);
    else if (active) vm_coeff_row<1>(P, xr, rowA);
    __syncwarp();
    if (active) {
#pragma unroll 1
      for (int l = 0; l < P.n_levels; ++l) {
        const VmLevel& L = P.lv[l];
        const float scale = FFB_DIV(msize, L.freq);
        float u[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) u[d] = map_coord(xr[d], P.lo[d], scale, P.mapping, nullptr);
        if (UNROLL) {
          TapSet<2, false> tb[3];
          float b[3][4];
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            vm_plane_taps(P, L, m, u, tb[m]);
            if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb[m], b[m]);
            else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb[m], b[m]);
          }
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const int q0 = L.col + m * L.C;
            float gb[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              gb[j] = 0.0f;
              if (j < L.C) {
                const int p = __ldg(P.perm + q0 + j);
                const float g = rowG[p];
                gb[j] = g * rowA[p];                   // d/d basis = g * coefficient
                rowA[p] = g * b[m][j];                 // d/d coefficient, in place
              }
            }
            if (G.plane[l][m]) {
              if (L.C == 4) scatter_vec<2, false, 4>(G.plane[l][m], 4, 0, tb[m], gb);
              else scatter_vec<2, false, 2>(G.plane[l][m], 2, 0, tb[m], gb);
            }
          }
        } else {
#pragma unroll 1
          for (int m = 0; m < 3; ++m) {
            TapSet<2, false> tb;
            vm_plane_taps(P, L, m, u, tb);
            float b[4], gb[4];
            if (L.C == 4) gather_vec<2, false, 4>(L.plane[m], 4, 0, tb, b);
            else gather_vec<2, false, 2>(L.plane[m], 2, 0, tb, b);
            const int q0 = L.col + m * L.C;
            for (int j = 0; j < 4; ++j) {
              gb[j] = 0.0f;
              if (j < L.C) {
                const int p = __ldg(P.perm + q0 + j);
                const float g = rowG[p];
                gb[j] = g * rowA[p];
                rowA[p] = g * b[j];
              }
            }
            if (G.plane[l][m]) {
              if (L.C == 4) scatter_vec<2, false, 4>(G.plane[l][m], 4, 0, tb, gb);
              else scatter_vec<2, false, 2>(G.plane[l][m], 2, 0, tb, gb);
            }
          }
        }
      }
      // coefficient lines: 1-D scatter of the 3 x Cc gradient row
#pragma unroll 1
      for (int m = 0; m < 3; ++m) {
        float* gl = G.cline[m];
        if (!gl) continue;
        const LTap t = vm_line_tap(P, m, xr);
        if (P.Cc == 18) {
          float g[18];
#pragma unroll
          for (int c = 0; c < 18; ++c) g[c] = rowA[m * 18 + c];
          if (t.ok0 && t.w0 != 0.0f) red_texel18(gl, (size_t)t.i0, g, t.w0);
          if (t.ok1 && t.w1 != 0.0f) red_texel18(gl, (size_t)(t.i0 + 1), g, t.w1);
        } else {
          for (int c = 0; c < P.Cc; c += 2) {
            const float g0 = rowA[m * P.Cc + c], g1 = rowA[m * P.Cc + c + 1];
            if (t.ok0 && t.w0 != 0.0f) red_add_v2(gl + (size_t)t.i0 * P.Cc + c, g0 * t.w0, g1 * t.w0);
            if (t.ok1 && t.w1 != 0.0f) red_add_v2(gl + (size_t)(t.i0 + 1) * P.Cc + c, g0 * t.w1, g1 * t.w1);
          }
        }
      }
    }
    __syncwarp();
  }
}

static bool build_vm_params(const ffb_field_desc& d, VmParams& P) {
  if (d.xdim != 3 || d.in_dim != 3 || d.mapping == FFB_MAP_TRIG || d.basis_is_x || !d.basis_perm) return false;
  if (d.n_cterms != 3 || d.n_bterms < 3 || d.n_bterms % 3 || d.n_bterms / 3 > VM_MAX_LEVELS) return false;
  if (d.coeff_width <= 0 || d.coeff_width != d.basis_width || d.coeff_width > 96 || (d.coeff_width & 1)) return false;
  int col = 0;
  for (int m = 0; m < 3; ++m) {
    const ffb_term& T = d.cterms[m];
    if (T.n_ops != 1 || T.col != col) return false;
    const ffb_gather_op& o = d.ops[T.op[0]];
    if (o.nd != 2 || o.size[0] != 1 || o.src[0] >= 0 || o.cst[0] != 0.0f || o.src[1] < 0 || o.src[1] > 2) return false;
    if (o.space != 0 || o.align_corners || !o.border || o.nearest || (o.C & 1) || ((uintptr_t)o.data & 15)) return false;
    if (m > 0 && (o.C != P.Cc || o.size[1] != P.Hc)) return false;
    P.Cc = o.C; P.Hc = o.size[1];
    P.cline[m] = o.data; P.caxis[m] = o.src[1]; P.cop[m] = T.op[0];
    col += o.C;
  }
  if (col != d.coeff_width) return false;
  col = 0;
  const int F = d.n_bterms / 3;
  for (int t = 0; t < d.n_bterms; ++t) {
    const ffb_term& T = d.bterms[t];
    const int l = t / 3, m = t % 3;
    if (T.n_ops != 1 || T.col != col) return false;
    const ffb_gather_op& o = d.ops[T.op[0]];
    if (o.nd != 2 || o.space != 1 || o.level != l || !o.align_corners || o.border || o.nearest || ((uintptr_t)o.data & 15)) return false;
    if (o.C != 2 && o.C != 4) return false;
    if (o.src[0] < 0 || o.src[0] > 2 || o.src[1] < 0 || o.src[1] > 2) return false;
    if (l == 0) { P.ax0[m] = o.src[0]; P.ax1[m] = o.src[1]; }
    else if (P.ax0[m] != o.src[0] || P.ax1[m] != o.src[1]) return false;
    VmLevel& L = P.lv[l];
    if (m == 0) { L.C = o.C; L.col = col; L.freq = d.freq[l]; }
    else if (o.C != L.C) return false;
    L.plane[m] = o.data; L.pw[m] = o.size[0]; L.ph[m] = o.size[1]; L.op[m] = T.op[0];
    col += o.C;
  }
  if (col != d.basis_width) return false;
  P.mapping = d.mapping; P.n_levels = F; P.W = d.coeff_width; P.perm = d.basis_perm;
  for (int k = 0; k < 3; ++k) { P.lo[k] = d.aabb_min[k]; P.hi[k] = d.aabb_max[k]; }
  return true;
}

static int g_planes_enabled = 1;
static int g_planes_v2 = 1;        // knob "field_planes_v2": warp-tile kernels (coalesced row traffic); 0: first-generation kernels
static int g_planes_unroll = 1;    // knob "field_planes_unroll": bit 0 forward, bit 1 backward — the three planes of a level in flight together
// measured at the -vm bench shape (423 k queries): forward 275 us (first generation) -> 220 (warp tiles) -> 175 us (+ unrolled planes);
// backward 722 -> 482 us (unrolled: 514 us, left off)

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_field_planes(int enabled) {
  g_planes_enabled = enabled ? 1 : 0;
  return FFB_OK;
}

int ffb_set_field_planes_tuning(int which, int value) {     // reached through ffb_set_tuning("field_planes_v2" | "field_planes_unroll")
  if (which == 0) g_planes_v2 = value;
  else g_planes_unroll = value;
  return FFB_OK;
}

int ffb_field_planes_eligible(ffb_field_t f) {
  if (!f || !g_planes_enabled) return 0;
  VmParams P;
  return build_vm_params(f->h, P) ? 1 : 0;
}

int ffb_field_planes_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, float* feats, float* coeff, float* basis, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  VmParams P;
  FFB_REQUIRE(g_planes_enabled && build_vm_params(f->h, P), "descriptor is not a vector-matrix (vm) field");
  if (n <= 0) return FFB_OK;
  static PerDeviceOnce once;
  if (once.first()) {
    FFB_CUDA(cudaFuncSetAttribute(vm_fwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    FFB_CUDA(cudaFuncSetAttribute(vm_fwd2_kernel<8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    FFB_CUDA(cudaFuncSetAttribute(vm_fwd2_kernel<6, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  const bool rows16 = (((uintptr_t)feats | (uintptr_t)coeff) & 15) == 0;
  if (g_planes_v2 && !basis && rows16) {
    const size_t smem2 = (size_t)(P.W + 1) * VM_NT * sizeof(float);
    FFB_REQUIRE(smem2 <= 64 * 1024, "coefficient row too wide");
    const int blocks = blocks_for(n, VM_NT, (int64_t)sm_count() * 64);
    if (g_planes_unroll & 1) vm_fwd2_kernel<6, true><<<blocks, VM_NT, smem2, (cudaStream_t)stream>>>(P, x, n, n_dev, feats, coeff);
    else vm_fwd2_kernel<8, false><<<blocks, VM_NT, smem2, (cudaStream_t)stream>>>(P, x, n, n_dev, feats, coeff);
    FFB_LAUNCHED();
    return FFB_OK;
  }
  const size_t smem = (size_t)P.W * VM_NT * sizeof(float);
  FFB_REQUIRE(smem <= 64 * 1024, "coefficient row too wide");
  vm_fwd_kernel<6><<<blocks_for(n, VM_NT, (int64_t)sm_count() * 64), VM_NT, smem, (cudaStream_t)stream>>>(P, x, n, n_dev, feats, coeff, basis);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_planes_bwd_saved(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                               const float* coeff, const float* basis, float* const* h_grads, void* stream) {
  FFB_REQUIRE(f && x, "null argument");
  VmParams P;
  FFB_REQUIRE(g_planes_enabled && build_vm_params(f->h, P), "descriptor is not a vector-matrix (vm) field");
  if (n <= 0) return FFB_OK;
  VmGrads G;
  auto grad_of = [&](int op) { return h_grads ? h_grads[op] : f->h.ops[op].grad; };
  bool aligned = true;
  for (int m = 0; m < 3; ++m) {
    G.cline[m] = grad_of(P.cop[m]);
    aligned = aligned && ((uintptr_t)G.cline[m] & 15) == 0;
  }
  for (int l = 0; l < VM_MAX_LEVELS; ++l)
    for (int m = 0; m < 3; ++m) {
      G.plane[l][m] = l < P.n_levels ? grad_of(P.lv[l].op[m]) : nullptr;
      aligned = aligned && ((uintptr_t)G.plane[l][m] & 15) == 0;
    }
  FFB_REQUIRE(aligned, "gradient tensors must be 16-byte aligned");
  const size_t smem = (size_t)P.W * VM_NT * sizeof(float);
  FFB_REQUIRE(smem <= 64 * 1024, "coefficient row too wide");
  static PerDeviceOnce once;
  if (once.first()) {
    FFB_CUDA(cudaFuncSetAttribute(vm_bwd_kernel<6>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    FFB_CUDA(cudaFuncSetAttribute(vm_bwd2_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    FFB_CUDA(cudaFuncSetAttribute(vm_bwd2_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
  }
  const size_t smem2 = (size_t)(P.W + 1) * VM_NT * 2 * sizeof(float);
  if (g_planes_v2 && g_feats && !g_coeff && !basis && smem2 <= 64 * 1024 && (((uintptr_t)g_feats | (uintptr_t)coeff) & 15) == 0) {
    // upstream gradient of the features only (the render / regression paths): warp-tile kernel; coeff = the row the forward returned
    const int blocks = blocks_for(n, VM_NT, (int64_t)sm_count() * 64);
    if (g_planes_unroll & 2) vm_bwd2_kernel<4, true><<<blocks, VM_NT, smem2, (cudaStream_t)stream>>>(P, G, x, n, n_dev, g_feats, coeff);
    else vm_bwd2_kernel<4, false><<<blocks, VM_NT, smem2, (cudaStream_t)stream>>>(P, G, x, n, n_dev, g_feats, coeff);
    FFB_LAUNCHED();
    return FFB_OK;
  }
  if (!basis) coeff = nullptr;      // the first-generation kernel takes both saved rows or neither
  vm_bwd_kernel<6><<<blocks_for(n, VM_NT, (int64_t)sm_count() * 64), VM_NT, smem, (cudaStream_t)stream>>>(P, G, x, n, n_dev, g_feats, g_coeff, coeff, basis);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_field_planes_bwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* g_feats, const float* g_coeff,
                         float* const* h_grads, void* stream) {
  return ffb_field_planes_bwd_saved(f, x, n, n_dev, g_feats, g_coeff, nullptr, nullptr, h_grads, stream);
}

}  // extern "C"
