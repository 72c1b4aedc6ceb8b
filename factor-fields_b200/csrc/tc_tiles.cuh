// tc_tiles.cuh — shared-memory operand tiles of the fused MLP kernels (mlp_fused.cu: linear_mat alone; field_mlp.cu: the
// field gather / scatter fused with linear_mat): staging of fp32 blocks as bf16-split UMMA tiles, descriptors for the
// K-major / MN-major roles of a tile, the split-product MMA issue loop, weight staging and the row loaders.
#pragma once
#include "tc_common.cuh"

namespace ffb {

constexpr uint32_t TILE_SR = 128;   // bytes between 8-row groups inside one 8-column chunk

__device__ __forceinline__ uint64_t desc_k(uint32_t base, uint32_t sc, int kslice) {   // M/N = tile rows, K = tile cols
  return make_desc(base + (uint32_t)kslice * 2u * sc, sc, TILE_SR);
}
__device__ __forceinline__ uint64_t desc_mn(uint32_t base, uint32_t sc, int kslice) {  // M/N = tile cols, K = tile rows
  return make_desc(base + (uint32_t)kslice * 2u * TILE_SR, TILE_SR, sc);
}

// Stage a [rows x 8*nchunks] fp32 block as TERMS bf16 operand tiles (tile column chunk stride sc = tile_rows*16).
// Warp-cooperative: a warp owns an 8-row group at a time; lane = (rr = lane>>2: row in the group, q = lane&3: column
// pair 2q, 2q+1 of an 8-wide chunk), so one store instruction writes 128 contiguous bytes.
template <int TERMS, class Load>
__device__ __forceinline__ void stage_tile(uint8_t* dst, uint32_t part_bytes, uint32_t sc, int rows, int nchunks, int warp, int nwarps,
                                           int lane, Load load) {
  const int rr = lane >> 2, q = lane & 3;
  const int items = (rows >> 3) * nchunks;                 // (row group, column chunk) pairs
  constexpr int BATCH = 8;
  for (int it0 = warp * BATCH; it0 < items; it0 += nwarps * BATCH) {
    float2 v[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int it = it0 + b;
      v[b] = make_float2(0.f, 0.f);
      if (it < items) v[b] = load((it / nchunks) * 8 + rr, (it % nchunks) * 8 + 2 * q);
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const int it = it0 + b;
      if (it < items) {
        uint8_t* p = dst + (uint32_t)(it % nchunks) * sc + (uint32_t)(it / nchunks) * TILE_SR + (uint32_t)rr * 16u + (uint32_t)q * 4u;
        uint32_t w[TERMS];
        split2_packed<TERMS>(v[b].x, v[b].y, w);
#pragma unroll
        for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint32_t*>(p + (uint32_t)t * part_bytes) = w[t];
      }
    }
  }
}

// The same staging split in two halves, so the global loads of the NEXT tile can be in flight while the current tile
// is computed: tile_load fills NB float2 registers per lane, tile_store converts and writes them.
template <int NB, class Load>
__device__ __forceinline__ void tile_load(float2 v[NB], int rows, int nchunks, int warp, int nwarps, int lane, Load load) {
  const int rr = lane >> 2, q = lane & 3, items = (rows >> 3) * nchunks;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int it = warp + nwarps * b;
    v[b] = make_float2(0.f, 0.f);
    if (it < items) v[b] = load((it / nchunks) * 8 + rr, (it % nchunks) * 8 + 2 * q);
  }
}
template <int TERMS, int NB>
__device__ __forceinline__ void tile_store(float2 v[NB], uint8_t* dst, uint32_t part_bytes, uint32_t sc, int rows, int nchunks, int warp,
                                           int nwarps, int lane) {
  const int rr = lane >> 2, q = lane & 3, items = (rows >> 3) * nchunks;
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    const int it = warp + nwarps * b;
    if (it < items) {
      uint8_t* p = dst + (uint32_t)(it % nchunks) * sc + (uint32_t)(it / nchunks) * TILE_SR + (uint32_t)rr * 16u + (uint32_t)q * 4u;
      uint32_t w[TERMS];
      split2_packed<TERMS>(v[b].x, v[b].y, w);
#pragma unroll
      for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint32_t*>(p + (uint32_t)t * part_bytes) = w[t];
    }
  }
}

// thread-per-row store of 8 consecutive columns (c0 % 8 == 0) of an activation tile, truncating split (exact with 3 parts)
template <int TERMS>
__device__ __forceinline__ void store_row8_trunc(uint8_t* dst, uint32_t part_bytes, uint32_t sc, int r, int c0, const float v[8]) {
  uint4 parts[TERMS];
  split8_trunc<TERMS>(v, parts);
  uint8_t* p = dst + (uint32_t)(c0 >> 3) * sc + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u;
#pragma unroll
  for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint4*>(p + (uint32_t)t * part_bytes) = parts[t];
}

// thread-per-row store of 8 consecutive columns (c0 % 8 == 0) of an activation tile
template <int TERMS>
__device__ __forceinline__ void store_row8(uint8_t* dst, uint32_t part_bytes, uint32_t sc, int r, int c0, const float v[8]) {
  uint4 parts[TERMS];
  split8_packed<TERMS>(v, parts);
  uint8_t* p = dst + (uint32_t)(c0 >> 3) * sc + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u;
#pragma unroll
  for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint4*>(p + (uint32_t)t * part_bytes) = parts[t];
}

// D (+)= A * B^T with every kept cross term of the bf16 split: parts (ta, tb) with ta + tb < TERMS
template <int TERMS, class DA, class DB>
__device__ __forceinline__ void issue_gemm(uint32_t tmem_d, uint32_t idesc, int kslices, bool accumulate, DA da, DB db) {
  uint32_t acc = accumulate ? 1u : 0u;
  for (int s = 0; s < kslices; ++s) {
#pragma unroll
    for (int ta = 0; ta < TERMS; ++ta)
#pragma unroll
      for (int tb = 0; tb < TERMS; ++tb) {
        if (ta + tb >= TERMS) continue;
        umma_f16(tmem_d, da(ta, s), db(tb, s), idesc, acc);
        acc = 1u;
      }
  }
}

// The same with every address offset a compile-time constant: operand tiles whose parts are A_PART / B_PART bytes apart and
// whose K = 16 slices are A_KSTEP / B_KSTEP bytes apart.  FIRST_ACC = false overwrites the accumulator with the first MMA.
template <int TERMS, int KSLICES, uint32_t A_PART, uint32_t A_KSTEP, uint32_t B_PART, uint32_t B_KSTEP, bool FIRST_ACC>
__device__ __forceinline__ void issue_gemm_c(uint32_t tmem_d, uint32_t idesc, const DescBase a, const DescBase b) {
#pragma unroll
  for (int s = 0; s < KSLICES; ++s) {
#pragma unroll
    for (int ta = 0; ta < TERMS; ++ta)
#pragma unroll
      for (int tb = 0; tb < TERMS; ++tb) {
        if (ta + tb >= TERMS) continue;
        const uint64_t da = desc_at(a, (uint32_t)ta * A_PART + (uint32_t)s * A_KSTEP), db = desc_at(b, (uint32_t)tb * B_PART + (uint32_t)s * B_KSTEP);
        if (s == 0 && ta == 0 && tb == 0 && !FIRST_ACC) umma_f16_c<false>(tmem_d, da, db, idesc);
        else umma_f16_c<true>(tmem_d, da, db, idesc);
      }
  }
}

struct Mlp2Shape {
  int K0, H, N;       // logical sizes
  int K0p, Np;        // padded to multiples of 16 (K0p includes the bias column)
};

// weights -> operand tiles (once per CTA).  W1 tile: rows j < H, cols k < K0p (col K0 = b1[j]);  W2 tile: rows n < Np, cols j < H
template <int TERMS>
__device__ __forceinline__ void stage_weights(const Mlp2Shape S, const float* __restrict__ W1, const float* __restrict__ b1,
                                              const float* __restrict__ W2, uint8_t* sW1, uint8_t* sW2, int tid, int nthreads) {
  const uint32_t szW1 = (uint32_t)S.H * S.K0p * 2, szW2 = (uint32_t)S.Np * S.H * 2;
  const uint32_t scW1 = (uint32_t)S.H * 16, scW2 = (uint32_t)S.Np * 16;
  for (int item = tid; item < S.H * (S.K0p / 8); item += nthreads) {
    const int j = item % S.H, c = item / S.H;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int k = c * 8 + i;
      v[i] = k < S.K0 ? __ldg(W1 + (int64_t)j * S.K0 + k) : ((k == S.K0 && b1) ? __ldg(b1 + j) : 0.0f);
    }
    store_row8<TERMS>(sW1, szW1, scW1, j, c * 8, v);
  }
  for (int item = tid; item < S.Np * (S.H / 8); item += nthreads) {
    const int nn = item % S.Np, c = item / S.Np;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = nn < S.N ? __ldg(W2 + (int64_t)nn * S.H + c * 8 + i) : 0.0f;
    store_row8<TERMS>(sW2, szW2, scW2, nn, c * 8, v);
  }
}

// x rows with the all-ones bias column at index K0 (rows beyond n are zero, including that column)
struct XLoader {
  const float* x;
  int64_t row0, n;
  int K0;
  bool vec2;
  __device__ __forceinline__ float2 operator()(int r, int c) const {
    const int64_t row = row0 + r;
    float2 v = make_float2(0.f, 0.f);
    if (row < n) {
      if (c + 1 < K0) {
        if (vec2) v = *reinterpret_cast<const float2*>(x + row * K0 + c);
        else { v.x = x[row * K0 + c]; v.y = x[row * K0 + c + 1]; }
      } else {
        v.x = c < K0 ? x[row * K0 + c] : (c == K0 ? 1.0f : 0.0f);
        v.y = c + 1 < K0 ? x[row * K0 + c + 1] : (c + 1 == K0 ? 1.0f : 0.0f);
      }
    }
    return v;
  }
};
__device__ __forceinline__ XLoader x_loader(const Mlp2Shape& S, const float* x, int64_t row0, int64_t n) {
  return XLoader{x, row0, n, S.K0, ((S.K0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0)};
}
// generic row-major [n, N] block (columns beyond N and rows beyond n read as zero)
struct RowLoader {
  const float* g;
  int64_t row0, n;
  int N;
  bool vec2;
  __device__ __forceinline__ float2 operator()(int r, int c) const {
    const int64_t row = row0 + r;
    float2 v = make_float2(0.f, 0.f);
    if (row < n && c < N) {
      if (vec2 && c + 1 < N) v = *reinterpret_cast<const float2*>(g + row * N + c);
      else { v.x = g[row * N + c]; if (c + 1 < N) v.y = g[row * N + c + 1]; }
    }
    return v;
  }
};

}  // namespace ffb
