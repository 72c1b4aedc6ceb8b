// mlp_pipe.cu — `linear_mat` (MLPMixer 2 layers, FactorFields.py:144-159) as PIPELINED warp-specialised tcgen05 kernels.
//
// mlp_fused.cu runs one 128-row tile per CTA as a serial chain (stage -> MMA -> wait -> epilogue -> MMA -> wait -> store) with
// every thread of the CTA spinning on each MMA round trip: ncu shows 2 900 (forward) / 4 800 (backward) thread-instructions
// per row, two thirds of them in the wait loops, tensor pipe 13 % busy.  Here one persistent CTA per SM splits the roles:
//   * PRODUCER warps keep a ring of operand-tile slots full (rows prefetched into registers before the slot is free, bf16-split
//     into the canonical no-swizzle UMMA layout);
//   * ONE thread issues the first-stage MMAs of tile t+1, t+2 while earlier tiles are still in their epilogues (accumulators
//     double-buffered in TMEM), tcgen05.commit frees the slot / publishes the accumulator through mbarriers;
//   * EPILOGUE warps (one per TMEM lane quarter and column half) do the ReLU / masking / re-splitting and the output stores,
//     and issue the second-stage MMAs themselves — the only threads that ever wait on an MMA.
// Same arithmetic as mlp_fused.cu (3 bf16 parts forward, 2 parts for gradients, ReLU decisions taken once in the forward pass,
// layer-1 bias as an all-ones input column, weight gradients accumulated in TMEM across all tiles of the CTA).
#include "tc_tiles.cuh"
#include "ffb_math.h"

namespace ffb {

constexpr int MP_H = 64, MP_K0P = 32, MP_NP = 32;

// Optional timeline trace (experiment knob ffb_mlp2p_trace): CTA 0 appends (event id, tile, globaltimer ns) triples.
__device__ long long* g_mp_trace = nullptr;
__device__ int g_mp_trace_n = 0;
__device__ __forceinline__ void mp_trace(int ev, long long tile) {
  long long* buf = g_mp_trace;
  if (buf && blockIdx.x == 0) {
    const int k = atomicAdd(&g_mp_trace_n, 1);
    if (k < 20000) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      buf[3 * k] = ev; buf[3 * k + 1] = tile; buf[3 * k + 2] = (long long)t;
    }
  }
}
constexpr uint32_t MP_SC = 2048;               // bytes between 8-column chunks of a 128-row tile

template <int TERMS>
__device__ __forceinline__ void mp_store2(uint8_t* row, uint32_t part_bytes, int c, float a, float b) {
  uint32_t w[TERMS];
  split2_packed<TERMS>(a, b, w);
  uint8_t* p = row + (uint32_t)(c >> 3) * MP_SC + (uint32_t)(c & 7) * 2u;
#pragma unroll
  for (int t = 0; t < TERMS; ++t) *reinterpret_cast<uint32_t*>(p + (uint32_t)t * part_bytes) = w[t];
}

// ---------------------------------------------------------------------------------------------------------
// forward:  y = relu([x | 1] [W1 | b1]^T) W2^T
// warps: 0..6 producers, 7 layer-1 issuer, 8.. epilogue — group = (warp - 8) >> 3 takes tiles t = group mod MPF_NGROUPS; inside a group
// q = warp & 3 is the TMEM lane quarter and half = ((warp - 8) >> 2) & 1 the column half.  (The SM's warp arbiter favours
// the highest warp ids: the latency-critical epilogue warps sit there, the producers — who mostly wait for a free slot — below.)
// ---------------------------------------------------------------------------------------------------------
constexpr int MPF_NGROUPS = 3;      // epilogue groups (tiles in the epilogue stage at once); each owns an accumulator pair and a hidden tile
constexpr int MPF_THREADS = (8 + 8 * MPF_NGROUPS) * 32, MPF_NPROD = 7, MPF_ISSUER = 7, MPF_EPI0 = 8, MPF_NSLOT = 2, MPF_TERMS = 3;
constexpr uint32_t MPF_TMEM_COLS = 512;      // NGROUPS x (64 + 32) accumulator columns
constexpr uint32_t MPF_XPART = 4 * MP_SC, MPF_XSLOT = MPF_TERMS * MPF_XPART;      // 8 KB / 24 KB
constexpr uint32_t MPF_HPART = 8 * MP_SC, MPF_HBUF = MPF_TERMS * MPF_HPART;       // 16 KB / 48 KB
struct MpfSmem {
  static constexpr uint32_t W1 = 0;
  static constexpr uint32_t W2 = W1 + MPF_TERMS * MP_H * MP_K0P * 2;
  static constexpr uint32_t X = W2 + MPF_TERMS * MP_NP * MP_H * 2;
  static constexpr uint32_t HID = X + MPF_NSLOT * MPF_XSLOT;
  static constexpr uint32_t BAR = HID + MPF_NGROUPS * MPF_HBUF;
  static constexpr uint32_t N_BAR = 2 * MPF_NSLOT + 3 * MPF_NGROUPS;     // slot_full, slot_free, d1_full[G], d1_free[G], d2_full[G]
  static constexpr uint32_t MISC = BAR + N_BAR * 8;
  static constexpr uint32_t TOTAL = MISC + 16;
};

__global__ void __launch_bounds__(MPF_THREADS, 1) mlp2p_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W1,
                                                                   const float* __restrict__ b1, const float* __restrict__ W2,
                                                                   float* __restrict__ y, uint16_t* __restrict__ mask, int64_t n_cap,
                                                                   const int32_t* __restrict__ n_dev, const Mlp2Shape S) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sW1 = smem + MpfSmem::W1;
  uint8_t* sW2 = smem + MpfSmem::W2;
  uint8_t* sX = smem + MpfSmem::X;
  uint8_t* sH = smem + MpfSmem::HID;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MpfSmem::BAR);
  uint64_t* slot_full = bars;
  uint64_t* slot_free = bars + MPF_NSLOT;
  uint64_t* d1_full = bars + 2 * MPF_NSLOT;
  uint64_t* d1_free = d1_full + MPF_NGROUPS;
  uint64_t* d2_full = d1_free + MPF_NGROUPS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + MpfSmem::MISC);
  int* next_chunk = reinterpret_cast<int*>(smem + MpfSmem::MISC + 4);
  const int K0 = S.K0, N = S.N;

  if (warp == MPF_ISSUER) tmem_alloc(tmem_slot, MPF_TMEM_COLS);
  if (tid == 0) {
    for (int s = 0; s < MPF_NSLOT; ++s) {
      mbar_init(slot_full + s, 4);
      mbar_init(slot_free + s, 1);
    }
    for (int b = 0; b < MPF_NGROUPS; ++b) {
      mbar_init(d1_full + b, 1);
      mbar_init(d1_free + b, 256);
      mbar_init(d2_full + b, 1);
    }
    *next_chunk = 0;
  }
  stage_weights<MPF_TERMS>(S, W1, b1, W2, sW1, sW2, tid, MPF_THREADS);
  for (uint32_t o = tid * 16u; o < MPF_NSLOT * MPF_XSLOT; o += MPF_THREADS * 16u) *reinterpret_cast<uint4*>(sX + o) = make_uint4(0, 0, 0, 0);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int64_t n_tiles = (n + 127) >> 7;
  const int64_t Tc = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

  if (warp < MPF_NPROD) {
    // ------------------------------ producers: one 32-row chunk at a time, rows prefetched before the slot wait
    const int n_chunks = (int)(4 * Tc);
    const bool vec2 = ((K0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    for (;;) {
      int j = 0;
      if (lane == 0) j = atomicAdd(next_chunk, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= n_chunks) break;
      const int tseq = j >> 2, rg = j & 3, slot = tseq % MPF_NSLOT;
      const int64_t i = (((int64_t)blockIdx.x + (int64_t)tseq * gridDim.x) << 7) + rg * 32 + lane;
      float v[MP_K0P];
#pragma unroll
      for (int c = 0; c < MP_K0P; ++c) v[c] = 0.0f;
      if (i < n) {
        const float* xr = x + i * K0;
        if (vec2) {
#pragma unroll
          for (int c = 0; c < MP_K0P; c += 2)
            if (c < K0) {
              const float2 t2 = *reinterpret_cast<const float2*>(xr + c);
              v[c] = t2.x;
              v[c + 1] = t2.y;
            }
        } else {
#pragma unroll
          for (int c = 0; c < MP_K0P; ++c)
            if (c < K0) v[c] = xr[c];
        }
#pragma unroll
        for (int c = 0; c < MP_K0P; ++c)
          if (c == K0) v[c] = 1.0f;                 // bias column
      }
      if (lane == 0) mp_trace(10 + rg, tseq);      // rows loaded
      mbar_wait(slot_free + slot, (uint32_t)(((tseq / MPF_NSLOT) & 1) ^ 1));
      if (lane == 0) mp_trace(20 + rg, tseq);      // slot free
      const int r = rg * 32 + lane;
      uint8_t* xrow = sX + (uint32_t)slot * MPF_XSLOT + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u;
#pragma unroll
      for (int c0 = 0; c0 < MP_K0P; c0 += 8) {
        if (c0 <= K0) {                              // chunks beyond the bias column stay zero (cleared once at start)
          uint4 parts[MPF_TERMS];
          split8_packed<MPF_TERMS>(v + c0, parts);
#pragma unroll
          for (int t = 0; t < MPF_TERMS; ++t) *reinterpret_cast<uint4*>(xrow + (uint32_t)(c0 >> 3) * MP_SC + (uint32_t)t * MPF_XPART) = parts[t];
        }
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_full + slot);
      if (lane == 0) mp_trace(30 + rg, tseq);      // chunk delivered
    }
  } else if (warp == MPF_ISSUER) {
    // ------------------------------ layer-1 issuer
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc(MP_H, 0, 0);
      const DescBase dW1b = desc_base(smem_u32(sW1), MP_H * 16, TILE_SR);
      for (int64_t t = 0; t < Tc; ++t) {
        const int slot = (int)(t % MPF_NSLOT), b = (int)(t % MPF_NGROUPS);
        mbar_wait(slot_full + slot, (uint32_t)((t / MPF_NSLOT) & 1));
        mp_trace(40, t);                              // slot full
        mbar_wait(d1_free + b, (uint32_t)(((t / MPF_NGROUPS) & 1) ^ 1));
        mp_trace(41, t);                              // accumulator free
        tc_fence_after();
        const uint32_t aX = smem_u32(sX) + (uint32_t)slot * MPF_XSLOT;
        // A = x tile (K-major: K slice = 2 chunks), B = [W1|b1] tile (K-major, chunk stride H*16)
        issue_gemm_c<MPF_TERMS, MP_K0P / 16, MPF_XPART, 2 * MP_SC, MP_H * MP_K0P * 2, 2 * MP_H * 16, false>(
            tmem + (uint32_t)b * MP_H, idesc1, desc_base(aX, MP_SC, TILE_SR), dW1b);
        umma_commit(slot_free + slot);
        umma_commit(d1_full + b);
        mp_trace(42, t);                              // layer 1 issued
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue: group g handles tiles g, g+2, ...; a thread owns one row and one column half
    const int ew = warp - MPF_EPI0, g = ew >> 3, half = (ew >> 2) & 1, q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t d1 = tmem + (uint32_t)g * MP_H, d2 = tmem + (uint32_t)MPF_NGROUPS * MP_H + (uint32_t)g * MP_NP;
    const uint32_t idesc2 = make_idesc(MP_NP, 0, 0);
    uint8_t* myH = sH + (uint32_t)g * MPF_HBUF;
    const DescBase dHb = desc_base(smem_u32(myH), MP_SC, TILE_SR), dW2b = desc_base(smem_u32(sW2), MP_NP * 16, TILE_SR);
    const bool tr = half == 0 && q == 0 && lane == 0;
    auto epi2 = [&](int64_t t, int64_t k) {
      if (tr) mp_trace(50, t);
      mbar_wait(d2_full + g, (uint32_t)(k & 1));
      if (tr) mp_trace(51, t);                        // layer 2 complete
      tc_fence_after();
      const int64_t grow = (((int64_t)blockIdx.x + t * gridDim.x) << 7) + row;
      const int c0 = half * 16;
      float v[16];
      tmem_ld16(d2 + lane_base + (uint32_t)c0, v);
      if (grow < n) {
        float* yr = y + grow * N + c0;
        if ((N & 3) == 0 && c0 + 16 <= N) {
#pragma unroll
          for (int k2 = 0; k2 < 16; k2 += 4) *reinterpret_cast<float4*>(yr + k2) = make_float4(v[k2], v[k2 + 1], v[k2 + 2], v[k2 + 3]);
        } else {
#pragma unroll
          for (int k2 = 0; k2 < 16; ++k2)
            if (c0 + k2 < N) yr[k2] = v[k2];
        }
      }
    };
    int64_t k = 0;
    for (int64_t t = g; t < Tc; t += MPF_NGROUPS, ++k) {
      if (k > 0) epi2(t - MPF_NGROUPS, k - 1);          // also: layer 2 of this group's previous tile has finished reading the hidden tile
      if (tr) mp_trace(52, t);                        // epilogue 2 done
      mbar_wait(d1_full + g, (uint32_t)(k & 1));
      if (tr) mp_trace(53, t);                        // layer 1 complete
      tc_fence_after();
      const int64_t grow = (((int64_t)blockIdx.x + t * gridDim.x) << 7) + row;
#pragma unroll
      for (int c0 = half * 32; c0 < half * 32 + 32; c0 += 16) {
        float v[16];
        tmem_ld16(d1 + lane_base + (uint32_t)c0, v);
        uint32_t bits = 0;
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          bits |= (v[i2] > 0.0f ? 1u : 0u) << i2;
          v[i2] = fmaxf(v[i2], 0.0f);
        }
        if (mask && grow < n) mask[grow * (MP_H >> 4) + (c0 >> 4)] = (uint16_t)bits;
        store_row8_trunc<MPF_TERMS>(myH, MPF_HPART, MP_SC, row, c0, v);
        store_row8_trunc<MPF_TERMS>(myH, MPF_HPART, MP_SC, row, c0 + 8, v + 8);
      }
      tc_fence_before();
      mbar_arrive(d1_free + g);
      proxy_fence();
      if (tr) mp_trace(54, t);                        // epilogue 1 done (this thread)
      named_sync(1 + g, 256);
      if (half == 0 && q == 0 && elect_one()) {
        mp_trace(55, t);                              // group synchronised
        tc_fence_after();
        issue_gemm_c<MPF_TERMS, MP_H / 16, MPF_HPART, 2 * MP_SC, MP_NP * MP_H * 2, 2 * MP_NP * 16, false>(d2, idesc2, dHb, dW2b);
        umma_commit(d2_full + g);
      }
    }
    if (k > 0) epi2(g + MPF_NGROUPS * (k - 1), k - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MPF_ISSUER) tmem_dealloc(tmem, MPF_TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------
// backward.  Tiles (2 bf16 parts, 128 rows):  B2 = [ g_y | x 1 ]  (64 columns, ring of 3),  A2 = [ h | g_h ]  (128 columns).
//   stage 1 (issuer warp, runs ahead):  D1 = X [W1|b1]^T   D2 = GY W2                       (double-buffered in TMEM)
//   epilogue 1: ReLU decisions (forward's bits) -> A2
//   stage 2 (issued by the epilogue):   D3 = GH W1 (g_x)   DW += A2^T B2  (gW2^T and [gW1|gb1], resident in TMEM)
//   epilogue 2 (of the PREVIOUS tile, under stage 2 of this one): g_x tile -> fp32 staging -> coalesced stores
// warps: 0..6 producers, 7 stage-1 issuer, 8..23 epilogue (q = warp & 3 lane quarter, cq = (warp - 8) >> 2 column quarter)
// ---------------------------------------------------------------------------------------------------------
constexpr int MPB_THREADS = 24 * 32, MPB_NPROD = 7, MPB_ISSUER = 7, MPB_EPI0 = 8, MPB_NEPI = 512, MPB_NSLOT = 3, MPB_TERMS = 2;
constexpr uint32_t MPB_BPART = 8 * MP_SC, MPB_BSLOT = MPB_TERMS * MPB_BPART;      // 16 KB / 32 KB
constexpr uint32_t MPB_APART = 16 * MP_SC, MPB_ABUF = MPB_TERMS * MPB_APART;      // 32 KB / 64 KB
struct MpbSmem {
  static constexpr uint32_t W1 = 0;
  static constexpr uint32_t W2 = W1 + MPB_TERMS * MP_H * MP_K0P * 2;
  static constexpr uint32_t B2 = W2 + MPB_TERMS * MP_NP * MP_H * 2;
  static constexpr uint32_t A2 = B2 + MPB_NSLOT * MPB_BSLOT;
  static constexpr uint32_t STG = A2 + MPB_ABUF;                    // g_x staging: 128 x 32 floats
  static constexpr uint32_t BAR = STG + 128 * MP_K0P * 4;
  static constexpr uint32_t N_BAR = 2 * MPB_NSLOT + 2 + 2 + 1;      // slot_full, slot_free, d12_full[2], d12_free[2], cd_full
  static constexpr uint32_t MISC = BAR + N_BAR * 8;
  static constexpr uint32_t TOTAL = MISC + 16;
};

__global__ void __launch_bounds__(MPB_THREADS, 1) mlp2p_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                   const float* __restrict__ W1, const float* __restrict__ b1,
                                                                   const float* __restrict__ W2, const uint16_t* __restrict__ mask,
                                                                   float* __restrict__ gx, float* __restrict__ gW1, float* __restrict__ gb1,
                                                                   float* __restrict__ gW2, int64_t n_cap, const int32_t* __restrict__ n_dev,
                                                                   const Mlp2Shape S, const float* __restrict__ g0 = nullptr,
                                                                   const int32_t* __restrict__ row_slot = nullptr,
                                                                   const float* __restrict__ g_rows = nullptr) {
  // g0 != nullptr (SPARSE upstream gradient, the render path): column 0 of row i is g0[i]; columns 1.. are row row_slot[i] of the
  // compact g_rows [*, N] when row_slot[i] >= 0 and zero otherwise (85 % of the samples at nerf.yaml) — gy is not read
  extern __shared__ __align__(128) uint8_t smem[];
  const int64_t n = resolve_n(n_cap, n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sW1 = smem + MpbSmem::W1;
  uint8_t* sW2 = smem + MpbSmem::W2;
  uint8_t* sB2 = smem + MpbSmem::B2;
  uint8_t* sA2 = smem + MpbSmem::A2;
  float* stg = reinterpret_cast<float*>(smem + MpbSmem::STG);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + MpbSmem::BAR);
  uint64_t* slot_full = bars;
  uint64_t* slot_free = bars + MPB_NSLOT;
  uint64_t* d12_full = bars + 2 * MPB_NSLOT;
  uint64_t* d12_free = d12_full + 2;
  uint64_t* cd_full = d12_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + MpbSmem::MISC);
  int* next_chunk = reinterpret_cast<int*>(smem + MpbSmem::MISC + 4);
  const int K0 = S.K0, N = S.N, H = MP_H;
  constexpr uint32_t offX = (uint32_t)(MP_NP / 8) * MP_SC, offGH = (uint32_t)(MP_H / 8) * MP_SC;      // x inside B2, g_h inside A2
  constexpr int NB2 = MP_NP + MP_K0P;

  if (warp == MPB_ISSUER) tmem_alloc(tmem_slot, 512u);
  if (tid == 0) {
    for (int s = 0; s < MPB_NSLOT; ++s) {
      mbar_init(slot_full + s, 4);
      mbar_init(slot_free + s, 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d12_full + b, 1);
      mbar_init(d12_free + b, MPB_NEPI);
    }
    mbar_init(cd_full, 1);
    *next_chunk = 0;
  }
  stage_weights<MPB_TERMS>(S, W1, b1, W2, sW1, sW2, tid, MPB_THREADS);
  for (uint32_t o = tid * 16u; o < MPB_NSLOT * MPB_BSLOT; o += MPB_THREADS * 16u) *reinterpret_cast<uint4*>(sB2 + o) = make_uint4(0, 0, 0, 0);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  // TMEM columns: D1[2] 0..127, D2[2] 128..255, D3[2] 256..319, DW 320..383
  const int64_t n_tiles = (n + 127) >> 7;
  const int64_t Tc = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t szW1 = (uint32_t)H * MP_K0P * 2, szW2 = (uint32_t)MP_NP * H * 2, scW1 = (uint32_t)H * 16, scW2 = (uint32_t)MP_NP * 16;
  const uint32_t aW1 = smem_u32(sW1), aW2 = smem_u32(sW2), aA2 = smem_u32(sA2);

  if (warp < MPB_NPROD) {
    // ------------------------------ producers
    const int n_chunks = (int)(4 * Tc);
    const bool vecX = ((K0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(x) & 7) == 0);
    const bool vecG = ((N & 3) == 0) && ((reinterpret_cast<uintptr_t>(g0 ? g_rows : gy) & 15) == 0);
    for (;;) {
      int j = 0;
      if (lane == 0) j = atomicAdd(next_chunk, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= n_chunks) break;
      const int tseq = j >> 2, rg = j & 3, slot = tseq % MPB_NSLOT;
      const int64_t i = (((int64_t)blockIdx.x + (int64_t)tseq * gridDim.x) << 7) + rg * 32 + lane;
      float v[MP_K0P], gv[MP_NP];
#pragma unroll
      for (int c = 0; c < MP_K0P; ++c) v[c] = 0.0f;
#pragma unroll
      for (int c = 0; c < MP_NP; ++c) gv[c] = 0.0f;
      if (i < n) {
        const float* xr = x + i * K0;
        const float* gr = gy + i * N;
        if (g0) {
          const int sl = row_slot[i];
          if (sl >= 0) {
            const float* sr = g_rows + (int64_t)sl * N;
            if (vecG) {
#pragma unroll
              for (int c = 0; c < MP_NP; c += 4)
                if (c < N) {
                  const float4 t4 = *reinterpret_cast<const float4*>(sr + c);
                  gv[c] = t4.x; gv[c + 1] = t4.y; gv[c + 2] = t4.z; gv[c + 3] = t4.w;
                }
            } else {
#pragma unroll
              for (int c = 0; c < MP_NP; ++c)
                if (c < N) gv[c] = sr[c];
            }
          }
          gv[0] = g0[i];
        } else if (vecG) {
#pragma unroll
          for (int c = 0; c < MP_NP; c += 4)
            if (c < N) {
              const float4 t4 = *reinterpret_cast<const float4*>(gr + c);
              gv[c] = t4.x; gv[c + 1] = t4.y; gv[c + 2] = t4.z; gv[c + 3] = t4.w;
            }
        } else {
#pragma unroll
          for (int c = 0; c < MP_NP; ++c)
            if (c < N) gv[c] = gr[c];
        }
        if (vecX) {
#pragma unroll
          for (int c = 0; c < MP_K0P; c += 2)
            if (c < K0) {
              const float2 t2 = *reinterpret_cast<const float2*>(xr + c);
              v[c] = t2.x;
              v[c + 1] = t2.y;
            }
        } else {
#pragma unroll
          for (int c = 0; c < MP_K0P; ++c)
            if (c < K0) v[c] = xr[c];
        }
#pragma unroll
        for (int c = 0; c < MP_K0P; ++c)
          if (c == K0) v[c] = 1.0f;
      }
      mbar_wait(slot_free + slot, (uint32_t)(((tseq / MPB_NSLOT) & 1) ^ 1));
      const int r = rg * 32 + lane;
      uint8_t* brow = sB2 + (uint32_t)slot * MPB_BSLOT + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u;
#pragma unroll
      for (int c0 = 0; c0 < MP_NP; c0 += 8) {
        uint4 parts[MPB_TERMS];
        split8_packed<MPB_TERMS>(gv + c0, parts);
#pragma unroll
        for (int t = 0; t < MPB_TERMS; ++t) *reinterpret_cast<uint4*>(brow + (uint32_t)(c0 >> 3) * MP_SC + (uint32_t)t * MPB_BPART) = parts[t];
      }
#pragma unroll
      for (int c0 = 0; c0 < MP_K0P; c0 += 8) {
        uint4 parts[MPB_TERMS];
        split8_packed<MPB_TERMS>(v + c0, parts);
#pragma unroll
        for (int t = 0; t < MPB_TERMS; ++t) *reinterpret_cast<uint4*>(brow + offX + (uint32_t)(c0 >> 3) * MP_SC + (uint32_t)t * MPB_BPART) = parts[t];
      }
      proxy_fence();
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_full + slot);
    }
  } else if (warp == MPB_ISSUER) {
    // ------------------------------ stage-1 issuer: D1 = X [W1|b1]^T (hidden pre-activation), D2 = GY W2
    if (elect_one()) {
      const uint32_t id_h = make_idesc(H, 0, 0), id_gh = make_idesc(H, 0, 1);
      const DescBase dW1k = desc_base(aW1, MP_H * 16, TILE_SR), dW2mn = desc_base(aW2, TILE_SR, MP_NP * 16);
      for (int64_t t = 0; t < Tc; ++t) {
        const int slot = (int)(t % MPB_NSLOT), b = (int)(t & 1);
        mbar_wait(slot_full + slot, (uint32_t)((t / MPB_NSLOT) & 1));
        mbar_wait(d12_free + b, (uint32_t)(((t >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t aB2 = smem_u32(sB2) + (uint32_t)slot * MPB_BSLOT;
        issue_gemm_c<MPB_TERMS, MP_K0P / 16, MPB_BPART, 2 * MP_SC, MP_H * MP_K0P * 2, 2 * MP_H * 16, false>(
            tmem + (uint32_t)b * H, id_h, desc_base(aB2 + offX, MP_SC, TILE_SR), dW1k);                       // D1 = X [W1|b1]^T
        issue_gemm_c<MPB_TERMS, MP_NP / 16, MPB_BPART, 2 * MP_SC, MP_NP * MP_H * 2, 2 * TILE_SR, false>(
            tmem + 128u + (uint32_t)b * H, id_gh, desc_base(aB2, MP_SC, TILE_SR), dW2mn);                     // D2 = GY W2 (W2 tile read MN-major)
        umma_commit(d12_full + b);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------ epilogue (512 threads): row = lane quarter * 32 + lane, 16 hidden columns of quarter cq
    const int q = warp & 3, cq = (warp - MPB_EPI0) >> 2, row = q * 32 + lane, etid = tid - MPB_EPI0 * 32;          // etid in [0, 512)
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t dW = tmem + 320u;
    const uint32_t id_gx = make_idesc(MP_K0P, 0, 1), id_w = make_idesc(NB2, 1, 1, 128);
    const DescBase dGHk = desc_base(aA2 + offGH, MP_SC, TILE_SR), dW1mn = desc_base(aW1, TILE_SR, MP_H * 16), dA2mn = desc_base(aA2, TILE_SR, MP_SC);
    auto epi2 = [&](int64_t t) {       // g_x rows of tile t: D3[t & 1] -> staging -> coalesced 16-byte stores
      const uint32_t d3 = tmem + 256u + (uint32_t)(t & 1) * MP_K0P;
      const int64_t row0 = ((int64_t)blockIdx.x + t * gridDim.x) << 7;
      if (gx && cq < 2) {
        float v[16];
        tmem_ld16(d3 + lane_base + (uint32_t)(cq * 16), v);
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2)
          if (cq * 16 + i2 < K0) stg[row * K0 + cq * 16 + i2] = v[i2];
      }
      tc_fence_before();
      named_sync(2, MPB_NEPI);
      if (gx) {
        const int64_t rows = (n - row0) < 128 ? (n - row0) : 128;
        const int64_t total = rows > 0 ? rows * K0 : 0;
        float* dst = gx + row0 * K0;
        if (((row0 * K0) & 3) == 0 && (reinterpret_cast<uintptr_t>(gx) & 15) == 0) {
          const int64_t n4 = total >> 2;
          for (int64_t e = etid; e < n4; e += MPB_NEPI) reinterpret_cast<float4*>(dst)[e] = reinterpret_cast<const float4*>(stg)[e];
          for (int64_t e = (n4 << 2) + etid; e < total; e += MPB_NEPI) dst[e] = stg[e];
        } else {
          for (int64_t e = etid; e < total; e += MPB_NEPI) dst[e] = stg[e];
        }
      }
      named_sync(2, MPB_NEPI);         // staging free again
    };
    for (int64_t t = 0; t < Tc; ++t) {
      const int slot = (int)(t % MPB_NSLOT), b = (int)(t & 1);
      if (t > 0) {                     // stage 2 of tile t-1 complete: A2 free, D3[(t-1)&1] ready, its B2 slot released
        mbar_wait(cd_full, (uint32_t)((t - 1) & 1));
        tc_fence_after();
      }
      mbar_wait(d12_full + b, (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const int64_t grow = (((int64_t)blockIdx.x + t * gridDim.x) << 7) + row;
      const uint32_t d1 = tmem + (uint32_t)b * H, d2 = tmem + 128u + (uint32_t)b * H;
      {
        const int c0 = cq * 16;
        float h[16], g[16];
        tmem_ld16(d1 + lane_base + (uint32_t)c0, h);
        tmem_ld16(d2 + lane_base + (uint32_t)c0, g);
        uint32_t bits;
        if (mask) {
          bits = grow < n ? (uint32_t)mask[grow * (H >> 4) + (c0 >> 4)] : 0u;
        } else {
          bits = 0;
#pragma unroll
          for (int i2 = 0; i2 < 16; ++i2) bits |= (h[i2] > 0.0f ? 1u : 0u) << i2;
          if (grow >= n) bits = 0;
        }
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const bool on = (bits >> i2) & 1u;
          g[i2] = on ? g[i2] : 0.0f;
          h[i2] = on ? fmaxf(h[i2], 0.0f) : 0.0f;
        }
        store_row8<MPB_TERMS>(sA2, MPB_APART, MP_SC, row, c0, h);
        store_row8<MPB_TERMS>(sA2, MPB_APART, MP_SC, row, c0 + 8, h + 8);
        store_row8<MPB_TERMS>(sA2 + offGH, MPB_APART, MP_SC, row, c0, g);
        store_row8<MPB_TERMS>(sA2 + offGH, MPB_APART, MP_SC, row, c0 + 8, g + 8);
      }
      tc_fence_before();
      mbar_arrive(d12_free + b);
      proxy_fence();
      named_sync(1, MPB_NEPI);
      if (warp == MPB_EPI0 && elect_one()) {
        tc_fence_after();
        const uint32_t aB2 = smem_u32(sB2) + (uint32_t)slot * MPB_BSLOT;
        issue_gemm_c<MPB_TERMS, MP_H / 16, MPB_APART, 2 * MP_SC, MP_H * MP_K0P * 2, 2 * TILE_SR, false>(
            tmem + 256u + (uint32_t)b * MP_K0P, id_gx, dGHk, dW1mn);                                          // D3 = GH W1 (W1 tile read MN-major)
        const DescBase dB2mn = desc_base(aB2, TILE_SR, MP_SC);
        if (t > 0) issue_gemm_c<MPB_TERMS, 8, MPB_APART, 2 * TILE_SR, MPB_BPART, 2 * TILE_SR, true>(dW, id_w, dA2mn, dB2mn);   // DW += A2^T B2
        else issue_gemm_c<MPB_TERMS, 8, MPB_APART, 2 * TILE_SR, MPB_BPART, 2 * TILE_SR, false>(dW, id_w, dA2mn, dB2mn);
        umma_commit(slot_free + slot);
        umma_commit(cd_full);
      }
      if (t > 0) epi2(t - 1);          // under stage 2 of tile t
    }
    if (Tc > 0) {
      mbar_wait(cd_full, (uint32_t)((Tc - 1) & 1));
      tc_fence_after();
      epi2(Tc - 1);
      // ---- flush the weight gradients: accumulator row m (M = 128): m < 64 -> gW2^T row j = m; m >= 64 -> [gW1 | gb1] row j = m - 64
      const int m = row;
      {
        const int c0 = cq * 16;        // NB2 = 64 accumulator columns: one 16-column group per column quarter
        float v[16];
        tmem_ld16(dW + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int i2 = 0; i2 < 16; ++i2) {
          const int c = c0 + i2;
          if (v[i2] == 0.0f) continue;
          if (m < H) {
            if (c < N && gW2) atomicAdd(gW2 + (int64_t)c * H + m, v[i2]);
          } else if (c >= MP_NP) {
            const int kk = c - MP_NP, jj = m - H;
            if (kk < K0) { if (gW1) atomicAdd(gW1 + (int64_t)jj * K0 + kk, v[i2]); }
            else if (kk == K0 && gb1) atomicAdd(gb1 + jj, v[i2]);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MPB_ISSUER) tmem_dealloc(tmem, 512u);
}

static int g_pipe_enabled = 1;

static bool pipe_shape_ok(int K0, int H, int N, Mlp2Shape* S) {
  if (H != MP_H || K0 < 1 || K0 + 1 > MP_K0P || N < 1 || N > MP_NP) return false;
  S->K0 = K0; S->H = H; S->N = N; S->K0p = MP_K0P; S->Np = MP_NP;
  return true;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

/* experiment knob: device buffer of 60000 int64 for the timeline trace of CTA 0 (NULL switches it off) */
int ffb_mlp2p_trace(long long* d_buf) {
  int zero = 0;
  FFB_CUDA(cudaMemcpyToSymbol(g_mp_trace, &d_buf, sizeof(d_buf)));
  FFB_CUDA(cudaMemcpyToSymbol(g_mp_trace_n, &zero, sizeof(zero)));
  return FFB_OK;
}

int ffb_set_mlp_pipelined(int enabled) {
  g_pipe_enabled = enabled ? 1 : 0;
  return FFB_OK;
}

/* 1 when the pipelined kernels take this shape: hidden width 64, K0 <= 31 inputs, N <= 32 outputs (linear_mat of nerf.yaml). */
int ffb_mlp2_pipelined_eligible(int32_t K0, int32_t H, int32_t N) {
  Mlp2Shape S;
  if (!g_pipe_enabled || !ffb_tensor_cores_enabled() || !pipe_shape_ok(K0, H, N, &S)) return 0;
  return smem_optin_bytes() >= (int)(MpfSmem::TOTAL > MpbSmem::TOTAL ? MpfSmem::TOTAL : MpbSmem::TOTAL) ? 1 : 0;
}

int ffb_mlp2p_fwd(const float* x, const float* W1, const float* b1, const float* W2, float* y, uint16_t* relu_mask, int64_t n,
                  const int32_t* n_dev, int32_t K0, int32_t H, int32_t N, void* stream) {
  FFB_REQUIRE(x && W1 && b1 && W2 && y, "null argument");
  Mlp2Shape S;
  FFB_REQUIRE(pipe_shape_ok(K0, H, N, &S), "MLP shape not eligible for the pipelined tensor-core path");
  if (n <= 0) return FFB_OK;
  static PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(mlp2p_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MpfSmem::TOTAL));
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  mlp2p_fwd_kernel<<<grid, MPF_THREADS, MpfSmem::TOTAL, (cudaStream_t)stream>>>(x, W1, b1, W2, y, relu_mask, n, n_dev, S);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_mlp2p_bwd(const float* x, const float* gy, const float* W1, const float* b1, const float* W2, const uint16_t* relu_mask, float* gx,
                  float* gW1, float* gb1, float* gW2, int64_t n, const int32_t* n_dev, int32_t K0, int32_t H, int32_t N, void* stream) {
  FFB_REQUIRE(x && gy && W1 && b1 && W2, "null argument");
  Mlp2Shape S;
  FFB_REQUIRE(pipe_shape_ok(K0, H, N, &S), "MLP shape not eligible for the pipelined tensor-core path");
  if (n <= 0) return FFB_OK;
  static PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(mlp2p_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MpbSmem::TOTAL));
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  mlp2p_bwd_kernel<<<grid, MPB_THREADS, MpbSmem::TOTAL, (cudaStream_t)stream>>>(x, gy, W1, b1, W2, relu_mask, gx, gW1, gb1, gW2, n, n_dev, S);
  FFB_LAUNCHED();
  return FFB_OK;
}

/* The same backward for a SPARSE upstream gradient (the render path: every sample has a density gradient, only the shaded ones have
 * feature gradients): gy[i, 0] = g0[i]; gy[i, 1:] = g_rows[row_slot[i], 1:] where row_slot[i] >= 0, zero elsewhere.  Saves the dense
 * [n, N] gradient tensor (126 MB written and read per nerf.yaml step). */
int ffb_mlp2p_bwd_sparse(const float* x, const float* g0, const int32_t* row_slot, const float* g_rows, const float* W1, const float* b1,
                         const float* W2, const uint16_t* relu_mask, float* gx, float* gW1, float* gb1, float* gW2, int64_t n,
                         const int32_t* n_dev, int32_t K0, int32_t H, int32_t N, void* stream) {
  FFB_REQUIRE(x && g0 && row_slot && g_rows && W1 && b1 && W2, "null argument");
  Mlp2Shape S;
  FFB_REQUIRE(pipe_shape_ok(K0, H, N, &S) && g_pipe_enabled, "MLP shape not eligible for the pipelined tensor-core path");
  if (n <= 0) return FFB_OK;
  static PerDeviceOnce once;
  if (once.first()) FFB_CUDA(cudaFuncSetAttribute(mlp2p_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MpbSmem::TOTAL));
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  mlp2p_bwd_kernel<<<grid, MPB_THREADS, MpbSmem::TOTAL, (cudaStream_t)stream>>>(x, nullptr, W1, b1, W2, relu_mask, gx, gW1, gb1, gW2, n, n_dev, S,
                                                                                g0, row_slot, g_rows);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
