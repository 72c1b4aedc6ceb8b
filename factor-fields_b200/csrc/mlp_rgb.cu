// mlp_rgb.cu — the appearance MLP (MLPRender_Fea, FactorFields.py:162-203) as ONE tcgen05 kernel:
//   input assembly [features | viewdirs | PE(features, fea_pe) | PE(viewdirs, view_pe)]  (:188-196)
//   -> Linear(K0 -> 128) + ReLU -> Linear(128 -> 128) + ReLU -> Linear(128 -> 3, no bias) -> sigmoid   (:197-202)
// for the shaded samples (weight > rayMarch_weight_thres, :879-885).  The per-layer path (mlp_tc.cu) writes the [Na, 194]
// input and both [Na, 128] hidden activations to HBM and reads them back; here a 128-sample tile stays on the SM from the
// gathered feature row to the colour.
//
// Precision: every product is the 3-part bf16 split (6 MMAs, ~3e-7 relative, tc_common.cuh) — with 2 parts enough ReLU
// decisions flip against the reference that hidden-layer weight gradients are off by 5e-4 (scratch/diag_relu_flip.py).
//
// Shared memory cannot hold 3 parts of W1 (160 KB) + W2 (96 KB) next to a 3-part activation tile (160 KB), so the
// weights are pre-split ONCE per step into bf16 operand slices in UMMA layout (rgb_pack_kernel; they change every
// optimiser step) and streamed through a 4-slot ring by the TMA engine (cp.async.bulk: the bytes are already in
// operand layout, no register pass) while the activation tile is resident.  Warp-specialised: warps 0-7 assemble /
// run the epilogues, warp 8 issues the MMAs, warp 9 feeds the ring; mbarriers carry every hand-off.
#include "tc_common.cuh"
#include "ffb_math.h"

namespace ffb {

constexpr int RGB_H = 128;               // hidden width of both layers
constexpr int RGB_WORKERS = 512;         // forward: threads of the 16 worker warps (the kernel is bound by per-warp instruction latency)
constexpr int RGB_THREADS = RGB_WORKERS + 64;   // + MMA warp + producer warp
constexpr int RGB_B_WORKERS = 512;       // backward: 16 worker warps as well
constexpr int RGB_B_THREADS = RGB_B_WORKERS + 64;
constexpr int RGB_NSLOT = 4;
constexpr uint32_t RGB_SLICE = 4096;     // one K=16 slice of a 128-row operand tile, one bf16 part
constexpr uint32_t RGB_SLOT = 3 * RGB_SLICE;
constexpr uint32_t RGB_W3_PART = 16 * RGB_H * 2;   // W3 tile: 16 rows (3 used) x 128 cols
constexpr int RGB_MAXQ = 10;             // channel quads per row: Cf + 3 <= 40
constexpr int RGB_NPRE = RGB_MAXQ;       // prefetched raw values per lane (one 8-row group per worker warp)

struct RgbShape {
  int Cf, view_pe, fea_pe;
  int K0;      // 3 + Cf + 6 view_pe + 2 fea_pe Cf
  int K0p;     // K0 + 1 (bias column) rounded up to 16
  int n1;      // K0p / 16 slices of W1
};

__host__ __device__ inline bool rgb_shape(int Cf, int view_pe, int fea_pe, RgbShape* S) {
  if (Cf < 1 || Cf + 3 > 4 * RGB_MAXQ || view_pe < 0 || fea_pe < 0 || view_pe > 16 || fea_pe > 16) return false;
  S->Cf = Cf; S->view_pe = view_pe; S->fea_pe = fea_pe;
  S->K0 = 3 + Cf + 6 * view_pe + 2 * fea_pe * Cf;
  S->K0p = (S->K0 + 1 + 15) / 16 * 16;
  S->n1 = S->K0p / 16;
  // activation tile: 3 parts x 128 x K0p x 2 B must leave room for the ring (<= 208); the backward kernel produces g_x in a
  // 128-column block plus a (K0p - 128)-column block, so narrower inputs stay on the per-layer kernels
  return S->K0p > 128 && S->K0p <= 208;
}
__host__ __device__ inline size_t rgb_pack_fwd_bytes(const RgbShape& S) { return (size_t)(S.n1 + RGB_H / 16) * RGB_SLOT + 3 * RGB_W3_PART; }
// backward operands (2 parts): W2^T slices [8][2][4096], W1^T rows 0..127 slices [8][2][4096], W1^T rows 128..K0p-1 slices [8][2][(K0p-128)*32]
__host__ __device__ inline uint32_t rgb_w1b_part(const RgbShape& S) { return (uint32_t)(S.K0p - 128) * 32u; }
__host__ __device__ inline size_t rgb_pack_bwd_bytes(const RgbShape& S) { return (size_t)8 * 2 * (4096 + 4096 + rgb_w1b_part(S)); }
__host__ __device__ inline size_t rgb_pack_bytes(const RgbShape& S) { return rgb_pack_fwd_bytes(S) + rgb_pack_bwd_bytes(S); }

// ---- weights -> bf16 3-part operand slices -----------------------------------------------------------------
// slice s of a [128 x K] weight (rows = output unit j, cols = k): for part t, element (j, kk = k - 16 s) at
//   s * RGB_SLOT + t * RGB_SLICE + (kk / 8) * 2048 + (j / 8) * 128 + (j % 8) * 16 + (kk % 8) * 2
// i.e. a K-major operand tile of 128 rows with column-chunk stride 2048 (LBO 2048, SBO 128).
__global__ void rgb_pack_kernel(const float* __restrict__ W1, const float* __restrict__ b1, const float* __restrict__ W2,
                                const float* __restrict__ W3, uint8_t* __restrict__ out, const RgbShape S) {
  const int nW1 = RGB_H * S.K0p, nW2 = RGB_H * RGB_H, nW3 = 16 * RGB_H;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nW1 + nW2 + nW3; e += gridDim.x * blockDim.x) {
    float v;
    uint8_t* dst;
    uint32_t part_stride;
    if (e < nW1 + nW2) {
      const bool first = e < nW1;
      const int q = first ? e : e - nW1;
      const int K = first ? S.K0p : RGB_H;
      const int j = q / K, k = q % K;
      if (first) v = k < S.K0 ? W1[(size_t)j * S.K0 + k] : (k == S.K0 ? b1[j] : 0.0f);
      else v = W2[(size_t)j * RGB_H + k];
      const int s = (k >> 4) + (first ? 0 : S.n1), kk = k & 15;
      dst = out + (size_t)s * RGB_SLOT + (uint32_t)(kk >> 3) * 2048u + (uint32_t)(j >> 3) * 128u + (uint32_t)(j & 7) * 16u + (uint32_t)(kk & 7) * 2u;
      part_stride = RGB_SLICE;
    } else {
      const int q = e - nW1 - nW2;
      const int r = q / RGB_H, c = q % RGB_H;
      v = r < 3 ? W3[(size_t)r * RGB_H + c] : 0.0f;
      dst = out + (size_t)(S.n1 + RGB_H / 16) * RGB_SLOT + (uint32_t)(c >> 3) * 256u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u +
            (uint32_t)(c & 7) * 2u;
      part_stride = RGB_W3_PART;
    }
    __nv_bfloat16 parts[3];
    split_bf16<3>(v, parts);
#pragma unroll
    for (int t = 0; t < 3; ++t) *reinterpret_cast<__nv_bfloat16*>(dst + (size_t)t * part_stride) = parts[t];
  }
}

// Transposed weights for the input-gradient GEMMs (B operand rows = the layer's INPUT index, K = its output index), 2 parts:
//   W2^T : element (j1, j2) = W2[j2][j1];   W1^T : element (k, j1) = W1[j1][k] for k < K0, 0 for the bias / padding rows
// slice s = K columns 16 s .. 16 s + 15, K-major: (kk / 8) * (rows * 16) + (row / 8) * 128 + (row % 8) * 16 + (kk % 8) * 2.
__global__ void rgb_pack_bwd_kernel(const float* __restrict__ W1, const float* __restrict__ W2, uint8_t* __restrict__ out, const RgbShape S) {
  const int nA = RGB_H * RGB_H, nB = RGB_H * RGB_H, nC = (S.K0p - 128) * RGB_H;
  const uint32_t pB = rgb_w1b_part(S);
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < nA + nB + nC; e += gridDim.x * blockDim.x) {
    int row, kcol, rows;
    float v;
    uint8_t* base;
    uint32_t part;
    if (e < nA) {                      // W2^T
      row = e / RGB_H; kcol = e % RGB_H; rows = 128; part = 4096;
      v = W2[(size_t)kcol * RGB_H + row];
      base = out;
    } else if (e < nA + nB) {          // W1^T, k = 0..127
      const int q = e - nA;
      row = q / RGB_H; kcol = q % RGB_H; rows = 128; part = 4096;
      v = row < S.K0 ? W1[(size_t)kcol * S.K0 + row] : 0.0f;
      base = out + (size_t)8 * 2 * 4096;
    } else {                           // W1^T, k = 128..K0p-1
      const int q = e - nA - nB;
      row = q / RGB_H; kcol = q % RGB_H; rows = S.K0p - 128; part = pB;
      v = (row + 128) < S.K0 ? W1[(size_t)kcol * S.K0 + row + 128] : 0.0f;
      base = out + (size_t)8 * 2 * 4096 * 2;
    }
    const int sidx = kcol >> 4, kk = kcol & 15;
    uint8_t* dst = base + (size_t)sidx * 2 * part + (uint32_t)(kk >> 3) * (uint32_t)(rows * 16) + (uint32_t)(row >> 3) * 128u + (uint32_t)(row & 7) * 16u +
                   (uint32_t)(kk & 7) * 2u;
    __nv_bfloat16 parts[2];
    split_bf16<2>(v, parts);
    *reinterpret_cast<__nv_bfloat16*>(dst) = parts[0];
    *reinterpret_cast<__nv_bfloat16*>(dst + part) = parts[1];
  }
}

struct RgbFwdArgs {
  const float* feat; int ld_feat;          // [Nv, ld]: linear_mat output, features = columns 1..Cf
  const float* rays;                       // [R, 6]: view direction = columns 3..5
  const int32_t* ray_id;                   // [Nv] or NULL (row i of feat uses rays[i])
  const int32_t* app_idx;                  // [n] rows of feat that are shaded, or NULL (identity)
  const uint8_t* wpack;                    // rgb_pack_kernel output
  const float* b2;
  float* rgb;                              // [n, 3]
  uint16_t* bits;                          // [n, 16]: ReLU decisions, layer 1 = words 0..7, layer 2 = words 8..15 (or NULL)
  float* x_out; float* h1_out; float* h2_out;   // optional fp32 copies of the MLP input [n, K0] / hidden activations [n, 128]
  uint8_t* sx; uint8_t* sh1; uint8_t* sh2;      // optional bf16 2-part row-slice streams for rgb_bwd_kernel (layout below)
  int64_t n; const int32_t* n_dev;
  RgbShape S;
};

// Row-slice streams handed to the backward kernel: per 128-row tile, 8 slices of 16 rows; per slice [part 0][part 1]; per
// part, element (r, c) at (c / 8) * 256 + ((r % 16) / 8) * 128 + (r % 8) * 16 + (c % 8) * 2 — an MN-major operand slice
// (K = the 16 rows: LBO 128; M/N = columns: SBO 256) that the TMA engine can drop into shared memory as it is.
__host__ __device__ inline uint32_t rgb_stream_part(int cols) { return (uint32_t)(cols / 8) * 256u; }
__host__ __device__ inline size_t rgb_stream_tile(int cols) { return (size_t)8 * 2 * rgb_stream_part(cols); }
__device__ __forceinline__ uint8_t* stream_ptr(uint8_t* base, int cols, int64_t tile, int r, int c8, int part) {
  return base + (size_t)tile * rgb_stream_tile(cols) + (size_t)(r >> 4) * 2 * rgb_stream_part(cols) + (size_t)part * rgb_stream_part(cols) +
         (uint32_t)c8 * 256u + (uint32_t)((r & 15) >> 3) * 128u + (uint32_t)(r & 7) * 16u;
}

// element (r, c) of the resident activation tile, part t: t * szA + (c/8) * 2048 + (r/8) * 128 + (r%8) * 16 + (c%8) * 2
__device__ __forceinline__ uint32_t a_off(int r, int c) {
  return (uint32_t)(c >> 3) * 2048u + (uint32_t)(r >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(c & 7) * 2u;
}
__device__ __forceinline__ void a_store3(uint8_t* sA, uint32_t szA, int r, int c, float v) {
  __nv_bfloat16 parts[3];
  split_bf16<3>(v, parts);
  uint8_t* p = sA + a_off(r, c);
#pragma unroll
  for (int t = 0; t < 3; ++t) *reinterpret_cast<__nv_bfloat16*>(p + (size_t)t * szA) = parts[t];
}

// two adjacent columns (c even) of one row: one 4-byte store per part
__device__ __forceinline__ void a_store3_pair(uint8_t* sA, uint32_t szA, int r, int c, float v0, float v1) {
  uint32_t w[3];
  split2_packed<3>(v0, v1, w);
  uint8_t* p = sA + a_off(r, c);
#pragma unroll
  for (int t = 0; t < 3; ++t) *reinterpret_cast<uint32_t*>(p + (size_t)t * szA) = w[t];
}

__global__ void __launch_bounds__(RGB_THREADS, 1) rgb_fwd_kernel(const RgbFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const RgbShape S = a.S;
  const int64_t n = resolve_n(a.n, a.n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t szA = 128u * (uint32_t)S.K0p * 2u;                 // one part of the activation tile
  uint8_t* sA = smem;                                               // 3 parts: X, then H1, then H2 (first 128 columns)
  uint8_t* sRing = sA + 3 * szA;
  uint8_t* sW3 = sRing + RGB_NSLOT * RGB_SLOT;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sW3 + 3 * RGB_W3_PART);
  uint64_t* full = bars;                    // [NSLOT] ring slot filled (tx bytes)
  uint64_t* empty = bars + RGB_NSLOT;       // [NSLOT] ring slot consumed (tcgen05.commit)
  uint64_t* a_ready = bars + 2 * RGB_NSLOT; // workers -> MMA warp: the activation tile is staged
  uint64_t* mma_done = a_ready + 1;         // MMA warp -> workers: the layer's accumulator is complete
  uint64_t* w3_full = mma_done + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w3_full + 1);

  if (warp == RGB_WORKERS / 32) tmem_alloc(tmem_slot, 256);
  if (tid == 0) {
    for (int i = 0; i < RGB_NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(a_ready, 1);
    mbar_init(mma_done, 1);
    mbar_init(w3_full, 1);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t d1 = tmem, d2 = tmem + RGB_H, d3 = tmem;           // D3 re-uses D1's first columns (D1 is consumed by then)
  const int64_t n_tiles = (n + 127) / 128;
  const int n_slices = S.n1 + RGB_H / 16;                            // ring traffic per tile: W1 slices then W2 slices
  const int rot1 = (int)(blockIdx.x % (unsigned)S.n1), rot2 = (int)(blockIdx.x % (unsigned)(RGB_H / 16));

  if (warp == RGB_WORKERS / 32 + 1) {
    // ---------------- producer: W3 once, then the same (W1, W2) slice sequence for every tile of this CTA
    if ((int64_t)blockIdx.x < n_tiles && elect_one()) {
      mbar_expect_tx(w3_full, 3 * RGB_W3_PART);
      bulk_g2s(sW3, a.wpack + (size_t)n_slices * RGB_SLOT, 3 * RGB_W3_PART, w3_full);   // == wpack + fwd bytes - W3 tile
      // Every CTA walks the K slices of a layer in its own rotation (the sum over slices is order-free): the 148 CTAs run
      // in near lockstep, and without the rotation they all pull the same L2 lines at the same moment.
      uint32_t seq = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
        for (int q = 0; q < n_slices; ++q, ++seq) {
          const int s = q < S.n1 ? (q + rot1) % S.n1 : S.n1 + (q - S.n1 + rot2) % (RGB_H / 16);
          const uint32_t slot = seq % RGB_NSLOT, use = seq / RGB_NSLOT;
          if (use > 0) mbar_wait(empty + slot, (use - 1) & 1);
          mbar_expect_tx(full + slot, RGB_SLOT);
          bulk_g2s(sRing + slot * RGB_SLOT, a.wpack + (size_t)s * RGB_SLOT, RGB_SLOT, full + slot);
        }
    }
  } else if (warp == RGB_WORKERS / 32) {
    // ---------------- MMA issuer
    if ((int64_t)blockIdx.x < n_tiles && elect_one()) {
      const uint32_t aA = smem_u32(sA), aRing = smem_u32(sRing), aW3 = smem_u32(sW3);
      const uint32_t idH = make_idesc(RGB_H, 0, 0), id3 = make_idesc(16, 0, 0);
      uint32_t seq = 0, ph_a = 0;
      mbar_wait(w3_full, 0);
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int layer = 0; layer < 2; ++layer) {
          mbar_wait(a_ready, ph_a);
          ph_a ^= 1;
          tc_fence_after();
          const int ns = layer == 0 ? S.n1 : RGB_H / 16;
          const uint32_t dst = layer == 0 ? d1 : d2;
          for (int q = 0; q < ns; ++q, ++seq) {
            const int s = (q + (layer == 0 ? rot1 : rot2)) % ns;       // the slice the producer put in this slot
            const uint32_t slot = seq % RGB_NSLOT, use = seq / RGB_NSLOT;
            mbar_wait(full + slot, use & 1);
            tc_fence_after();
            const uint32_t wb = aRing + slot * RGB_SLOT;
            uint32_t acc = q > 0 ? 1u : 0u;
#pragma unroll
            for (int ta = 0; ta < 3; ++ta)
#pragma unroll
              for (int tb = 0; tb < 3; ++tb) {
                if (ta + tb >= 3) continue;
                umma_f16(dst, make_desc(aA + ta * szA + (uint32_t)s * 4096u, 2048u, 128u), make_desc(wb + tb * RGB_SLICE, 2048u, 128u), idH, acc);
                acc = 1u;
              }
            umma_commit(empty + slot);      // the slot may be refilled once these MMAs have read it
          }
          umma_commit(mma_done);
        }
        // colour head: D3[128 x 16] = H2 * W3^T, W3 resident
        mbar_wait(a_ready, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        for (int s = 0; s < RGB_H / 16; ++s) {
          uint32_t acc = s > 0 ? 1u : 0u;
#pragma unroll
          for (int ta = 0; ta < 3; ++ta)
#pragma unroll
            for (int tb = 0; tb < 3; ++tb) {
              if (ta + tb >= 3) continue;
              umma_f16(d3, make_desc(aA + ta * szA + (uint32_t)s * 4096u, 2048u, 128u),
                       make_desc(aW3 + tb * RGB_W3_PART + (uint32_t)s * 512u, 256u, 128u), id3, acc);
              acc = 1u;
            }
        }
        umma_commit(mma_done);
      }
    }
  } else {
    // ---------------- workers: input assembly and the three epilogues
    const int rloc = (warp & 3) * 32 + lane, quarter = warp >> 2;      // TMEM lane quarter = warp % 4; column blocks by warp / 4
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int C = S.Cf, nch = C + 3;
    const int oV = C, oPF = C + 3, oPV = C + 3 + 2 * S.fea_pe * C;
    const int ncq = (nch + 3) >> 2;
    float pre[RGB_NPRE];
    // raw MLP inputs of this lane's (row, channel) items: feat[app_idx[j], 1 + ch] or the ray's view direction
    auto load_raw = [&](int64_t row0) {
      const int64_t j = row0 + warp * 8 + (lane >> 2);
      const int64_t src = j < n ? (a.app_idx ? (int64_t)__ldg(a.app_idx + j) : j) : -1;
      const int64_t ray = src < 0 ? -1 : (a.ray_id ? (int64_t)__ldg(a.ray_id + src) : src);
#pragma unroll
      for (int k = 0; k < RGB_NPRE; ++k) {
        const int ch = k * 4 + (lane & 3);
        pre[k] = 0.0f;
        if (k < ncq && ch < nch && src >= 0) pre[k] = ch < C ? __ldg(a.feat + src * a.ld_feat + 1 + ch) : __ldg(a.rays + ray * 6 + 3 + (ch - C));
      }
    };
    uint32_t ph_m = 0;
    if ((int64_t)blockIdx.x < n_tiles) load_raw((int64_t)blockIdx.x * 128);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row0 = tile * 128;
      // --- X tile: [features | viewdirs | sin/cos PE of both | 1 (bias column) | 0 padding]; rows beyond n are zero
      // Work item = (8-row group, 4 consecutive channels); lane = (row in group, channel in quad).  The shared-memory bank of
      // tile element (r, c) is 4 (r % 8) + (c % 8) / 2, so the 32 lanes of a warp hit 32 different banks (or the two halves
      // of one word) for the identity columns and for every octave of the sin / cos blocks.  Worker warp w owns row group
      // w; the raw values were fetched into registers while the previous tile was in the tensor pipe.
#pragma unroll
      for (int k = 0; k < RGB_NPRE; ++k) {
        const int r = warp * 8 + (lane >> 2), ch = k * 4 + (lane & 3);
        if (k >= ncq || ch >= nch) continue;
        const int64_t j = row0 + r;
        const bool live = j < n;
        const float v = pre[k];
        int pe, o_id, o_sin, o_cos;
        if (ch < C) {
          pe = S.fea_pe; o_id = ch; o_sin = oPF + ch * S.fea_pe; o_cos = oPF + C * S.fea_pe + ch * S.fea_pe;
        } else {
          const int d = ch - C;
          pe = S.view_pe; o_id = oV + d; o_sin = oPV + d * S.view_pe; o_cos = oPV + 3 * S.view_pe + d * S.view_pe;
        }
        a_store3(sA, szA, r, o_id, v);
        if (a.x_out && live) a.x_out[j * S.K0 + o_id] = v;
        float sa, ca;
        sincosf(v, &sa, &ca);
        if (!live) sa = ca = 0.0f;       // rows beyond n are all-zero (doubling keeps them zero)
        const bool pairs = ((pe | o_sin | o_cos) & 1) == 0;     // octaves (2k, 2k+1) share a 4-byte word
        for (int kk = 0; kk < pe; kk += 2) {
          const float s2 = 2.0f * sa * ca, c2 = (ca - sa) * (ca + sa);     // exact angle doubling for the next octave
          if (pairs) {
            a_store3_pair(sA, szA, r, o_sin + kk, sa, s2);
            a_store3_pair(sA, szA, r, o_cos + kk, ca, c2);
          } else {
            a_store3(sA, szA, r, o_sin + kk, sa);
            a_store3(sA, szA, r, o_cos + kk, ca);
            if (kk + 1 < pe) { a_store3(sA, szA, r, o_sin + kk + 1, s2); a_store3(sA, szA, r, o_cos + kk + 1, c2); }
          }
          if (a.x_out && live) {
            a.x_out[j * S.K0 + o_sin + kk] = sa; a.x_out[j * S.K0 + o_cos + kk] = ca;
            if (kk + 1 < pe) { a.x_out[j * S.K0 + o_sin + kk + 1] = s2; a.x_out[j * S.K0 + o_cos + kk + 1] = c2; }
          }
          sa = 2.0f * s2 * c2;
          ca = (c2 - s2) * (c2 + s2);
        }
      }
      for (int it = tid; it < 128 * (S.K0p - S.K0); it += RGB_WORKERS) {
        const int r = it / (S.K0p - S.K0), c = S.K0 + it % (S.K0p - S.K0);
        a_store3(sA, szA, r, c, (c == S.K0 && row0 + r < n) ? 1.0f : 0.0f);
      }
      proxy_fence();
      named_sync(1, RGB_WORKERS);
      if (tid == 0) mbar_arrive(a_ready);
      if (tile + gridDim.x < n_tiles) load_raw((tile + gridDim.x) * 128);      // next tile's inputs travel during this tile's GEMMs
      const int64_t row = row0 + rloc;
      if (a.sx) {
        // the staged X tile, parts 0 and 1, re-laid as row slices for the backward kernel; a thread copies exactly the
        // chunks of its row that its own epilogue will overwrite later (plus its share of the columns beyond 128)
        for (int c8 = 0; c8 < S.K0p / 8; ++c8) {
          if (((c8 >> 1) & 3) != quarter) continue;
#pragma unroll
          for (int t = 0; t < 2; ++t)
            *reinterpret_cast<uint4*>(stream_ptr(a.sx, S.K0p, tile, rloc, c8, t)) = *reinterpret_cast<const uint4*>(sA + (size_t)t * szA + a_off(rloc, c8 * 8));
        }
      }
      // --- hidden layers: TMEM -> (+ bias) -> ReLU (+ decision bits) -> bf16 parts -> operand tile of the next GEMM
      for (int layer = 0; layer < 2; ++layer) {
        mbar_wait(mma_done, ph_m);
        ph_m ^= 1;
        tc_fence_after();
        const uint32_t src = layer == 0 ? d1 : d2;
        float* hout = layer == 0 ? a.h1_out : a.h2_out;
        uint8_t* sh = layer == 0 ? a.sh1 : a.sh2;
        for (int c0 = quarter * 16; c0 < RGB_H; c0 += 64) {
          float v[16];
          tmem_ld16(src + lane_base + (uint32_t)c0, v);
          uint32_t bits = 0;
          if (layer == 1) {
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 bb = __ldg(reinterpret_cast<const float4*>(a.b2 + c0 + i));
              v[i] += bb.x; v[i + 1] += bb.y; v[i + 2] += bb.z; v[i + 3] += bb.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            bits |= (v[i] > 0.0f ? 1u : 0u) << i;
            v[i] = fmaxf(v[i], 0.0f);
          }
          if (row < n) {
            if (a.bits) a.bits[row * 16 + layer * 8 + (c0 >> 4)] = (uint16_t)bits;
            if (hout) {
#pragma unroll
              for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(hout + row * RGB_H + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
            }
          }
          uint4 parts[3];
          split8_packed<3>(v, parts);
          uint8_t* p = sA + a_off(rloc, c0);
#pragma unroll
          for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(p + (size_t)t * szA) = parts[t];
          if (sh) {
            *reinterpret_cast<uint4*>(stream_ptr(sh, RGB_H, tile, rloc, c0 >> 3, 0)) = parts[0];
            *reinterpret_cast<uint4*>(stream_ptr(sh, RGB_H, tile, rloc, c0 >> 3, 1)) = parts[1];
          }
          split8_packed<3>(v + 8, parts);
          p = sA + a_off(rloc, c0 + 8);
#pragma unroll
          for (int t = 0; t < 3; ++t) *reinterpret_cast<uint4*>(p + (size_t)t * szA) = parts[t];
          if (sh) {
            *reinterpret_cast<uint4*>(stream_ptr(sh, RGB_H, tile, rloc, (c0 >> 3) + 1, 0)) = parts[0];
            *reinterpret_cast<uint4*>(stream_ptr(sh, RGB_H, tile, rloc, (c0 >> 3) + 1, 1)) = parts[1];
          }
        }
        tc_fence_before();
        proxy_fence();
        named_sync(1, RGB_WORKERS);
        if (tid == 0) mbar_arrive(a_ready);
      }
      // --- colour head: sigmoid (FactorFields.py:200-202)
      mbar_wait(mma_done, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      if (quarter == 0) {
        float v[16];
        tmem_ld16(d3 + lane_base, v);
        if (row < n) {
          a.rgb[row * 3 + 0] = sigmoid_f(v[0]);
          a.rgb[row * 3 + 1] = sigmoid_f(v[1]);
          a.rgb[row * 3 + 2] = sigmoid_f(v[2]);
        }
      }
      tc_fence_before();
      named_sync(1, RGB_WORKERS);     // every TMEM read of this tile is done before the next tile's a_ready can be signalled
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RGB_WORKERS / 32) {
    __syncwarp();
    tmem_dealloc(tmem, 256);
  }
}

// =============================================================================================================
// backward.  Per 128-row tile (2-part bf16 split for every operand, as the per-layer gradient kernels):
//   g3   = g_rgb * rgb (1 - rgb)                                  (sigmoid')                        [SIMT]
//   G2   = (g3 W3) .* [h2 > 0]                                     K = 3: formed in the epilogue     [SIMT -> smem tile]
//   gW3 += H2^T g3          gW2 += G2^T H1        gb2 += G2^T 1
//   G1   = (G2 W2) .* [h1 > 0]
//   g_x  = G1 W1            gW1 | gb1 += G1^T [x | 1]
// Only the G tile (G2, later G1) lives in shared memory; H2, H1 and X arrive as the row-slice streams the forward kernel
// wrote, and W2^T / W1^T as pre-split slices, all through one TMA ring (they are B operands — or, for H2, an MN-major
// A operand — so a 16-row / 16-column slice at a time is all a GEMM step needs).  The four weight-gradient
// accumulators stay in TMEM across every tile of the persistent CTA and are flushed once with atomics.
// TMEM columns: gW1|gb1 [0, K0p)   gW2 [208, 336)   gW3 [336, 352)   gb2 [352, 368)   scratch (G2 W2, then g_x halves) [368, 496)
// =============================================================================================================
constexpr int RGB_B_NSLOT = 10;
constexpr uint32_t RGB_B_SLOT = 13312;          // largest ring item: an X slice, 2 parts x 26 chunks x 256 B
constexpr uint32_t RGB_G_PART = 128 * RGB_H * 2;  // one part of the G tile

struct RgbBwdArgs {
  const float* g_rgb; const float* rgb;     // [n, 3]: upstream gradient and the forward output
  const uint16_t* bits;                     // [n, 16]
  const uint8_t* sx; const uint8_t* sh1; const uint8_t* sh2;   // forward streams
  const uint8_t* wpack;                     // backward part of the workspace (rgb_pack_bwd_kernel)
  const float* W3;                          // [3, 128] fp32
  float* g_x; int ld_gx;                    // [n, ld_gx], ld_gx >= K0 and a multiple of 4 (16-byte aligned rows)
  float* gW1; float* gb1; float* gW2; float* gb2; float* gW3;   // accumulated into (atomics)
  int64_t n; const int32_t* n_dev;
  RgbShape S;
};

__device__ __forceinline__ void red_add2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void g_store2(uint8_t* sG, int r, int c0, const float v[8]) {
  uint4 parts[2];
  split8_packed<2>(v, parts);
  uint8_t* p = sG + a_off(r, c0);
  *reinterpret_cast<uint4*>(p) = parts[0];
  *reinterpret_cast<uint4*>(p + RGB_G_PART) = parts[1];
}

template <class DA, class DB>
__device__ __forceinline__ void mma2(uint32_t d, uint32_t idesc, bool accumulate, DA da, DB db) {
  umma_f16(d, da(0), db(0), idesc, accumulate ? 1u : 0u);
  umma_f16(d, da(0), db(1), idesc, 1u);
  umma_f16(d, da(1), db(0), idesc, 1u);
}

__global__ void __launch_bounds__(RGB_B_THREADS, 1) rgb_bwd_kernel(const RgbBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const RgbShape S = a.S;
  const int64_t n = resolve_n(a.n, a.n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sG = smem;                                   // G tile, 2 parts (K-major; also read MN-major)
  uint8_t* sG3 = sG + 2 * RGB_G_PART;                   // g3 tile [128 x 16], 2 parts of 4096 B
  uint8_t* sOnes = sG3 + 2 * 4096;                      // [128 x 16] of 1.0 (one part)
  float* sW3 = reinterpret_cast<float*>(sOnes + 4096);  // [3][128]
  uint8_t* sRing = reinterpret_cast<uint8_t*>(sW3 + 3 * RGB_H);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sRing + RGB_B_NSLOT * RGB_B_SLOT);
  uint64_t* full = bars;
  uint64_t* empty = bars + RGB_B_NSLOT;
  uint64_t* a_ready = bars + 2 * RGB_B_NSLOT;
  uint64_t* mma_done = a_ready + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mma_done + 1);

  if (warp == RGB_B_WORKERS / 32) tmem_alloc(tmem_slot, 512);
  if (tid == 0) {
    for (int i = 0; i < RGB_B_NSLOT; ++i) { mbar_init(full + i, 1); mbar_init(empty + i, 1); }
    mbar_init(a_ready, 1);
    mbar_init(mma_done, 1);
  }
  for (int i = tid; i < 3 * RGB_H; i += blockDim.x) sW3[i] = a.W3[i];
  for (int i = tid; i < 128 * 16; i += blockDim.x)     // ones tile, same layout as the g3 tile
    *reinterpret_cast<__nv_bfloat16*>(sOnes + a_off(i >> 4, i & 15)) = __float2bfloat16_rn(1.0f);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t dW1 = tmem, dW2 = tmem + 208, dW3 = tmem + 336, dB2 = tmem + 352, dS = tmem + 368;
  const int64_t n_tiles = (n + 127) / 128;
  const int xcols = S.K0p, nb = S.K0p - 128;             // X stream width; rows of the second W1^T block
  const uint32_t pH = rgb_stream_part(RGB_H), pX = rgb_stream_part(xcols), pW1b = rgb_w1b_part(S);
  const uint8_t* wW2 = a.wpack;
  const uint8_t* wW1a = a.wpack + (size_t)8 * 2 * 4096;
  const uint8_t* wW1b = a.wpack + (size_t)8 * 2 * 4096 * 2;
  const int rot = (int)(blockIdx.x & 7u);

  if (warp == RGB_B_WORKERS / 32 + 1) {
    // ---------------- producer: 48 ring items per tile, in the order the MMA warp consumes them
    if (elect_one()) {
      uint32_t seq = 0;
      auto put = [&](const uint8_t* src, uint32_t bytes) {
        const uint32_t slot = seq % RGB_B_NSLOT, use = seq / RGB_B_NSLOT;
        if (use > 0) mbar_wait(empty + slot, (use - 1) & 1);
        mbar_expect_tx(full + slot, bytes);
        bulk_g2s(sRing + slot * RGB_B_SLOT, src, bytes, full + slot);
        ++seq;
      };
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint8_t* th2 = a.sh2 + (size_t)tile * rgb_stream_tile(RGB_H);
        const uint8_t* th1 = a.sh1 + (size_t)tile * rgb_stream_tile(RGB_H);
        const uint8_t* tx = a.sx + (size_t)tile * rgb_stream_tile(xcols);
        for (int s = 0; s < 8; ++s) put(th2 + (size_t)s * 2 * pH, 2 * pH);
        for (int s = 0; s < 8; ++s) {
          put(wW2 + (size_t)((s + rot) & 7) * 8192, 8192);        // weight slices in this CTA's own rotation (see rgb_fwd_kernel)
          put(th1 + (size_t)s * 2 * pH, 2 * pH);
        }
        for (int s = 0; s < 8; ++s) {
          put(wW1a + (size_t)((s + rot) & 7) * 8192, 8192);
          put(tx + (size_t)s * 2 * pX, 2 * pX);
        }
        for (int s = 0; s < 8; ++s) put(wW1b + (size_t)((s + rot) & 7) * 2 * pW1b, 2 * pW1b);
      }
    }
  } else if (warp == RGB_B_WORKERS / 32) {
    // ---------------- MMA issuer
    if (elect_one()) {
      const uint32_t aG = smem_u32(sG), aG3 = smem_u32(sG3), aOnes = smem_u32(sOnes), aRing = smem_u32(sRing);
      const uint32_t id_gh = make_idesc(RGB_H, 0, 0);          // scratch = G (K-major) x W^T slice (K-major)
      const uint32_t id_gxb = make_idesc(nb, 0, 0);
      const uint32_t id_w2 = make_idesc(RGB_H, 1, 1);          // gW2 += G^T (MN-major) x H1 slice (MN-major)
      const uint32_t id_w1 = make_idesc(xcols, 1, 1);
      const uint32_t id_16 = make_idesc(16, 1, 1);             // gW3 += H2^T g3 ; gb2 += G^T 1
      uint32_t seq = 0, ph_a = 0;
      bool any = false;
      auto take = [&]() {
        const uint32_t slot = seq % RGB_B_NSLOT, use = seq / RGB_B_NSLOT;
        mbar_wait(full + slot, use & 1);
        tc_fence_after();
        return aRing + slot * RGB_B_SLOT;
      };
      auto release = [&]() {
        umma_commit(empty + (seq % RGB_B_NSLOT));
        ++seq;
      };
      // A operand descriptors of the resident G tile: K-major slice s (K = columns) / MN-major slice s (K = rows)
      auto gk = [&](int t, int s) { return make_desc(aG + t * RGB_G_PART + (uint32_t)s * 4096u, 2048u, 128u); };
      auto gmn = [&](int t, int s) { return make_desc(aG + t * RGB_G_PART + (uint32_t)s * 256u, 128u, 2048u); };
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---- phase A: needs G2 and g3
        mbar_wait(a_ready, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        for (int s = 0; s < 8; ++s) {                             // gW3[j2, c] += sum_r h2[r, j2] g3[r, c]
          const uint32_t b = take();
          mma2(dW3, id_16, any || s > 0, [&](int t) { return make_desc(b + t * pH, 128u, 256u); },
               [&](int t) { return make_desc(aG3 + t * 4096u + (uint32_t)s * 256u, 128u, 2048u); });
          release();
        }
        for (int s = 0; s < 8; ++s) {
          uint32_t b = take();                                    // scratch[r, j1] = sum_j2 G2[r, j2] W2[j2, j1]
          mma2(dS, id_gh, s > 0, [&](int t) { return gk(t, (s + rot) & 7); }, [&](int t) { return make_desc(b + t * 4096u, 2048u, 128u); });
          release();
          b = take();                                             // gW2[j2, j1] += sum_r G2[r, j2] h1[r, j1]
          mma2(dW2, id_w2, any || s > 0, [&](int t) { return gmn(t, s); }, [&](int t) { return make_desc(b + t * pH, 128u, 256u); });
          umma_f16(dB2, gmn(0, s), make_desc(aOnes + (uint32_t)s * 256u, 128u, 2048u), id_16, (any || s > 0) ? 1u : 0u);   // gb2 += G2^T 1
          umma_f16(dB2, gmn(1, s), make_desc(aOnes + (uint32_t)s * 256u, 128u, 2048u), id_16, 1u);
          release();
        }
        umma_commit(mma_done);
        // ---- phase B: needs G1
        mbar_wait(a_ready, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        for (int s = 0; s < 8; ++s) {
          uint32_t b = take();                                    // scratch[r, k] = sum_j1 G1[r, j1] W1[j1, k], k < 128
          mma2(dS, id_gh, s > 0, [&](int t) { return gk(t, (s + rot) & 7); }, [&](int t) { return make_desc(b + t * 4096u, 2048u, 128u); });
          release();
          b = take();                                             // gW1|gb1[j1, k] += sum_r G1[r, j1] [x | 1][r, k]
          mma2(dW1, id_w1, any || s > 0, [&](int t) { return gmn(t, s); }, [&](int t) { return make_desc(b + t * pX, 128u, 256u); });
          release();
        }
        umma_commit(mma_done);
        // ---- phase C: the remaining columns of g_x through the same scratch columns
        mbar_wait(a_ready, ph_a);
        ph_a ^= 1;
        tc_fence_after();
        for (int s = 0; s < 8; ++s) {
          const uint32_t b = take();
          mma2(dS, id_gxb, s > 0, [&](int t) { return gk(t, (s + rot) & 7); }, [&](int t) { return make_desc(b + t * pW1b, (uint32_t)nb * 16u, 128u); });
          release();
        }
        umma_commit(mma_done);
        any = true;
      }
    }
  } else {
    // ---------------- workers
    const int rloc = (warp & 3) * 32 + lane, quarter = warp >> 2;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    uint32_t ph_m = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t row = tile * 128 + rloc;
      const bool live = row < n;
      // g3 = g_rgb * sigmoid'  and  G2 = (g3 W3) .* [h2 > 0]
      float g3[3] = {0.f, 0.f, 0.f};
      if (live) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const float y = a.rgb[row * 3 + c];
          g3[c] = a.g_rgb[row * 3 + c] * y * (1.0f - y);
        }
      }
      if (quarter == 0) {
        float v[8] = {g3[0], g3[1], g3[2], 0.f, 0.f, 0.f, 0.f, 0.f};
        const float z[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        uint4 parts[2];
        split8_parts<2>(v, parts);
        *reinterpret_cast<uint4*>(sG3 + a_off(rloc, 0)) = parts[0];
        *reinterpret_cast<uint4*>(sG3 + 4096 + a_off(rloc, 0)) = parts[1];
        split8_parts<2>(z, parts);
        *reinterpret_cast<uint4*>(sG3 + a_off(rloc, 8)) = parts[0];
        *reinterpret_cast<uint4*>(sG3 + 4096 + a_off(rloc, 8)) = parts[1];
      }
      for (int c0 = quarter * 16; c0 < RGB_H; c0 += 64) {
        const uint32_t bits = live ? (uint32_t)a.bits[row * 16 + 8 + (c0 >> 4)] : 0u;
        float v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int j = c0 + i;
          const float g = g3[0] * sW3[j] + g3[1] * sW3[RGB_H + j] + g3[2] * sW3[2 * RGB_H + j];
          v[i] = ((bits >> i) & 1u) ? g : 0.0f;
        }
        g_store2(sG, rloc, c0, v);
        g_store2(sG, rloc, c0 + 8, v + 8);
      }
      proxy_fence();
      named_sync(1, RGB_B_WORKERS);
      if (tid == 0) mbar_arrive(a_ready);
      // G1 = scratch .* [h1 > 0]
      mbar_wait(mma_done, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      for (int c0 = quarter * 16; c0 < RGB_H; c0 += 64) {
        const uint32_t bits = live ? (uint32_t)a.bits[row * 16 + (c0 >> 4)] : 0u;
        float v[16];
        tmem_ld16(dS + lane_base + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = ((bits >> i) & 1u) ? v[i] : 0.0f;
        g_store2(sG, rloc, c0, v);
        g_store2(sG, rloc, c0 + 8, v + 8);
      }
      tc_fence_before();
      proxy_fence();
      named_sync(1, RGB_B_WORKERS);
      if (tid == 0) mbar_arrive(a_ready);
      // g_x columns [0, 128)
      mbar_wait(mma_done, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      for (int c0 = quarter * 16; c0 < 128; c0 += 64) {
        float v[16];
        tmem_ld16(dS + lane_base + (uint32_t)c0, v);
        if (live) {       // 64 contiguous bytes per lane
#pragma unroll
          for (int i = 0; i < 16; i += 4) *reinterpret_cast<float4*>(a.g_x + row * a.ld_gx + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
      named_sync(1, RGB_B_WORKERS);
      if (tid == 0) mbar_arrive(a_ready);
      // g_x columns [128, K0)
      mbar_wait(mma_done, ph_m);
      ph_m ^= 1;
      tc_fence_after();
      for (int c0 = quarter * 16; c0 < nb; c0 += 64) {
        float v[16];
        tmem_ld16(dS + lane_base + (uint32_t)c0, v);
        if (live) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (128 + c0 + i + 4 <= a.ld_gx) *reinterpret_cast<float4*>(a.g_x + row * a.ld_gx + 128 + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      tc_fence_before();
      named_sync(1, RGB_B_WORKERS);     // scratch drained and G1 consumed before the next tile's G2 / a_ready
    }
    // ---- flush the weight gradients (accumulator row = TMEM lane): wait for the last tile's MMAs first (they were all
    // committed to mma_done, which this thread has already waited on for every phase)
    if ((int64_t)blockIdx.x < n_tiles) {
      tc_fence_after();
      const int m = rloc;
      for (int c0 = quarter * 16; c0 < xcols; c0 += 64) {            // gW1 | gb1: lane j1, column k
        float v[16];
        tmem_ld16(dW1 + lane_base + (uint32_t)c0, v);
        const bool vec2 = a.gW1 && ((S.K0 & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.gW1) & 7) == 0);   // 8-byte aligned column pairs
#pragma unroll
        for (int i = 0; i < 16; i += 2) {
          const int k = c0 + i;
          if (vec2 && k + 1 < S.K0) {
            if (v[i] != 0.0f || v[i + 1] != 0.0f) red_add2(a.gW1 + (size_t)m * S.K0 + k, v[i], v[i + 1]);
            continue;
          }
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            if (v[i + u] == 0.0f) continue;
            if (k + u < S.K0) { if (a.gW1) atomicAdd(a.gW1 + (size_t)m * S.K0 + k + u, v[i + u]); }
            else if (k + u == S.K0 && a.gb1) atomicAdd(a.gb1 + m, v[i + u]);
          }
        }
      }
      for (int c0 = quarter * 16; c0 < RGB_H; c0 += 64) {            // gW2: lane j2, column j1
        float v[16];
        tmem_ld16(dW2 + lane_base + (uint32_t)c0, v);
        if (a.gW2) {
          if ((reinterpret_cast<uintptr_t>(a.gW2) & 15) == 0) {      // 16-byte vector reductions: a quarter of the L2 atomics
#pragma unroll
            for (int i = 0; i < 16; i += 4) red_add4(a.gW2 + (size_t)m * RGB_H + c0 + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (v[i] != 0.0f) atomicAdd(a.gW2 + (size_t)m * RGB_H + c0 + i, v[i]);
          }
        }
      }
      if (quarter == 0) {
        float v[16];
        tmem_ld16(dW3 + lane_base, v);                             // gW3: lane j2, column c
        if (a.gW3) {
#pragma unroll
          for (int c = 0; c < 3; ++c)
            if (v[c] != 0.0f) atomicAdd(a.gW3 + (size_t)c * RGB_H + m, v[c]);
        }
        tmem_ld16(dB2 + lane_base, v);                             // gb2: lane j2 (every column holds the same sum)
        if (a.gb2 && v[0] != 0.0f) atomicAdd(a.gb2 + m, v[0]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RGB_B_WORKERS / 32) {
    __syncwarp();
    tmem_dealloc(tmem, 512);
  }
}

static size_t rgb_bwd_smem() {
  return (size_t)2 * RGB_G_PART + 2 * 4096 + 4096 + 3 * RGB_H * 4 + (size_t)RGB_B_NSLOT * RGB_B_SLOT + (2 * RGB_B_NSLOT + 2) * 8 + 16;
}

static int rgb_smem_optin() { return smem_optin_bytes(); }
static size_t rgb_fwd_smem(const RgbShape& S) {
  return (size_t)3 * 128 * S.K0p * 2 + (size_t)RGB_NSLOT * RGB_SLOT + 3 * RGB_W3_PART + (2 * RGB_NSLOT + 3) * 8 + 16;
}
static int g_rgb_fused = 1;

}  // namespace ffb

using namespace ffb;

extern "C" {

int ffb_set_fused_rgbmlp(int enabled) {
  g_rgb_fused = enabled ? 1 : 0;
  return FFB_OK;
}

int64_t ffb_rgbmlp_workspace_bytes(int32_t Cf, int32_t hidden, int32_t view_pe, int32_t fea_pe) {
  RgbShape S;
  if (!g_rgb_fused || !ffb_tensor_cores_enabled() || hidden != RGB_H || !rgb_shape(Cf, view_pe, fea_pe, &S)) return 0;
  if (rgb_fwd_smem(S) > (size_t)rgb_smem_optin()) return 0;
  return (int64_t)rgb_pack_bytes(S);
}

int ffb_rgbmlp_pack(const float* W1, const float* b1, const float* W2, const float* W3, void* workspace, int32_t Cf, int32_t view_pe,
                    int32_t fea_pe, void* stream) {
  RgbShape S;
  FFB_REQUIRE(W1 && b1 && W2 && W3 && workspace, "null argument");
  FFB_REQUIRE(rgb_shape(Cf, view_pe, fea_pe, &S), "appearance-MLP shape not eligible for the fused tensor-core path");
  FFB_REQUIRE(((uintptr_t)workspace & 15) == 0, "workspace must be 16-byte aligned");
  rgb_pack_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(W1, b1, W2, W3, (uint8_t*)workspace, S);
  FFB_LAUNCHED();
  rgb_pack_bwd_kernel<<<64, 256, 0, (cudaStream_t)stream>>>(W1, W2, (uint8_t*)workspace + rgb_pack_fwd_bytes(S), S);
  FFB_LAUNCHED();
  return FFB_OK;
}

int ffb_rgbmlp_fwd(const float* feat, int32_t ld_feat, const float* rays, const int32_t* ray_id, const int32_t* app_idx,
                   const void* workspace, const float* b2, float* rgb, uint16_t* relu_bits, float* x_out, float* h1_out, float* h2_out,
                   void* stream_x, void* stream_h1, void* stream_h2, int64_t n, const int32_t* n_dev, int32_t Cf, int32_t view_pe,
                   int32_t fea_pe, void* stream) {
  RgbFwdArgs a;
  FFB_REQUIRE(feat && rays && workspace && b2 && rgb, "null argument");
  FFB_REQUIRE(rgb_shape(Cf, view_pe, fea_pe, &a.S), "appearance-MLP shape not eligible for the fused tensor-core path");
  FFB_REQUIRE(ld_feat >= Cf + 1, "feature rows are narrower than 1 + Cf");
  if (n <= 0) return FFB_OK;
  a.feat = feat; a.ld_feat = ld_feat; a.rays = rays; a.ray_id = ray_id; a.app_idx = app_idx;
  a.wpack = (const uint8_t*)workspace; a.b2 = b2; a.rgb = rgb; a.bits = relu_bits;
  a.x_out = x_out; a.h1_out = h1_out; a.h2_out = h2_out; a.n = n; a.n_dev = n_dev;
  a.sx = (uint8_t*)stream_x; a.sh1 = (uint8_t*)stream_h1; a.sh2 = (uint8_t*)stream_h2;
  FFB_REQUIRE((!stream_x && !stream_h1 && !stream_h2) || (stream_x && stream_h1 && stream_h2), "pass all three activation streams or none");
  const size_t smem = rgb_fwd_smem(a.S);
  FFB_REQUIRE(smem <= (size_t)rgb_smem_optin(), "not enough shared memory for the fused appearance MLP");
  static PerDeviceOnce attr_done;
  if (attr_done.first()) {
    FFB_CUDA(cudaFuncSetAttribute(rgb_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rgb_smem_optin()));
  }
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  rgb_fwd_kernel<<<grid, RGB_THREADS, smem, (cudaStream_t)stream>>>(a);
  FFB_LAUNCHED();
  return FFB_OK;
}

int64_t ffb_rgbmlp_stream_bytes(int32_t Cf, int32_t view_pe, int32_t fea_pe, int64_t n, int32_t which) {
  RgbShape S;
  if (!rgb_shape(Cf, view_pe, fea_pe, &S) || n < 0) return 0;
  const int64_t tiles = (n + 127) / 128;
  return tiles * (int64_t)rgb_stream_tile(which == 0 ? S.K0p : RGB_H);
}

int ffb_rgbmlp_bwd(const float* g_rgb, const float* rgb, const uint16_t* relu_bits, const void* stream_x, const void* stream_h1,
                   const void* stream_h2, const void* workspace, const float* W3, float* g_x, int32_t ld_gx, float* gW1, float* gb1,
                   float* gW2, float* gb2, float* gW3, int64_t n, const int32_t* n_dev, int32_t Cf, int32_t view_pe, int32_t fea_pe,
                   void* stream) {
  RgbBwdArgs a;
  FFB_REQUIRE(g_rgb && rgb && relu_bits && stream_x && stream_h1 && stream_h2 && workspace && W3 && g_x, "null argument");
  FFB_REQUIRE(rgb_shape(Cf, view_pe, fea_pe, &a.S), "appearance-MLP shape not eligible for the fused tensor-core path");
  FFB_REQUIRE(a.S.K0p > 128 && a.S.K0p <= 208, "fused backward expects 128 < padded input width <= 208");
  FFB_REQUIRE(ld_gx >= a.S.K0 && (ld_gx & 3) == 0 && ((uintptr_t)g_x & 15) == 0, "g_x rows must be 16-byte aligned (ld_gx a multiple of 4, >= K0)");
  if (n <= 0) return FFB_OK;
  a.g_rgb = g_rgb; a.rgb = rgb; a.bits = relu_bits;
  a.sx = (const uint8_t*)stream_x; a.sh1 = (const uint8_t*)stream_h1; a.sh2 = (const uint8_t*)stream_h2;
  a.wpack = (const uint8_t*)workspace + rgb_pack_fwd_bytes(a.S); a.W3 = W3; a.g_x = g_x; a.ld_gx = ld_gx;
  a.gW1 = gW1; a.gb1 = gb1; a.gW2 = gW2; a.gb2 = gb2; a.gW3 = gW3; a.n = n; a.n_dev = n_dev;
  const size_t smem = rgb_bwd_smem();
  static PerDeviceOnce attr_done;
  if (attr_done.first()) {
    FFB_CUDA(cudaFuncSetAttribute(rgb_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rgb_smem_optin()));
  }
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  rgb_bwd_kernel<<<grid, RGB_B_THREADS, smem, (cudaStream_t)stream>>>(a);
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
