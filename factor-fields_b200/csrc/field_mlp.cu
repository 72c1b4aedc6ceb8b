// field_mlp.cu — the field query fused with `linear_mat` (K1 + K3 of DESIGN.md in one launch per direction).
//
// Replaces, in one persistent warp-specialised kernel per direction (one CTA of 1024 threads per SM):
//   forward : grid_mapping + 7x grid_sample + cat + mul (FactorFields.py:425-533)  ->  MLPMixer 18->64->32 (:144-159)
//   backward: MLPMixer autograd  ->  7x grid_sampler_backward (atomic scatter)
// The separate kernels left both halves of the SM idle in turn: the gather is bound by the latency of L2-resident loads
// (issue slots 46 % busy, tensor pipe 0 %), the MLP by its serial MMA round trips (LSU / L2 idle), and a 72 B/query feature
// row went to HBM and back between them.  Here
//   * 27 GATHER warps (one lane per query, consecutive lanes = consecutive samples of a ray, exactly the arithmetic of
//     field_fast.cuh) write each feature row straight into a shared-memory operand tile — bf16 x3 split, canonical
//     no-swizzle UMMA layout — of a ring of 128-query tile slots; the coefficient / basis rows the backward pass needs leave
//     as coalesced column-blocked stores;
//   * one MMA warp issues layer 1 (tcgen05.mma, fp32 accumulators in TMEM, double-buffered) as soon as the four 32-row
//     chunks of a slot have arrived (mbarrier), and frees the slot with tcgen05.commit;
//   * four EPILOGUE warps (one per TMEM lane quarter) turn the layer-1 accumulator into the ReLU'd hidden tile (+ decision
//     bits), issue layer 2, and store the [n, 32] output — while the gather warps are already many tiles ahead.
// Nothing but x, y, the ReLU bits and the two saved rows touches HBM.
//
// Precision is that of mlp_fused.cu: 3 bf16 parts / 6 MMAs per product in the forward pass (~3e-7 relative), layer-1 bias
// riding in the GEMM as an all-ones input column.
#include "field_fast.cuh"
#include "tc_tiles.cuh"

namespace ffb {

constexpr int FM_THREADS = 1024;
constexpr int FM_NG = 27;             // gather warps 0..26
constexpr int FM_MMA_WARP = 27;       // layer-1 issuer, owns the TMEM allocation
constexpr int FM_EPI_WARP0 = 28;      // warps 28..31 -> TMEM lane quarters 0..3 (= warp % 4)
constexpr int FM_TERMS = 3;
constexpr uint32_t FM_SC = 2048;      // bytes between 8-column chunks of a 128-row activation tile
constexpr uint32_t FM_XCHUNKS = 3;    // real chunks of an x tile (columns 0..23); chunk 3 (24..31) is the shared zero block
constexpr uint32_t FM_XPART = FM_XCHUNKS * FM_SC;          // 6 KB per bf16 part
constexpr uint32_t FM_XSLOT = FM_TERMS * FM_XPART;         // 18 KB per slot
constexpr uint32_t FM_HPART = 8 * FM_SC;                   // hidden tile: 64 columns
constexpr int FM_H = 64, FM_K0P = 32, FM_NP = 32;

template <int FM_NSLOT>      // ring of x-tile slots: 27 chunks in flight span < 8 tiles, so with 8 slots a claim never waits
struct FmSmemT {
  // byte offsets into dynamic shared memory
  static constexpr uint32_t W1 = 0;                                   // 3 x 4 KB
  static constexpr uint32_t W2 = W1 + FM_TERMS * FM_H * FM_K0P * 2;    // 3 x 4 KB
  static constexpr uint32_t X = W2 + FM_TERMS * FM_NP * FM_H * 2;      // FM_NSLOT x 18 KB
  static constexpr uint32_t ZERO = X + FM_NSLOT * FM_XSLOT;            // 2 KB of zeros: chunk 3 of every x tile / part
  static constexpr uint32_t HID = ZERO + FM_SC;                        // 3 x 16 KB
  static constexpr uint32_t BAR = HID + FM_TERMS * FM_HPART;           // mbarriers
  static constexpr uint32_t N_BAR = 2 * FM_NSLOT + 2 + 2 + 1;          // slot_full, slot_free, d1_full[2], d1_free[2], d2_full
  static constexpr uint32_t MISC = BAR + N_BAR * 8;                    // tmem slot, chunk counter
  static constexpr uint32_t TOTAL = MISC + 16;
};

__device__ __forceinline__ void fm_store2(uint8_t* xrow, int c, float a, float b) {
  uint32_t w[FM_TERMS];
  split2_packed<FM_TERMS>(a, b, w);
  uint8_t* p = xrow + (uint32_t)(c >> 3) * FM_SC + (uint32_t)(c & 7) * 2u;
#pragma unroll
  for (int t = 0; t < FM_TERMS; ++t) *reinterpret_cast<uint32_t*>(p + (uint32_t)t * FM_XPART) = w[t];
}

struct FmFwdArgs {
  FastParams P;
  const float* x;
  int64_t n;
  const int32_t* n_dev;
  const float *W1, *b1, *W2;
  float* y;             // [n, N]
  uint16_t* bits;       // [n, 4] ReLU decisions (may be null)
  float* coeff_blk;     // blocked by 32 rows (blk_idx), may be null
  float* basis_blk;     // blocked by 32 rows, may be null
  float* feats;         // optional row-major [n, W] copy of the feature row (tests / callers without the fused backward)
  float* coeff;         // optional row-major [n, W] coefficient row (get_coding's second output)
  Mlp2Shape S;
  int n_gather;         // gather warps actually used (<= FM_NG and <= 4 * slots: a claim may run at most one use of a slot ahead)
  int debug;            // experiment knob "fm_debug": bit 0 = gather warps skip the field arithmetic, bit 1 = epilogue warps skip their work
};

template <int DB, int DC, int FM_NSLOT>
__global__ void __launch_bounds__(FM_THREADS, 1) field_mlp_fwd_kernel(const FmFwdArgs a) {
  using FmSmem = FmSmemT<FM_NSLOT>;
  extern __shared__ __align__(128) uint8_t smem[];
  const FastParams& P = a.P;
  const int64_t n = resolve_n(a.n, a.n_dev);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sW1 = smem + FmSmem::W1;
  uint8_t* sW2 = smem + FmSmem::W2;
  uint8_t* sX = smem + FmSmem::X;
  uint8_t* sH = smem + FmSmem::HID;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FmSmem::BAR);
  uint64_t* slot_full = bars;
  uint64_t* slot_free = bars + FM_NSLOT;
  uint64_t* d1_full = bars + 2 * FM_NSLOT;
  uint64_t* d1_free = d1_full + 2;
  uint64_t* d2_full = d1_free + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + FmSmem::MISC);
  int* next_chunk = reinterpret_cast<int*>(smem + FmSmem::MISC + 4);
  const int W = P.W;

  // ---- one-time setup: weights -> operand tiles, x slots zeroed with the all-ones bias column in place, barriers, TMEM
  if (warp == FM_MMA_WARP) tmem_alloc(tmem_slot, 256u);
  if (tid == 0) {
    for (int s = 0; s < FM_NSLOT; ++s) {
      mbar_init(slot_full + s, 4);      // the four 32-row chunks of a tile
      mbar_init(slot_free + s, 1);      // tcgen05.commit of layer 1
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(d1_full + b, 1);
      mbar_init(d1_free + b, 128);      // every epilogue thread, after its TMEM loads
    }
    mbar_init(d2_full, 1);
    *next_chunk = 0;
  }
  stage_weights<FM_TERMS>(a.S, a.W1, a.b1, a.W2, sW1, sW2, tid, FM_THREADS);
  for (uint32_t o = tid * 16u; o < FM_NSLOT * FM_XSLOT + FM_SC; o += FM_THREADS * 16u) *reinterpret_cast<uint4*>(sX + o) = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int it = tid; it < FM_NSLOT * 128; it += FM_THREADS) {    // x[:, W] = 1 (bf16 part 0): the bias column of layer 1
    const int s = it >> 7, r = it & 127;
    uint8_t* p = sX + (uint32_t)s * FM_XSLOT + (uint32_t)(W >> 3) * FM_SC + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u + (uint32_t)(W & 7) * 2u;
    *reinterpret_cast<uint16_t*>(p) = 0x3F80u;
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  const int64_t n_tiles = (n + 127) >> 7;
  const int64_t Tc = n_tiles > (int64_t)blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;   // tiles of this CTA

  if (warp < a.n_gather) {
    // =================================== gather warps ===================================
    const float msize = fast_msize(P);
    const int n_chunks = (int)(4 * Tc);
    for (;;) {
      int j = 0;
      if (lane == 0) j = atomicAdd(next_chunk, 1);
      j = __shfl_sync(0xffffffffu, j, 0);
      if (j >= n_chunks) break;
      const int tseq = j >> 2, rg = j & 3, slot = tseq % FM_NSLOT;
      const int64_t i = (((int64_t)blockIdx.x + (int64_t)tseq * gridDim.x) << 7) + rg * 32 + lane;
      const bool active = i < n;
      mbar_wait(slot_free + slot, (uint32_t)(((tseq / FM_NSLOT) & 1) ^ 1));     // layer 1 of the slot's previous tile has read it
      const int r = rg * 32 + lane;
      uint8_t* xrow = sX + (uint32_t)slot * FM_XSLOT + (uint32_t)(r >> 3) * TILE_SR + (uint32_t)(r & 7) * 16u;
      if (active && !(a.debug & 1)) {
        float xr[3];
        for (int d = 0; d < P.xdim; ++d) xr[d] = a.x[i * P.xdim + d];
        TapSet<DC, false> tc;
        coeff_taps<DC, false>(P, xr, tc);
        float* frow = a.feats ? a.feats + i * W : nullptr;
        float* crow = a.coeff ? a.coeff + i * W : nullptr;
        for (int l = 0; l < P.n_levels; ++l) {
          const FastLevel L = P.lv[l];
          TapSet<DB, false> tb;
          basis_taps<DB, false>(P, L, xr, msize, tb);
          if ((L.C & 3) == 0) {
            for (int c0 = 0; c0 < L.C; c0 += 4) {
              float b[4], ca[2], cb[2];
              gather_vec<DB, false, 4>(L.data, L.C, c0, tb, b);
              gather_vec<DC, false, 2>(P.cdata, W, L.col + c0, tc, ca);
              gather_vec<DC, false, 2>(P.cdata, W, L.col + c0 + 2, tc, cb);
              const int o = L.col + c0;
              const float f0 = b[0] * ca[0], f1 = b[1] * ca[1], f2 = b[2] * cb[0], f3 = b[3] * cb[1];
              fm_store2(xrow, o, f0, f1);
              fm_store2(xrow, o + 2, f2, f3);
              if (a.coeff_blk) {
                a.coeff_blk[blk_idx(i, o, W)] = ca[0]; a.coeff_blk[blk_idx(i, o + 1, W)] = ca[1];
                a.coeff_blk[blk_idx(i, o + 2, W)] = cb[0]; a.coeff_blk[blk_idx(i, o + 3, W)] = cb[1];
              }
              if (a.basis_blk) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a.basis_blk[blk_idx(i, o + q, W)] = b[q];
              }
              if (frow) {
                *reinterpret_cast<float2*>(frow + o) = make_float2(f0, f1);
                *reinterpret_cast<float2*>(frow + o + 2) = make_float2(f2, f3);
              }
              if (crow) {
                *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
                *reinterpret_cast<float2*>(crow + o + 2) = make_float2(cb[0], cb[1]);
              }
            }
          } else {
            for (int c0 = 0; c0 < L.C; c0 += 2) {
              float b[2], ca[2];
              gather_vec<DB, false, 2>(L.data, L.C, c0, tb, b);
              gather_vec<DC, false, 2>(P.cdata, W, L.col + c0, tc, ca);
              const int o = L.col + c0;
              const float f0 = b[0] * ca[0], f1 = b[1] * ca[1];
              fm_store2(xrow, o, f0, f1);
              if (a.coeff_blk) { a.coeff_blk[blk_idx(i, o, W)] = ca[0]; a.coeff_blk[blk_idx(i, o + 1, W)] = ca[1]; }
              if (a.basis_blk) { a.basis_blk[blk_idx(i, o, W)] = b[0]; a.basis_blk[blk_idx(i, o + 1, W)] = b[1]; }
              if (frow) *reinterpret_cast<float2*>(frow + o) = make_float2(f0, f1);
              if (crow) *reinterpret_cast<float2*>(crow + o) = make_float2(ca[0], ca[1]);
            }
          }
        }
      } else {
        for (int c = 0; c < W; c += 2) fm_store2(xrow, c, 0.0f, 0.0f);     // rows past n: defined (zero) operands
      }
      proxy_fence();                    // generic-proxy writes of this lane -> visible to the tensor-core (async) proxy
      __syncwarp();
      if (lane == 0) mbar_arrive(slot_full + slot);
    }
  } else if (warp < FM_NG) {
    // idle (experiment configurations with fewer gather warps)
  } else if (warp == FM_MMA_WARP) {
    // =================================== layer-1 issuer ===================================
    if (elect_one()) {
      const uint32_t idesc1 = make_idesc(FM_H, 0, 0);
      const uint32_t aW1 = smem_u32(sW1), aZero = smem_u32(smem + FmSmem::ZERO);
      const uint32_t szW1 = FM_H * FM_K0P * 2, scW1 = FM_H * 16;
      for (int64_t t = 0; t < Tc; ++t) {
        const int slot = (int)(t % FM_NSLOT), b = (int)(t & 1);
        mbar_wait(slot_full + slot, (uint32_t)((t / FM_NSLOT) & 1));
        mbar_wait(d1_free + b, (uint32_t)(((t >> 1) & 1) ^ 1));
        tc_fence_after();
        const uint32_t aX = smem_u32(sX) + (uint32_t)slot * FM_XSLOT;
        const uint32_t d1 = tmem + (uint32_t)b * FM_H;
        uint32_t acc = 0;
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
          for (int ta = 0; ta < FM_TERMS; ++ta) {
#pragma unroll
            for (int tb = 0; tb < FM_TERMS; ++tb) {
              if (ta + tb >= FM_TERMS) continue;
              const uint32_t xa = aX + (uint32_t)ta * FM_XPART + (uint32_t)ks * 2u * FM_SC;
              // K slice 0: chunks 0,1 (LBO = chunk stride).  K slice 1: chunk 2 and, as its K-neighbour, the zero block.
              const uint64_t da = ks == 0 ? make_desc(xa, FM_SC, TILE_SR) : make_desc(xa, aZero - xa, TILE_SR);
              umma_f16(d1, da, desc_k(aW1 + (uint32_t)tb * szW1, scW1, ks), idesc1, acc);
              acc = 1u;
            }
          }
        }
        umma_commit(slot_free + slot);
        umma_commit(d1_full + b);
      }
    }
    __syncwarp();
  } else {
    // =================================== epilogue warps ===================================
    const int q = warp & 3, row = q * 32 + lane;
    const uint32_t lane_base = (uint32_t)(q * 32) << 16;
    const uint32_t d2 = tmem + 2u * FM_H;
    const uint32_t idesc2 = make_idesc(FM_NP, 0, 0);
    const uint32_t aH = smem_u32(sH), aW2 = smem_u32(sW2);
    const uint32_t szW2 = FM_NP * FM_H * 2, scW2 = FM_NP * 16;
    const int N = a.S.N;
    auto epi2 = [&](int64_t t) {      // y rows of tile t from the layer-2 accumulator
      mbar_wait(d2_full, (uint32_t)(t & 1));
      tc_fence_after();
      const int64_t grow = (((int64_t)blockIdx.x + t * gridDim.x) << 7) + row;
#pragma unroll
      for (int c0 = 0; c0 < FM_NP; c0 += 16) {
        if (a.debug & 2) break;
        float v[16];
        tmem_ld16(d2 + lane_base + (uint32_t)c0, v);
        if (grow < n) {
          float* yr = a.y + grow * N + c0;
          if ((N & 3) == 0 && c0 + 16 <= N) {
#pragma unroll
            for (int k = 0; k < 16; k += 4) *reinterpret_cast<float4*>(yr + k) = make_float4(v[k], v[k + 1], v[k + 2], v[k + 3]);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (c0 + k < N) yr[k] = v[k];
          }
        }
      }
    };
    for (int64_t t = 0; t < Tc; ++t) {
      const int b = (int)(t & 1);
      if (t > 0) epi2(t - 1);          // also: layer 2 of tile t-1 has finished reading the hidden tile
      mbar_wait(d1_full + b, (uint32_t)((t >> 1) & 1));
      tc_fence_after();
      const int64_t grow = (((int64_t)blockIdx.x + t * gridDim.x) << 7) + row;
      const uint32_t d1 = tmem + (uint32_t)b * FM_H;
#pragma unroll
      for (int c0 = 0; c0 < FM_H; c0 += 16) {
        if (a.debug & 2) break;
        float v[16];
        tmem_ld16(d1 + lane_base + (uint32_t)c0, v);
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          bits |= (v[k] > 0.0f ? 1u : 0u) << k;
          v[k] = fmaxf(v[k], 0.0f);
        }
        if (a.bits && grow < n) a.bits[grow * (FM_H >> 4) + (c0 >> 4)] = (uint16_t)bits;
        store_row8<FM_TERMS>(sH, FM_HPART, FM_SC, row, c0, v);
        store_row8<FM_TERMS>(sH, FM_HPART, FM_SC, row, c0 + 8, v + 8);
      }
      tc_fence_before();
      mbar_arrive(d1_free + b);
      proxy_fence();
      named_sync(1, 128);
      if (warp == FM_EPI_WARP0 && elect_one()) {
        tc_fence_after();
        issue_gemm<FM_TERMS>(d2, idesc2, FM_H / 16, false, [&](int tt, int s) { return desc_k(aH + (uint32_t)tt * FM_HPART, FM_SC, s); },
                             [&](int tt, int s) { return desc_k(aW2 + (uint32_t)tt * szW2, scW2, s); });
        umma_commit(d2_full);
      }
    }
    if (Tc > 0) epi2(Tc - 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == FM_MMA_WARP) tmem_dealloc(tmem, 256u);
}

static bool fm_shape_ok(const FastParams& P, int K0, int H, int N, Mlp2Shape* S) {
  if (H != FM_H || K0 != P.W || (P.W & 1) || P.W + 1 > 24 || N < 1 || N > FM_NP) return false;
  S->K0 = K0; S->H = H; S->N = N; S->K0p = FM_K0P; S->Np = FM_NP;
  return true;
}

}  // namespace ffb

using namespace ffb;

extern "C" {

static int g_fm_enabled = 1;
static int g_fm_debug = 0, g_fm_nslot = 8;
int ffb_field_mlp_tuning(int debug, int nslot) { g_fm_debug = debug; g_fm_nslot = nslot; return FFB_OK; }

int ffb_set_field_mlp(int enabled) {
  g_fm_enabled = enabled ? 1 : 0;
  return FFB_OK;
}

/* 1 when get_coding + linear_mat of this field can run as the fused kernels: grid x grid field with linear taps (the
 * nerf.yaml / sdf.yaml shape class, 2-D or 3-D), W = K0 <= 22 feature columns, hidden width 64, at most 32 outputs. */
int ffb_field_mlp_eligible(ffb_field_t f, int32_t K0, int32_t H, int32_t N) {
  if (!f || !g_fm_enabled || !ffb_tensor_cores_enabled()) return 0;
  FastParams P;
  int idx[FAST_MAX_LEVELS + 1];
  if (!build_params(f->h, P, idx)) return 0;
  const bool nb = f->h.ops[f->h.bterms[0].op[0]].nearest, nc = f->h.ops[f->h.cterms[0].op[0]].nearest;
  if (nb || nc) return 0;
  if (!((P.in_dim == 3 && P.xdim == 3) || (P.in_dim == 2 && P.xdim == 2))) return 0;
  Mlp2Shape S;
  if (!fm_shape_ok(P, K0, H, N, &S)) return 0;
  return smem_optin_bytes() >= (int)FmSmemT<8>::TOTAL ? 1 : 0;
}

int ffb_field_mlp_fwd(ffb_field_t f, const float* x, int64_t n, const int32_t* n_dev, const float* W1, const float* b1, const float* W2,
                      float* y, uint16_t* relu_bits, float* coeff_blk, float* basis_blk, float* feats, float* coeff, int32_t K0, int32_t H,
                      int32_t N, void* stream) {
  FFB_REQUIRE(f && x && W1 && b1 && W2 && y, "null argument");
  FFB_REQUIRE(ffb_field_mlp_eligible(f, K0, H, N) == 1, "field / MLP shape is not eligible for the fused field + linear_mat kernels");
  if (n <= 0) return FFB_OK;
  FmFwdArgs a;
  int idx[FAST_MAX_LEVELS + 1];
  build_params(f->h, a.P, idx);
  fm_shape_ok(a.P, K0, H, N, &a.S);
  a.x = x; a.n = n; a.n_dev = n_dev; a.W1 = W1; a.b1 = b1; a.W2 = W2; a.y = y; a.bits = relu_bits;
  a.coeff_blk = coeff_blk; a.basis_blk = basis_blk; a.feats = feats; a.coeff = coeff;
  const int64_t tiles = (n + 127) / 128;
  const unsigned grid = (unsigned)(tiles < sm_count() ? tiles : sm_count());
  cudaStream_t s = (cudaStream_t)stream;
  a.debug = g_fm_debug;
  a.n_gather = g_fm_nslot * 4 < FM_NG ? g_fm_nslot * 4 : FM_NG;
#define FM_LAUNCH(DB_, DC_, NS_)                                                                                                    \
  do {                                                                                                                              \
    static PerDeviceOnce once;                                                                                                      \
    if (once.first())                                                                                                               \
      FFB_CUDA(cudaFuncSetAttribute(field_mlp_fwd_kernel<DB_, DC_, NS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FmSmemT<NS_>::TOTAL)); \
    field_mlp_fwd_kernel<DB_, DC_, NS_><<<grid, FM_THREADS, FmSmemT<NS_>::TOTAL, s>>>(a);                                           \
  } while (0)
  if (a.P.in_dim == 3) {
    if (g_fm_nslot == 4) FM_LAUNCH(3, 3, 4);
    else if (g_fm_nslot == 6) FM_LAUNCH(3, 3, 6);
    else FM_LAUNCH(3, 3, 8);
  } else {
    FM_LAUNCH(2, 2, 8);
  }
  FFB_LAUNCHED();
  return FFB_OK;
}

}  // extern "C"
