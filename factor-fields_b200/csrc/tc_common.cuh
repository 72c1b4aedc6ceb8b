// tc_common.cuh — tcgen05 / TMEM / mbarrier primitives (inline PTX) and the bf16-split helpers shared by the
// tensor-core MLP kernels (mlp_tc.cu: one layer per launch; mlp_fused.cu: whole decoder MLPs per launch).
//
// Shared-memory operand tiles use the canonical NO-SWIZZLE UMMA layout built from 8x8 bf16 core matrices (128
// contiguous bytes, 16 bytes per row).  For a tile of R rows x C columns, element (r, c) lives at
//     (r/8)*SR + (c/8)*SC + (r%8)*16 + (c%8)*2
// and the SAME bytes serve two roles:
//   * K-major operand  (M/N index = r, K index = c): descriptor LBO = SC, SBO = SR, one K=16 slice = 2*SC bytes;
//   * MN-major operand (M/N index = c, K index = r): descriptor LBO = SR, SBO = SC, one K=16 slice = 2*SR bytes
// (LBO is always the stride between core matrices along K, SBO along M/N), which is what lets the backward kernels
// use one staged activation tile both as the A operand of the input-gradient GEMM and, transposed, as an operand of
// the weight-gradient GEMM.
#pragma once
#include <cuda_bf16.h>
#include "ffb_common.cuh"

namespace ffb {

// One lane of a CONVERGED warp (elect.sync).  The tensor-core / TMA instructions are warp-uniform in SASS (UTCHMMA, UBLKCP, UTCBAR):
// inside an `if (lane == 0)` branch nvcc cannot prove that and wraps EVERY such instruction in an ELECT / BRA.U.ANY waterfall
// loop (~150 cycles per MMA, measured: the issue rate, not the tensor pipe, bounded every MLP kernel); behind elect.sync it
// emits them back to back.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  while (!ok) {
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA-engine bulk copy (no tensor map): `bytes` contiguous bytes global -> shared, completion counted on `bar`
// (16-byte aligned addresses, size a multiple of 16).  The operand tiles it moves are already in UMMA layout.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// named barrier among the first `count` threads' warps (the worker warps of a warp-specialised kernel)
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{ .reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Cheap MMA issue.  A timeline trace of the pipelined kernels (scratch/trace_mlp2p.py) showed the ISSUE of the MMAs — one thread
// rebuilding two 64-bit descriptors per instruction — to be the pipeline's bottleneck (~170 cycles per tcgen05.mma, 36 per
// tile).  A descriptor of a fixed layout differs between the MMAs of a tile only in its 14-bit start-address field, so the
// base is built once and each MMA adds a compile-time constant to the low word.
struct DescBase {
  uint32_t lo, hi;
};
__device__ __forceinline__ DescBase desc_base(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  DescBase d;
  d.lo = ((saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
  d.hi = (sbo >> 4) | (1u << 14);       // bit 46 of the descriptor: version 1 (Blackwell)
  return d;
}
__device__ __forceinline__ uint64_t desc_at(const DescBase d, uint32_t byte_off) {
  return ((uint64_t)d.hi << 32) | (uint64_t)(d.lo + (byte_off >> 4));
}
template <bool ACC>
__device__ __forceinline__ void umma_f16_c(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
  if (ACC)
    asm volatile("{ .reg .pred p; setp.eq.u32 p, 1, 1; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc)
                 : "memory");
  else
    asm volatile("{ .reg .pred p; setp.eq.u32 p, 1, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p; }" ::"r"(tmem_d), "l"(adesc),
                 "l"(bdesc), "r"(idesc)
                 : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32 lanes x 16 consecutive columns (fp32) of this warp's TMEM lane quarter
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float v[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// Shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell): start>>4 | LBO>>4 <<16 | SBO>>4 <<32 | 1<<46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
// Instruction descriptor for kind::f16: D=f32, A=B=bf16, M=128, N=n; major bits: 0 = K-major, 1 = MN-major
__device__ __forceinline__ uint32_t make_idesc(int n, int a_mn, int b_mn, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(n >> 3) << 17) |
         ((uint32_t)(m >> 4) << 24);
}


// x = p0 + p1 (+ p2) + O(2^-9*TERMS x): round-to-nearest bf16 parts of an fp32 value
template <int TERMS>
__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16 out[TERMS]) {
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    out[t] = __float2bfloat16_rn(v);
    v -= __bfloat162float(out[t]);
  }
}

// Two fp32 values -> TERMS packed bf16x2 words (low half = a): one cvt.rn.bf16x2 per part, residuals by integer
// re-expansion of the halves (bf16 -> fp32 is a 16-bit shift) — the same round-to-nearest parts as split_bf16.
template <int TERMS>
__device__ __forceinline__ void split2_packed(float a, float b, uint32_t out[TERMS]) {
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    const __nv_bfloat162 p = __floats2bfloat162_rn(a, b);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&p);
    out[t] = w;
    if (t + 1 < TERMS) {
      a -= __uint_as_float(w << 16);
      b -= __uint_as_float(w & 0xffff0000u);
    }
  }
}
template <int TERMS>
__device__ __forceinline__ void split8_packed(const float x[8], uint4 out[TERMS]) {
  uint32_t w[4][TERMS];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2_packed<TERMS>(x[2 * i], x[2 * i + 1], w[i]);
#pragma unroll
  for (int t = 0; t < TERMS; ++t) out[t] = make_uint4(w[0][t], w[1][t], w[2][t], w[3][t]);
}

// Truncating variant: each part takes the top 16 bits of the running residual (PRMT, no F2FP conversion).  With three parts
// the decomposition of a normal fp32 value is EXACT (8 + 8 + 8 significant bits); with fewer parts the residual is up to
// twice that of the round-to-nearest split, so the gradient kernels (2 parts) keep split2_packed.
template <int TERMS>
__device__ __forceinline__ void split2_trunc(float a, float b, uint32_t out[TERMS]) {
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    const uint32_t ua = __float_as_uint(a), ub = __float_as_uint(b);
    out[t] = __byte_perm(ua, ub, 0x7632);        // low half = high 16 bits of a, high half = high 16 bits of b
    if (t + 1 < TERMS) {
      a -= __uint_as_float(ua & 0xffff0000u);
      b -= __uint_as_float(ub & 0xffff0000u);
    }
  }
}
template <int TERMS>
__device__ __forceinline__ void split8_trunc(const float x[8], uint4 out[TERMS]) {
  uint32_t w[4][TERMS];
#pragma unroll
  for (int i = 0; i < 4; ++i) split2_trunc<TERMS>(x[2 * i], x[2 * i + 1], w[i]);
#pragma unroll
  for (int t = 0; t < TERMS; ++t) out[t] = make_uint4(w[0][t], w[1][t], w[2][t], w[3][t]);
}

// 8 consecutive fp32 values -> one 16-byte chunk (8 bf16) per part
template <int TERMS>
__device__ __forceinline__ void split8_parts(const float x[8], uint4 out[TERMS]) {
  float r[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) r[i] = x[i];
#pragma unroll
  for (int t = 0; t < TERMS; ++t) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat16 a = __float2bfloat16_rn(r[2 * i]), b = __float2bfloat16_rn(r[2 * i + 1]);
      r[2 * i] -= __bfloat162float(a);
      r[2 * i + 1] -= __bfloat162float(b);
      __nv_bfloat162 ab = __halves2bfloat162(a, b);
      w[i] = *reinterpret_cast<uint32_t*>(&ab);
    }
    out[t] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

}  // namespace ffb
