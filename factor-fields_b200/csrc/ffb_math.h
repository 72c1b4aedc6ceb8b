// ffb_math.h — coordinate arithmetic shared by every kernel.  Host+device inline so that the exact
// same expressions can be unit-tested on the CPU (tests/ builds a tiny host shim around this file).
//
// Everything that feeds a DECISION (floor / nearest index / in-box test / mask threshold) is written
// with explicit round-to-nearest single operations (no FMA contraction), reproducing the reference's
// unfused eager-PyTorch op order.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FFB_HD __host__ __device__ __forceinline__
#else
#define FFB_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define FFB_MUL(a, b) __fmul_rn((a), (b))
#define FFB_ADD(a, b) __fadd_rn((a), (b))
#define FFB_SUB(a, b) __fsub_rn((a), (b))
#define FFB_DIV(a, b) __fdiv_rn((a), (b))
#else
// host build: compile with -ffp-contract=off
static inline float ffb_vol(float x) { volatile float y = x; return y; }
#define FFB_MUL(a, b) ffb_vol((a) * (b))
#define FFB_ADD(a, b) ffb_vol((a) + (b))
#define FFB_SUB(a, b) ffb_vol((a) - (b))
#define FFB_DIV(a, b) ffb_vol((a) / (b))
#endif

namespace ffb {

// torch.remainder for floats (ATen BinaryOpsKernel remainder): fmod, then shift into the sign of b.
FFB_HD float remainder_f(float a, float b) {
  float m = fmodf(a, b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) m = FFB_ADD(m, b);
  return m;
}

// torch floor_divide for floats (ATen div_floor_floating).
FFB_HD float floordiv_f(float a, float b) {
  if (b == 0.f) return FFB_DIV(a, b);
  float m = fmodf(a, b);
  float d = FFB_DIV(FFB_SUB(a, m), b);
  if (m != 0.f && ((b < 0.f) != (m < 0.f))) d = FFB_SUB(d, 1.f);
  float fd;
  if (d != 0.f) {
    fd = floorf(d);
    if (FFB_SUB(d, fd) > 0.5f) fd = FFB_ADD(fd, 1.f);
  } else {
    fd = copysignf(0.f, FFB_DIV(a, b));
  }
  return fd;
}

// FactorFields.normalize_coord (FactorFields.py:635-637): (x - aabb0) * (2 / (aabb1 - aabb0)) - 1
FFB_HD float normalize_coord(float x, float lo, float hi) {
  float inv = FFB_DIV(2.0f, FFB_SUB(hi, lo));
  return FFB_SUB(FFB_MUL(FFB_SUB(x, lo), inv), 1.0f);
}

// grid_mapping (FactorFields.py:11-33) for one coordinate and one frequency band.
// `scale` = max(aabbSize) / freq (fp32 division done by the caller once per band).
// For FFB_MAP_TRIG returns sin; cos is returned through *aux.
FFB_HD float map_coord(float x, float lo, float scale, int mapping, float* aux) {
  const float p = FFB_SUB(x, lo);
  const float kPi = 3.14159274101257324219f;      // float32(np.pi)
  const float kHalfPi = 1.57079637050628662109f;  // float32(np.pi / 2)
  switch (mapping) {
    case 0: {  // sawtooth
      float l = remainder_f(p, scale);
      l = FFB_SUB(FFB_DIV(l, FFB_DIV(scale, 2.0f)), 1.0f);
      return fminf(fmaxf(l, -1.0f), 1.0f);
    }
    case 1: {  // triangle
      float l = remainder_f(p, scale);
      float li = remainder_f(floordiv_f(p, scale), 2.0f);
      l = FFB_SUB(FFB_DIV(l, FFB_DIV(scale, 2.0f)), 1.0f);
      return (li == 1.0f) ? -l : l;
    }
    case 2:  // sinc
      return sinf(FFB_SUB(FFB_DIV(p, FFB_DIV(scale, kPi)), kHalfPi));
    case 3: {  // trigonometric
      float a = FFB_MUL(FFB_MUL(FFB_DIV(p, scale), 2.0f), kPi);
      if (aux) *aux = cosf(a);
      return sinf(a);
    }
    default:  // 'x'
      return FFB_DIV(p, scale);
  }
}

// ATen grid_sampler_unnormalize + clip_coordinates (padding_mode='border').
FFB_HD float source_index(float u, int size, int align_corners, int border) {
  float c;
  if (align_corners)
    c = FFB_MUL(FFB_DIV(FFB_ADD(u, 1.0f), 2.0f), (float)(size - 1));
  else
    c = FFB_DIV(FFB_SUB(FFB_MUL(FFB_ADD(u, 1.0f), (float)size), 1.0f), 2.0f);
  if (border) c = fminf((float)(size - 1), fmaxf(c, 0.0f));
  return c;
}

// One axis of a linear tap: low index, weight of the low and of the high corner (ATen: (ix_hi - ix), (ix - ix_lo)).
struct Axis {
  int i0;
  float w0, w1;
};
FFB_HD Axis linear_axis(float c) {
  Axis a;
  float f = floorf(c);
  a.i0 = (int)f;
  a.w0 = FFB_SUB(FFB_ADD(f, 1.0f), c);
  a.w1 = FFB_SUB(c, f);
  return a;
}
FFB_HD int nearest_index(float c) { return (int)nearbyintf(c); }  // round-half-even, as ATen

// sample_point (FactorFields.py:588-591): entry distance of a ray into the box.
FFB_HD float ray_tmin(const float o[3], const float d[3], const float lo[3], const float hi[3]) {
  float t = -INFINITY;
  for (int k = 0; k < 3; ++k) {
    float v = (d[k] == 0.0f) ? 1e-6f : d[k];
    float ra = FFB_DIV(FFB_SUB(hi[k], o[k]), v);
    float rb = FFB_DIV(FFB_SUB(lo[k], o[k]), v);
    t = fmaxf(t, fminf(ra, rb));
  }
  return fminf(fmaxf(t, 0.05f), 1e3f);
}

// interpx for sample s (FactorFields.py:592-597); jitter < 0 flags is_train=False (no add).
FFB_HD float sample_t(float tmin, float step_size, int s, float jitter, bool train) {
  float r = (float)s;
  if (train) r = FFB_ADD(r, jitter);
  return FFB_ADD(tmin, FFB_MUL(step_size, r));
}

FFB_HD bool sample_pos(const float o[3], const float d[3], float t, const float lo[3], const float hi[3], float p[3]) {
  bool inside = true;
  for (int k = 0; k < 3; ++k) {
    p[k] = FFB_ADD(o[k], FFB_MUL(d[k], t));
    inside = inside && !(lo[k] > p[k] || p[k] > hi[k]);
  }
  return inside;
}

// sample_point_unbound (FactorFields.py:625-633): p = o + d t; outside the unit cube (inf-norm > 1) the point is
// contracted to p / n * ((1 + bg) - bg / n).  Returns inner_mask.
FFB_HD bool sample_pos_unbound(const float o[3], const float d[3], float t, float bg_len, float p[3]) {
  float n = 0.0f;
  for (int k = 0; k < 3; ++k) {
    p[k] = FFB_ADD(o[k], FFB_MUL(d[k], t));
    n = fmaxf(n, fabsf(p[k]));
  }
  if (n <= 1.0f) return true;
  // `self.bg_len / norm` on a Python float is Tensor.__rtruediv__ = norm.reciprocal() * bg_len
  const float s = FFB_SUB((float)(1.0 + (double)bg_len), FFB_MUL(FFB_DIV(1.0f, n), bg_len));
  for (int k = 0; k < 3; ++k) p[k] = FFB_MUL(FFB_DIV(p[k], n), s);
  return false;
}

// AlphaGridMask.sample_alpha (FactorFields.py:103-110) on a 0/1 volume, ATen trilinear order, zeros padding,
// align_corners=True.
FFB_HD float alpha_lookup(const uint8_t* vol, const int size[3] /*W,H,D*/, const float amin[3], const float ainv[3],
                          const float p[3]) {
  Axis ax[3];
  for (int k = 0; k < 3; ++k) {
    float u = FFB_SUB(FFB_MUL(FFB_SUB(p[k], amin[k]), ainv[k]), 1.0f);
    ax[k] = linear_axis(source_index(u, size[k], 1, 0));
  }
  float out = 0.0f;
  for (int c = 0; c < 8; ++c) {
    int bx = c & 1, by = (c >> 1) & 1, bz = c >> 2;
    int ix = ax[0].i0 + bx, iy = ax[1].i0 + by, iz = ax[2].i0 + bz;
    if (ix < 0 || iy < 0 || iz < 0 || ix >= size[0] || iy >= size[1] || iz >= size[2]) continue;
    float w = FFB_MUL(FFB_MUL(bx ? ax[0].w1 : ax[0].w0, by ? ax[1].w1 : ax[1].w0), bz ? ax[2].w1 : ax[2].w0);
    float v = (float)vol[((size_t)iz * size[1] + iy) * size[0] + ix];
    out = FFB_ADD(out, FFB_MUL(v, w));
  }
  return out;
}

// F.softplus(beta=1, threshold=20)
FFB_HD float softplus_f(float x) { return x > 20.0f ? x : log1pf(expf(x)); }
FFB_HD float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }

}  // namespace ffb
