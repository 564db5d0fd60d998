// Fog (Koschmieder blend, data/preprocessing.py:116-123) with an fp32 SCREEN in front of the reference's fp64
// expression.  Per-thread code shared by fog_kernel (csrc/corrupt.cu) and the host emulation of
// tests/test_fog_fast_cpu.py.
//
// The reference evaluates  v = x * t + A * (1 - t),  t = exp(-beta * d),  in fp64 and stores trunc(clip(v, 0, 1) * 255).
// The byte is a step function of y = 255 v; an approximation of y decides it unless y lies within the
// approximation's error of an integer.  The screen computes Y = 2^15 y in fp32,
//     t32 = ex2.approx(fl(c * fl32(d))),  c = fl32(-beta * log2 e)
//     Y   = fma(u * 2^15, t32, fma(-a, t32, a)),  a = fl32(255 * A * 2^15)          (u = the input byte, exact)
// rounds it to an integer q with a magic-number add, takes the byte from bits 15..22 of q and sends the PIXEL to the
// exact path when any of its three values has (q + kBand) mod 2^15 <= 2 kBand, i.e. |y - integer| <= kBand / 2^15 =
// 3.05e-4.  Error budget of y (0 <= t <= 1, 0 <= A <= 1, the launcher checks the parameters and the kernel flags
// pixels whose exponent argument is positive or NaN):
//     t32: argument 3 roundings (d, c, product) -> 1.25e-7 |arg| 2^-|arg| <= 6.6e-8, ex2.approx 2 ulp -> 2.4e-7 t,
//          together <= 3.1e-7 absolute; y moves by |u - 255 A| <= 255 times that                        7.9e-5
//     a (one rounding of 255 A 2^15) 1.5e-5, the two fma roundings 2 x 7.6e-6, the reference's own
//     x = fl32(u / 255) 1.5e-5, the magic-number rounding 1.5e-5                                           6.1e-5
// 1.4e-4 in total against a band of 3.05e-4.  The exact path is the reference's expression as before
// (exp, __dmul_rn / __dadd_rn, no contraction).  ~1.8e-3 of the pixels take it; a thread (4 pixels) loops over its
// flagged pixels, so a warp pays ~110 instructions with probability ~0.2 instead of 75 per pixel always.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "blur_strip.cuh"  // packed pairs (F2, add2, fma2, ...)
#include "convert.cuh"

namespace awx {
namespace fog {

using strip::F2;

constexpr int kBand = 10;
constexpr float kTwo15 = 32768.0f, kTwo23 = 8388608.0f, kTwo38 = 274877906944.0f;

struct Params {
  double neg_beta;   // numpy evaluates (-beta) * depth
  double airlight;   // already rounded to fp32 by the host (A * ones_like(fp32))
  float c;           // fl32(-beta * log2 e)
  float a;           // fl32(255 * airlight * 2^15)
};

AWX_HD Params make_params(double beta, double airlight) {
  Params p;
  p.neg_beta = -beta;
  p.airlight = airlight;
  p.c = (float)(-beta * 1.4426950408889634);
  p.a = (float)(255.0 * airlight * 32768.0);
  return p;
}
// the screen's preconditions: transmission in [0, 1] (checked per pixel through the sign of the exponent) and
// airlight in [0, 1], so that 0 <= Y <= 255 * 2^15 < 2^23
AWX_HD bool params_ok(double beta, double airlight) { return beta >= 0.0 && airlight >= 0.0 && airlight <= 1.0; }

#if defined(__CUDA_ARCH__)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float as_float(unsigned u) { return __uint_as_float(u); }
__device__ __forceinline__ unsigned as_uint(float f) { return __float_as_uint(f); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ int trunc_to_int(double v) { return __double2int_rz(v); }
// byte b of w into the low byte of `hi24` (a float's upper 24 bits): one PRMT
__device__ __forceinline__ unsigned byte_into(unsigned w, int b, unsigned hi24) { return __byte_perm(w, hi24, 0x7650 + b); }
__device__ __forceinline__ unsigned min3u(unsigned a, unsigned b, unsigned c) { return min(min(a, b), c); }
#else
inline float& host_ex2_error() {  // test hook: relative error injected into the host stand-in for ex2.approx
  static float e = 0.0f;
  return e;
}
inline float ex2(float x) { return exp2f(x) * (1.0f + host_ex2_error()); }
inline float as_float(unsigned u) {
  float f;
  memcpy(&f, &u, 4);
  return f;
}
inline unsigned as_uint(float f) {
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
}
inline double dmul(double a, double b) { return a * b; }  // compiled with -ffp-contract=off
inline double dadd(double a, double b) { return a + b; }
inline double dsub(double a, double b) { return a - b; }
inline float fmul(float a, float b) {
  volatile float r = a * b;
  return r;
}
inline int trunc_to_int(double v) { return (int)v; }
inline unsigned byte_into(unsigned w, int b, unsigned hi24) { return (hi24 & 0xffffff00u) | ((w >> (8 * b)) & 0xffu); }
inline unsigned min3u(unsigned a, unsigned b, unsigned c) {
  const unsigned m = a < b ? a : b;
  return m < c ? m : c;
}
#endif

// (np.clip(v, 0, 1) * 255).astype(uint8): truncation toward zero and the clip commute, so the clamp is on integers
AWX_HD unsigned exact_byte(unsigned u, double tr, double veil) {
  const double v = dadd(dmul((double)unit_of_u8(u), tr), veil);
  const int i = trunc_to_int(dmul(v, 255.0));
  return (unsigned)(i < 0 ? 0 : (i > 255 ? 255 : i));
}

// the reference's expression for the three values of one pixel; `k0` = index of the pixel's first byte in w[0..2]
AWX_HD void exact_pixel(const unsigned (&w)[3], unsigned (&o)[3], int k0, double depth, const Params& fp) {
  const double tr = exp(dmul(fp.neg_beta, depth));
  const double veil = dmul(fp.airlight, dsub(1.0, tr));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int k = k0 + c, sh = (k & 3) * 8;
    const unsigned word = k < 4 ? w[0] : (k < 8 ? w[1] : w[2]);
    const unsigned r = exact_byte((word >> sh) & 0xffu, tr, veil) << sh, keep = ~(0xffu << sh);
    if (k < 4)
      o[0] = (o[0] & keep) | r;
    else if (k < 8)
      o[1] = (o[1] & keep) | r;
    else
      o[2] = (o[2] & keep) | r;
  }
}

// Four pixels (12 bytes in w[0..2]) through the screen; returns the mask of the pixels the exact path must redo.
AWX_HD unsigned screen4(const unsigned (&w)[3], const double (&depth)[4], const Params& fp, unsigned (&o)[3]) {
  float tr[4], ve[4];
  unsigned redo = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float arg = fmul((float)depth[j], fp.c);
    tr[j] = ex2(arg);
    ve[j] = fmaf(-fp.a, tr[j], fp.a);
    if (!(arg <= 0.0f)) redo |= 1u << j;  // transmission above 1, or NaN
  }
  unsigned q[12];
  const F2 m38 = strip::splat2(-kTwo38), m23 = strip::splat2(kTwo23);
  constexpr unsigned kHi = 0x52800000u;  // 2^38: a byte in the low mantissa bits weighs 2^15
#pragma unroll
  for (int jj = 0; jj < 2; ++jj) {  // pixel pairs (0, 2) and (1, 3): both halves of a packed operation share the channel
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int k0 = 3 * jj + c, k1 = 3 * (jj + 2) + c;
      const F2 us = strip::add2(strip::pair(as_float(byte_into(w[k0 >> 2], k0 & 3, kHi)), as_float(byte_into(w[k1 >> 2], k1 & 3, kHi))), m38);
      const F2 y = strip::fma2(us, strip::pair(tr[jj], tr[jj + 2]), strip::pair(ve[jj], ve[jj + 2]));
      const F2 z = strip::add2(y, m23);  // round to nearest integer: it sits in the mantissa
      q[k0] = as_uint(z.x);
      q[k1] = as_uint(z.y);
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const unsigned m = min3u((q[3 * j] + kBand) & 0x7fffu, (q[3 * j + 1] + kBand) & 0x7fffu, (q[3 * j + 2] + kBand) & 0x7fffu);
    if (m <= 2u * kBand) redo |= 1u << j;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    o[i] = strip::pack4(q[4 * i] >> 15, q[4 * i + 1] >> 15, q[4 * i + 2] >> 15, q[4 * i + 3] >> 15);
  return redo;
}

// screen + exact path for the flagged pixels
AWX_HD void fog4(const unsigned (&w)[3], const double (&depth)[4], const Params& fp, unsigned (&o)[3]) {
  unsigned redo = screen4(w, depth, fp, o);
  while (redo) {
    int j = 0;
    while (!((redo >> j) & 1u)) ++j;
    redo &= redo - 1u;
    const double d = j == 0 ? depth[0] : (j == 1 ? depth[1] : (j == 2 ? depth[2] : depth[3]));
    exact_pixel(w, o, 3 * j, d, fp);
  }
}

}  // namespace fog
}  // namespace awx
