// Row-walking separable Gaussian of the rain / snow corruptions (blur_strip_kernel in corrupt.cu).
//
// A CTA owns a STRIP of 32 (3 taps) or 27 (7 taps) units -- a unit = 16 pixels = 48 bytes = 48 ELEMENTS of the HWC
// row -- and walks down a segment of rows, NR rows per iteration, every input row being read, point-op'ed and
// row-filtered exactly once:
//
//   H phase  thread tau takes (row it*NR + tau / units, unit tau % units): for 3 taps warp = row, lane = unit.  The lane loads its 48 bytes plus the 16 bytes before and
//            after them (three pixels of halo each side; BORDER_REFLECT_101 at the image border is a byte shuffle of
//            the lane's own unit), converts, applies the point operation and the overlay, and runs the horizontal
//            filter ENTIRELY IN REGISTERS on 24 packed pairs {element k, element k + 24} -- both halves of a pair have
//            the same channel and all tap offsets (multiples of 3 elements) shift k by whole pairs, so every operand
//            is an aligned register pair: FMUL2 / FADD2 / FFMA2, two elements per issue slot.  The 24 filtered
//            pairs go to shared memory as 12 conflict-free 16-byte stores.
//   V phase  thread g (< 6 x units) owns the 4 pairs {48u + 4kg + j, 48u + 24 + 4kg + j}, j < 4, of unit u = g / 6 for the
//            whole walk: per row two 16-byte shared loads refill one slot of a (2R+1)-deep REGISTER window (the ring
//            index is static: NR is a multiple of 2R+1), the vertical filter runs packed, and two 32-bit stores write
//            the bytes (six lanes cover 24 contiguous bytes; the two stores of a warp interleave).
//
// Shared memory carries 4 bytes out and 4 bytes in per element and nothing else; the tile kernel it replaces moved 27.
// The arithmetic is OpenCV's (cv2.GaussianBlur on CV_32FC3, vector body, FMA build), order and fusion included:
//   row filter, 3 taps      fma(c, t0, fl((l + r) * t1))                     (SymmRowSmallVec_32f)
//   row filter, 7 taps      left-to-right chain fl(x0 * t3), fma(x1, t2, .), ...    (RowVec_32f)
//   column filter           fl(c * t0), fma(u1 + d1, t1, .), fma(u2 + d2, t2, .), ...   (SymmColumnVec_32f)
// checked against cv2 itself on the host by tests/test_blur_strip_cpu.py, which compiles THIS header with g++ and
// runs the same per-thread code under a sequential CTA emulation (the device kernel and the emulation share every
// function below; only the loop over threads and the barrier differ).
#pragma once

#include <math.h>
#include <stdint.h>

#include "convert.cuh"

namespace awx {
namespace strip {

constexpr int kUnitPx = 16;
constexpr int kUnitE = 48;
#ifndef AWX_STRIP_CTAS_R1
#define AWX_STRIP_CTAS_R1 3   // build knob (CTAs per SM of the 3-tap kernels): 3 = 96 registers, no spills, 18 warps
#endif
#ifndef AWX_STRIP_CTAS_R3
#define AWX_STRIP_CTAS_R3 2   // 7-tap kernels
#endif
#ifndef AWX_STRIP_UNITS_R3
#define AWX_STRIP_UNITS_R3 27  // build knob (units per strip of the 7-tap kernels): 32 = 7 warps, 128 registers with
                               // spills, 0.594 ms per 64 frames against 0.565
#endif
template <int R>
struct Geo {
  static constexpr int kD = 2 * R + 1;                // window depth
  static constexpr int kNR = R == 1 ? 6 : 7;          // rows per iteration; a multiple of kD
  // Units per strip.  H-phase task tau = threadIdx.x: row tau / kStripUnits of the iteration, unit tau % kStripUnits
  // of the strip.  3 taps: 32 units, 6 x 32 = 192 tasks = 6 warps, warp = row.  7 taps: the register window alone is
  // 56 registers, and a thread may only have 128 where a scheduler hosts 4 warps (2 CTAs of 7 warps: 14 warps on 4
  // schedulers) -- which spilled ~40 words per thread and iteration.  27 units make 7 x 27 = 189 tasks fit 6 warps:
  // 2 CTAs = 12 warps = 3 per scheduler = up to 168 registers, nothing spilled.
  static constexpr int kStripUnits = R == 1 ? 32 : AWX_STRIP_UNITS_R3;
  static constexpr int kGroups = kStripUnits * 6;     // V-phase threads (8 elements each)
  static constexpr int kThreads = ((kNR * kStripUnits + 31) / 32) * 32;
  // float4 slots per half row: one pad slot per 4 units, so that the H-phase stores (unit stride 6 slots) hit 8
  // distinct bank groups; rounded up to a multiple of 8
  static constexpr int kRowSlots = ((kGroups + (kStripUnits + 3) / 4 + 7) / 8) * 8;
  static constexpr int kRowFloat4 = 2 * kRowSlots;    // one filtered row in shared memory: [half][slot]
  // R = 3: the filtered rows are double buffered, one barrier per iteration, 2 CTAs per SM.  R = 1 gets by with far
  // fewer registers: a single buffer and a second barrier per iteration let several CTAs share an SM, which hides both
  // barriers and the dependent-issue latencies.  Measured per 64 frames, rain / snow-3: 2 CTAs 0.368 / 0.346 ms,
  // 3 CTAs (96 registers, no spills) 0.322 / 0.299, 4 CTAs (80 registers, a few spilled loop scalars) 0.357 / 0.311.
  static constexpr int kBuffers = R == 1 ? 1 : 2;
  static constexpr int kCtasPerSm = R == 1 ? AWX_STRIP_CTAS_R1 : AWX_STRIP_CTAS_R3;
  static constexpr int kSmemBytes = kBuffers * kNR * kRowFloat4 * 16;
  static_assert(kNR % kD == 0, "static ring index");
  static_assert(kThreads >= kGroups, "every V group has a thread");
  // shared-memory slot (in float4 units, within one filtered row) of quad q = 2 * kg + h of unit u of the strip ...
  AWX_HD static int quad_slot(int u, int q) { return (q & 1) * kRowSlots + 6 * u + (q >> 1) + (u >> 2); }
  // ... and of V group g = 6 * u + kg, half h
  AWX_HD static int group_slot(int g, int h) { return h * kRowSlots + g + g / 24; }
};

// ------------------------------------------------------------------------------------------------ packed pairs
#if defined(__CUDA_ARCH__)
using F2 = float2;
__device__ __forceinline__ unsigned long long& f2bits(F2& v) { return *reinterpret_cast<unsigned long long*>(&v); }
__device__ __forceinline__ const unsigned long long& f2bits(const F2& v) {
  return *reinterpret_cast<const unsigned long long*>(&v);
}
__device__ __forceinline__ F2 mul2(const F2 a, const F2 b) {
  F2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(f2bits(d)) : "l"(f2bits(a)), "l"(f2bits(b)));
  return d;
}
__device__ __forceinline__ F2 add2(const F2 a, const F2 b) {
  F2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(f2bits(d)) : "l"(f2bits(a)), "l"(f2bits(b)));
  return d;
}
__device__ __forceinline__ F2 fma2(const F2 a, const F2 b, const F2 c) {
  F2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(f2bits(d)) : "l"(f2bits(a)), "l"(f2bits(b)), "l"(f2bits(c)));
  return d;
}
__device__ __forceinline__ unsigned to_byte(float v255) {  // truncate toward zero, saturate to [0, 255]
  unsigned r;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(v255));
  return r;
}
#else
struct F2 {
  float x, y;
};
inline F2 mul2(const F2 a, const F2 b) {
  volatile float x = a.x * b.x, y = a.y * b.y;
  return F2{x, y};
}
inline F2 add2(const F2 a, const F2 b) {
  volatile float x = a.x + b.x, y = a.y + b.y;
  return F2{x, y};
}
inline F2 fma2(const F2 a, const F2 b, const F2 c) { return F2{fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)}; }
inline unsigned to_byte(float v255) {
  if (!(v255 > 0.0f)) return 0u;
  if (v255 >= 255.0f) return 255u;
  return (unsigned)(int)v255;
}
#endif
// the low bytes of four words -> one word: three byte permutes on the device
AWX_HD unsigned pack4(unsigned b0, unsigned b1, unsigned b2, unsigned b3) {
#if defined(__CUDA_ARCH__)
  return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
#else
  return (b0 & 0xffu) | ((b1 & 0xffu) << 8) | ((b2 & 0xffu) << 16) | (b3 << 24);
#endif
}
// clip(fl(x + k), 0, 1): one saturating add on the device
AWX_HD float add_clip01(float x, float k) {
#if defined(__CUDA_ARCH__)
  return __saturatef(__fadd_rn(x, k));
#else
  volatile float v = x + k;
  return fminf(fmaxf(v, 0.0f), 1.0f);
#endif
}
AWX_HD F2 pair(float x, float y) {
  F2 p;
  p.x = x;
  p.y = y;
  return p;
}
AWX_HD F2 splat2(float x) { return pair(x, x); }

AWX_HD int reflect101(int i, int n) {  // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

// ------------------------------------------------------------------------------------------------ one unit, raw
struct Raw {
  unsigned own[12];  // elements 0 .. 47
  unsigned hl[3];    // elements -12 .. -1  (the four pixels left of the unit)
  unsigned hr[3];    // elements 48 .. 59   (the four pixels right of it)
  unsigned m0, m1;   // the two mask words around the unit as loaded (see mask_words); finish_raw turns them into mb
  unsigned mb;       // overlay bits: bit 8 + j  <->  pixel j of the unit, j in [-8, 24)
};

AWX_HD unsigned pick4(const unsigned* w, int b0, int b1, int b2, int b3) {
  auto byte = [&](int b) { return (w[b >> 2] >> ((b & 3) * 8)) & 0xffu; };
  return byte(b0) | (byte(b1) << 8) | (byte(b2) << 16) | (byte(b3) << 24);
}

// BORDER_REFLECT_101 for the first / last unit of an image row: the halo pixels are the unit's own pixels mirrored.
AWX_HD void reflect_left(Raw& r) {
  // elements -12 .. -1 = pixels -4, -3, -2, -1 = pixels 4, 3, 2, 1
  const unsigned a = pick4(r.own, 12, 13, 14, 9), b = pick4(r.own, 10, 11, 6, 7), c = pick4(r.own, 8, 3, 4, 5);
  r.hl[0] = a;
  r.hl[1] = b;
  r.hl[2] = c;
  unsigned m = r.mb & ~0xffu;  // bit 8 - q := bit 8 + q, q = 1 .. 8
#pragma unroll
  for (int q = 1; q <= 8; ++q) m |= ((r.mb >> (8 + q)) & 1u) << (8 - q);
  r.mb = m;
}
AWX_HD void reflect_right(Raw& r) {
  // elements 48 .. 59 = pixels 16, 17, 18, 19 = pixels 14, 13, 12, 11  (bytes 42.., 39.., 36.., 33..)
  const unsigned a = pick4(r.own, 42, 43, 44, 39), b = pick4(r.own, 40, 41, 36, 37), c = pick4(r.own, 38, 33, 34, 35);
  r.hr[0] = a;
  r.hr[1] = b;
  r.hr[2] = c;
  unsigned m = r.mb & 0x00ffffffu;  // bit 24 + i := bit 22 - i, i = 0 .. 7
#pragma unroll
  for (int i = 0; i < 8; ++i) m |= ((r.mb >> (22 - i)) & 1u) << (24 + i);
  r.mb = m;
}

// The overlay bits of pixels 16u - 8 .. 16u + 23 straddle two words of the 1-bit-per-pixel mask row (`mrow`: WW 32-bit
// words).  The LOADS are issued with the rest of the unit (one iteration ahead); nothing may consume them there, or the
// prefetch stalls on its own loads -- finish_raw runs at the start of the H phase.
AWX_HD void mask_words(const unsigned* mrow, int unit, int WW, unsigned& m0, unsigned& m1) {
  const int wi = unit >> 1;
  if ((unit & 1) == 0) {
    m0 = wi > 0 ? mrow[wi - 1] : 0u;
    m1 = mrow[wi];
  } else {
    m0 = mrow[wi];
    m1 = wi + 1 < WW ? mrow[wi + 1] : 0u;
  }
}
AWX_HD void finish_raw(Raw& r, int unit, int n_units) {
  r.mb = (unit & 1) == 0 ? ((r.m1 << 8) | (r.m0 >> 24)) : ((r.m0 >> 8) | (r.m1 << 24));
  if (unit == 0) reflect_left(r);
  if (unit == n_units - 1) reflect_right(r);
}

// ------------------------------------------------------------------------------------------------ H phase
// Point operation + overlay + horizontal filter of one unit; the 24 filtered pairs are handed to `store(q, v)` as 12
// float4-sized quads q = 2 * kg + h: pairs 4kg + 2h and 4kg + 2h + 1.
struct PointParams {
  float k1, k2;      // rain: fl(fl(x * k1) + k2);  snow: clip(fl(x + k1), 0, 1)
  float t[4];        // half Gaussian kernel, t[0] = centre
  float negzero;     // a RUNTIME -0.0f: fma(x, k1, -0) == fl(x * k1), and ptxas cannot contract the following add
};

template <int R, bool RAIN, bool OV, class Store>
AWX_HD void h_row(const Raw& raw, const PointParams& pp, Store&& store) {
  constexpr int HE = 3 * R;           // halo elements per side
  constexpr int NP = 24 + 2 * HE;     // pairs to point-op: k = -HE .. 24 + HE - 1
  const F2 lo2 = splat2(-2.319175823606301e-10f), hi2 = splat2(0.003921568859368563f);  // convert.cuh: fl(u / 255)
  const F2 k1 = splat2(pp.k1), k2 = splat2(pp.k2), nz = splat2(pp.negzero);
  const F2 t0 = splat2(pp.t[0]), t1 = splat2(pp.t[1]), t2 = splat2(pp.t[2]), t3 = splat2(pp.t[3]);

  auto elem_byte = [&](int e) -> unsigned {  // e in [-12, 60): static after unrolling
    if (e < 0) {
      const int b = e + 12;
      return (raw.hl[b >> 2] >> ((b & 3) * 8)) & 0xffu;
    }
    if (e >= kUnitE) {
      const int b = e - kUnitE;
      return (raw.hr[b >> 2] >> ((b & 3) * 8)) & 0xffu;
    }
    return (raw.own[e >> 2] >> ((e & 3) * 8)) & 0xffu;
  };
  auto point_pair = [&](int k) -> F2 {  // pair k = {element k, element k + 24}
    const F2 uf = pair((float)elem_byte(k), (float)elem_byte(k + 24));
    const F2 x = fma2(uf, hi2, mul2(uf, lo2));
    F2 v;
    if (RAIN) {
      v = add2(fma2(x, k1, nz), k2);
    } else {
      v = pair(add_clip01(x.x, pp.k1), add_clip01(x.y, pp.k1));
    }
    if (OV) {
      const int c = ((k % 3) + 3) % 3;
      const int px = (k + 48) / 3 - 16;  // floor(k / 3)
      const float colour = RAIN ? (c == 0 ? 0.8f : (c == 1 ? 0.9f : 1.0f)) : 1.0f;
      if ((raw.mb >> (8 + px)) & 1u) v.x = colour;
      if ((raw.mb >> (16 + px)) & 1u) v.y = colour;
    }
    return v;
  };

  F2 P[NP];   // only a window of 2 * HE + 1 pairs is live at any time (straight-line code after unrolling)
  F2 Hq[2];
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    P[i] = point_pair(i - HE);
    const int k = i - 2 * HE;  // the output pair that just became computable: taps k - HE .. k + HE
    if (k >= 0) {
      const F2* c = P + (k + HE);
      F2 v;
      if (R == 1) {
        v = fma2(c[0], t0, mul2(add2(c[-3], c[3]), t1));
      } else {
        v = mul2(c[-9], t3);
        v = fma2(c[-6], t2, v);
        v = fma2(c[-3], t1, v);
        v = fma2(c[0], t0, v);
        v = fma2(c[3], t1, v);
        v = fma2(c[6], t2, v);
        v = fma2(c[9], t3, v);
      }
      Hq[k & 1] = v;
      if (k & 1) store(k >> 1, Hq[0], Hq[1]);
    }
  }
}

// ------------------------------------------------------------------------------------------------ V phase
template <int R>
struct Window {
  F2 w[2 * R + 1][4];
};

struct Quad {  // a float4: two pairs
  float a, b, c, d;
};

// Row J of an iteration (J static): refill slot J mod D with the filtered row just produced; if `emit`, filter the
// window centred R rows back and return the two output words (first-half bytes, second-half bytes).
template <int R, int J>
AWX_HD void v_row(Window<R>& win, const Quad& q0, const Quad& q1, const float (&t)[4], bool emit, unsigned& word_lo,
                  unsigned& word_hi) {
  constexpr int D = 2 * R + 1;
  constexpr int S = J % D;
  win.w[S][0] = pair(q0.a, q0.b);
  win.w[S][1] = pair(q0.c, q0.d);
  win.w[S][2] = pair(q1.a, q1.b);
  win.w[S][3] = pair(q1.c, q1.d);
  if (!emit) return;
  const F2 t0 = splat2(t[0]), t1 = splat2(t[1]), t2 = splat2(t[2]), t3 = splat2(t[3]), s255 = splat2(255.0f);
  // newest row = J (slot S); centre = J - R; rows J - 2R .. J are in the window
  constexpr int C = ((J - R) % D + D) % D;
  unsigned lo[4], hi[4];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    F2 v = mul2(win.w[C][p], t0);
    v = fma2(add2(win.w[(C + D - 1) % D][p], win.w[(C + 1) % D][p]), t1, v);
    if (R == 3) {
      v = fma2(add2(win.w[(C + D - 2) % D][p], win.w[(C + 2) % D][p]), t2, v);
      v = fma2(add2(win.w[(C + D - 3) % D][p], win.w[(C + 3) % D][p]), t3, v);
    }
    v = mul2(v, s255);
    lo[p] = to_byte(v.x);
    hi[p] = to_byte(v.y);
  }
  word_lo = pack4(lo[0], lo[1], lo[2], lo[3]);
  word_hi = pack4(hi[0], hi[1], hi[2], hi[3]);
}

}  // namespace strip
}  // namespace awx
