// Test-only host build of csrc/blur_strip.cuh: the per-thread code of blur_strip_kernel (csrc/corrupt.cu) under a
// sequential emulation of one CTA per (strip, row segment) -- threads run one after the other, the barrier is the
// end of a loop.  Built by tests/test_blur_strip_cpu.py with g++ -ffp-contract=off into tests/native/_build/.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "blur_strip.cuh"

using namespace awx::strip;

namespace {

template <int R, bool RAIN>
void run_cta(const uint8_t* src, uint8_t* dst, const unsigned* m, int H, int W, int WW, int strip, int ys, int seg,
             const PointParams& pp) {
  using G = Geo<R>;
  constexpr int NR = G::kNR;
  const int NU = W / kUnitPx;
  const int ye = ys + seg < H ? ys + seg : H;
  const int nH = ye - ys + 2 * R;
  const size_t row_bytes = (size_t)W * 3;
  std::vector<Quad> s_h((size_t)2 * NR * G::kRowFloat4);
  std::vector<Raw> raw(G::kThreads);
  std::vector<Window<R>> win(G::kThreads);
  constexpr unsigned kOvBits = ((1u << (kUnitPx + 2 * R)) - 1u) << (8 - R);

  auto load_raw = [&](int hr, int unit, Raw& r) {
    const int sy = reflect101(ys - R + hr, H);
    const uint8_t* g = src + (size_t)sy * row_bytes + (size_t)unit * kUnitE;
    memcpy(r.own, g, 48);
    memset(r.hl, 0, 12);
    memset(r.hr, 0, 12);
    if (unit > 0) memcpy(r.hl, g - 12, 12);
    if (unit + 1 < NU) memcpy(r.hr, g + 48, 12);
    mask_words(m + (size_t)sy * WW, unit, WW, r.m0, r.m1);
  };

  for (int t = 0; t < G::kThreads; ++t) {
    const int hrl = t / G::kStripUnits, unit = strip * G::kStripUnits + t % G::kStripUnits;
    if (hrl < NR && unit < NU && hrl < nH) load_raw(hrl, unit, raw[t]);
  }
  const int iters = (nH + NR - 1) / NR;
  int buf = 0;
  for (int it = 0; it < iters; ++it) {
    for (int t = 0; t < G::kThreads; ++t) {  // H phase
      const int hrl = t / G::kStripUnits, hul = t % G::kStripUnits, unit = strip * G::kStripUnits + hul;
      if (hrl >= NR) continue;
      const int hr = it * NR + hrl;
      const bool hrow = unit < NU && hr < nH;
      Quad* rowbuf = s_h.data() + (size_t)(buf * NR + hrl) * G::kRowFloat4;
      auto store = [&](int q, const F2& a, const F2& c) { rowbuf[G::quad_slot(hul, q)] = Quad{a.x, a.y, c.x, c.y}; };
      if (hrow) {
        finish_raw(raw[t], unit, NU);
        // the device votes per warp; both instruction streams give the same values, so the emulation picks per lane
        // on odd rows and always takes the select stream on even rows (both get exercised)
        if ((raw[t].mb & kOvBits) != 0u || (hr & 1) == 0)
          h_row<R, RAIN, true>(raw[t], pp, store);
        else
          h_row<R, RAIN, false>(raw[t], pp, store);
      }
      if (unit < NU && hr + NR < nH) load_raw(hr + NR, unit, raw[t]);
    }
    for (int g = 0; g < G::kGroups; ++g) {  // V phase (after the barrier)
      const int vu = g / 6, vkg = g - vu * 6, vunit = strip * G::kStripUnits + vu;
      if (vunit >= NU) continue;
      const Quad* vb = s_h.data() + (size_t)(buf * NR) * G::kRowFloat4;
      uint8_t* vdst = dst + (size_t)vunit * kUnitE + vkg * 4;
      auto row = [&](auto jc) {
        constexpr int J = decltype(jc)::value;
        if constexpr (J < NR) {
          const int hj = it * NR + J;
          if (hj < nH) {
            unsigned wlo = 0, whi = 0;
            const bool emit = hj >= 2 * R;
            v_row<R, J>(win[g], vb[J * G::kRowFloat4 + G::group_slot(g, 0)], vb[J * G::kRowFloat4 + G::group_slot(g, 1)], pp.t, emit,
                        wlo, whi);
            if (emit) {
              uint8_t* d = vdst + (size_t)(ys + hj - 2 * R) * row_bytes;
              memcpy(d, &wlo, 4);
              memcpy(d + 24, &whi, 4);
            }
          }
        }
      };
      row(std::integral_constant<int, 0>{});
      row(std::integral_constant<int, 1>{});
      row(std::integral_constant<int, 2>{});
      row(std::integral_constant<int, 3>{});
      row(std::integral_constant<int, 4>{});
      row(std::integral_constant<int, 5>{});
      row(std::integral_constant<int, 6>{});
    }
    buf ^= 1;
  }
}

template <int R, bool RAIN>
void run_image(const uint8_t* src, uint8_t* dst, const unsigned* m, int H, int W, int WW, int seg, const PointParams& pp) {
  const int strips = (W / kUnitPx + Geo<R>::kStripUnits - 1) / Geo<R>::kStripUnits;
  for (int s = 0; s < strips; ++s)
    for (int ys = 0; ys < H; ys += seg) run_cta<R, RAIN>(src, dst, m, H, W, WW, s, ys, seg, pp);
}

}  // namespace

// img / out: uint8 [H, W, 3]; mask: 1 bit per pixel, WW = (W + 31) / 32 words per row; taps: half kernel, taps[0] = centre
extern "C" int blur_strip_emulate(const uint8_t* img, uint8_t* out, const unsigned* mask, int H, int W, int radius, int rain,
                                  float k1, float k2, const float* taps, int seg) {
  if (W % kUnitPx != 0 || H < 1 || seg < 1 || (radius != 1 && radius != 3)) return 1;
  const int WW = (W + 31) / 32;
  PointParams pp;
  pp.k1 = k1;
  pp.k2 = k2;
  for (int i = 0; i < 4; ++i) pp.t[i] = i <= radius ? taps[i] : 0.0f;
  pp.negzero = -0.0f;
  if (radius == 1 && rain) run_image<1, true>(img, out, mask, H, W, WW, seg, pp);
  else if (radius == 1) run_image<1, false>(img, out, mask, H, W, WW, seg, pp);
  else if (rain) run_image<3, true>(img, out, mask, H, W, WW, seg, pp);
  else run_image<3, false>(img, out, mask, H, W, WW, seg, pp);
  return 0;
}
