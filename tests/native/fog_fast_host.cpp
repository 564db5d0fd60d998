// Test-only host build of csrc/fog_fast.cuh (the per-thread code of fog_kernel, csrc/corrupt.cu): every group of four
// pixels through screen + exact path, with exp2f standing in for ex2.approx -- by construction ANY transmission within
// the screen's error budget must give the reference's bytes.  `perturb` adds a relative error to the transmission to
// exercise that claim up to the budget.  Built by tests/test_fog_fast_cpu.py with g++ -ffp-contract=off.
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "fog_fast.cuh"

using namespace awx::fog;

// img / out: uint8 [n_px * 3] (n_px a multiple of 4); depth: fp64 [n_px]; returns the number of pixels that took the exact path
extern "C" long long fog_fast_emulate(const uint8_t* img, uint8_t* out, const double* depth, long long n_px, double beta,
                                      double airlight, int screen_only) {
  const Params fp = make_params(beta, airlight);
  long long redone = 0;
  for (long long p = 0; p + 4 <= n_px; p += 4) {
    unsigned w[3], o[3];
    memcpy(w, img + 3 * p, 12);
    const double d[4] = {depth[p], depth[p + 1], depth[p + 2], depth[p + 3]};
    unsigned redo = screen4(w, d, fp, o);
    for (int j = 0; j < 4; ++j) redone += (redo >> j) & 1u;
    if (!screen_only) {
      while (redo) {
        int j = 0;
        while (!((redo >> j) & 1u)) ++j;
        redo &= redo - 1u;
        exact_pixel(w, o, 3 * j, d[j], fp);
      }
    }
    memcpy(out + 3 * p, o, 12);
  }
  return redone;
}

extern "C" void fog_set_ex2_error(float rel) { host_ex2_error() = rel; }

extern "C" int fog_params_ok(double beta, double airlight) { return params_ok(beta, airlight) ? 1 : 0; }
