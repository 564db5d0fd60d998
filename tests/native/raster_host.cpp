// Test-only host build of csrc/raster.cuh (the product compiles the same header for the device).
// Built by tests/test_raster_cpu.py with g++ into tests/native/_build/.
#include <stddef.h>
#include <stdint.h>
#include "raster.cuh"
#include "convert.cuh"

extern "C" void raster_lines(const int32_t* items, int n, int H, int W, uint8_t* mask) {
  auto emit = [&](int y, int x0, int x1) {
    for (int x = x0; x <= x1; ++x) mask[(size_t)y * W + x] = 1;
  };
  for (int i = 0; i < n; ++i) {
    const int32_t* it = items + 5 * i;
    awx::raster::line(it[0], it[1], it[2], it[3], it[4], W, H, emit);
  }
}

extern "C" void raster_discs(const int32_t* items, int n, int H, int W, uint8_t* mask) {
  auto emit = [&](int y, int x0, int x1) {
    for (int x = x0; x <= x1; ++x) mask[(size_t)y * W + x] = 1;
  };
  for (int i = 0; i < n; ++i) {
    const int32_t* it = items + 5 * i;
    awx::raster::disc(it[0], it[1], it[2], W, H, emit);
  }
}

extern "C" void unit_table(float* out256) {
  for (unsigned u = 0; u < 256; ++u) out256[u] = awx::unit_of_u8(u);
}
