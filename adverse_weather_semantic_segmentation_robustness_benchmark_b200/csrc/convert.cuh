// Host/device scalar conversions shared by the corruption kernels (and compiled for the host by
// tests/test_raster_cpu.py to be checked exhaustively).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define AWX_HD __host__ __device__ __forceinline__
#else
#define AWX_HD inline
#endif

namespace awx {

// fl(u / 255) for a byte value without a table or a division: 1/255 split into fl(1/255) + a low word, the
// low product rounded on its own and the high one fused -- fl(u * hi + fl(u * lo)) is the correctly rounded
// quotient for all 256 values (checked exhaustively by tests/test_raster_cpu.py): I2F + FMUL + FFMA.
// A shared 256-entry table costs ~3.5-way bank conflicts on random bytes plus 256 divisions per CTA.
AWX_HD float unit_of_u8(unsigned u) {
  const float hi = 0.003921568859368563f;     // fl(1/255)
  const float lo = -2.319175823606301e-10f;   // fl(1/255 - hi)
  const float uf = (float)u;
#ifdef __CUDA_ARCH__
  return fmaf(uf, hi, __fmul_rn(uf, lo));
#else
  volatile float t = uf * lo;  // keep the product's own rounding whatever the host's contraction setting
  return fmaf(uf, hi, t);
#endif
}

}  // namespace awx
