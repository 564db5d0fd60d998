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

// fl(u / 255) for a byte value without a table or a division: q = u * fl(1/255) corrected by one exact FMA
// residual step is the correctly rounded quotient (Markstein; checked against u / 255 for all 256 values).
// A shared 256-entry table costs ~3.5-way bank conflicts on random bytes plus 256 divisions per CTA.
AWX_HD float unit_of_u8(unsigned u) {
  const float r = 0.003921568859368563f;  // fl(1/255)
  const float uf = (float)u;
#ifdef __CUDA_ARCH__
  const float q = __fmul_rn(uf, r);
#else
  const float q = uf * r;
#endif
  const float e = fmaf(-q, 255.0f, uf);
  return fmaf(e, r, q);
}

}  // namespace awx
