// Internal helpers shared by the libawx translation units (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>

#include "awx.h"

namespace awx {

// ----------------------------------------------------------------------------- errors
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* where);

#define AWX_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) {                   \
      ::awx::set_error(__VA_ARGS__); \
      return (code);                 \
    }                                \
  } while (0)

#define AWX_CUDA(call)                                            \
  do {                                                            \
    cudaError_t e__ = (call);                                     \
    if (e__ != cudaSuccess) return ::awx::cuda_fail(e__, #call);  \
  } while (0)

int sm_count();
int launch_gauss_f64(const double* in, void* out, int32_t out_dtype, double* tmp, int64_t batch, int32_t H, int32_t W,
                     double ramp_scale, double floor_value, const double* weights /*HOST*/, int32_t radius, cudaStream_t s);
void note_launch(int n = 1);  // counts kernels launched by this library (awx_launch_count)

// --------------------------------------------------------------------- device helpers
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float lg2_approx(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Streaming (read-once) global loads: read-only path, do not allocate in L1.
__device__ __forceinline__ float ld_stream(const float* p) {
  float v;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float2 ld_stream2(const float* p) {
  float2 v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ uint4 ld_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void st_stream_u4(void* p, uint4 v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w));
}

__device__ __forceinline__ bool is_nan(float x) { return x != x; }

// torch.argmax / torch.max ordering: NaN beats every number, first index wins ties.
__device__ __forceinline__ bool beats(float cand, float best) {
  return (cand > best) || (is_nan(cand) && !is_nan(best));
}

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

}  // namespace awx
