// Throughput micro-benchmark of the pipes awx_score leans on (dev tool): FFMA, FFMA2 (f32x2), MUFU.EX2, FMNMX,
// FSETP+SEL, LDS.  Each kernel runs ILP independent chains per thread; prints warp-instructions / clk / SM.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define ILP 8
#define ITERS 4096

__global__ void k_ffma(float* out, float a, float b) {
  float x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = fmaf(x[i], a, b);
  float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_ffma2(float* out, float a, float b) {
  u64 x[ILP];
  float2 av = make_float2(a, a), bv = make_float2(b, b);
  u64 A = *reinterpret_cast<u64*>(&av), B = *reinterpret_cast<u64*>(&bv);
  for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x + i, i); x[i] = *reinterpret_cast<u64*>(&t); }
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
  float s = 0; for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mufu(float* out, float a) {
  float x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = -(float)(threadIdx.x + i) * a;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fmnmx(float* out, float a) {
  float x[ILP];
  for (int i = 0; i < ILP; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("max.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(a + it));
  float s = 0; for (int i = 0; i < ILP; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_mix(float* out, float a, float b) {  // 1 MUFU : 2 FFMA2 : 2 ALU per chain step
  u64 x[ILP]; float y[ILP], z[ILP];
  float2 av = make_float2(a, a), bv = make_float2(b, b);
  u64 A = *reinterpret_cast<u64*>(&av), B = *reinterpret_cast<u64*>(&bv);
  for (int i = 0; i < ILP; ++i) { float2 t = make_float2(threadIdx.x + i, i); x[i] = *reinterpret_cast<u64*>(&t); y[i] = -i; z[i] = i; }
  for (int it = 0; it < ITERS; ++it)
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(A), "l"(B));
      asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(y[i]));
      asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(B), "l"(A));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(a + it));
      asm volatile("max.f32 %0, %0, %1;" : "+f"(z[i]) : "f"(b + it));
    }
  float s = 0; for (int i = 0; i < ILP; ++i) { float2 t = *reinterpret_cast<float2*>(&x[i]); s += t.x + t.y + y[i] + z[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
void run(const char* name, F launch, double instr_per_thread, int threads, int sms, double clock_ghz) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  launch(); cudaDeviceSynchronize();
  cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warp_instr = instr_per_thread * threads / 32.0 * sms;
  const double clks = ms * 1e-3 * clock_ghz * 1e9;
  printf("%-28s %8.3f ms  %6.3f warp-instr/clk/SM  (%d threads/SM)\n", name, ms, warp_instr / clks / sms, threads);
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  double ghz = khz * 1e-6; printf("%s: %d SMs, %.3f GHz (max)\n", p.name, sms, ghz);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024);
  for (int threads : {128, 512, 1024}) {
    run("FFMA", [&] { k_ffma<<<sms, threads>>>(out, 1.0001f, 0.5f); }, (double)ILP * ITERS, threads, sms, ghz);
    run("FFMA2 (f32x2)", [&] { k_ffma2<<<sms, threads>>>(out, 1.0001f, 0.5f); }, (double)ILP * ITERS, threads, sms, ghz);
    run("MUFU.EX2", [&] { k_mufu<<<sms, threads>>>(out, 0.01f); }, (double)ILP * ITERS, threads, sms, ghz);
    run("FMNMX", [&] { k_fmnmx<<<sms, threads>>>(out, 0.5f); }, (double)ILP * ITERS, threads, sms, ghz);
    run("mix 2 FFMA2 + 1 MUFU + 2 FMNMX", [&] { k_mix<<<sms, threads>>>(out, 1.0001f, 0.5f); }, 5.0 * ILP * ITERS, threads, sms, ghz);
  }
  return 0;
}
