// Building blocks of the TMA-fed shared-memory rings (score_v2.cu, fogloss.cu): mbarrier operations, the 1-D
// bulk async copy (cp.async.bulk, SASS UBLKCP) and the elected-thread idiom of the producer warps.
#pragma once
#include <cstdint>

namespace awx {
namespace {

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// true for exactly one (the first active) lane of a converged warp: the pattern ptxas recognises to keep the
// operands of the elected thread's instructions in uniform registers
__device__ __forceinline__ bool elect_one() {
  unsigned pred = 0;
  asm volatile("{\n .reg .pred p;\n elect.sync _|p, 0xffffffff;\n @p mov.u32 %0, 1;\n}" : "+r"(pred));
  return pred != 0;
}
// the same operations on 32-bit shared addresses (the hot loops keep one shared base register and add
// immediates; converting a generic pointer costs an S2R + LEA every time)
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Build knob AWX_WAIT_SLEEP_NS: back off that long after a failed attempt.  try_wait's suspend-time hint is only a hint
// -- a consumer warp of the score kernel re-polls every ~40 cycles (93 SYNCS per pixel) -- so an explicit sleep trades
// wake-up latency against ~150 issued instructions per pixel.  Measured: profiles/r2u_wait_backoff.md.
#ifndef AWX_WAIT_SLEEP_NS
#define AWX_WAIT_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(20000u)
        : "memory");
    if (AWX_WAIT_SLEEP_NS > 0 && !ok) __nanosleep(AWX_WAIT_SLEEP_NS);
  } while (!ok);
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(20000u)  // suspend-time hint (ns): park instead of spinning
        : "memory");
  } while (!ok);
}
// one bounded attempt: true once the phase with the given parity has completed (sleeps up to `ns` otherwise)
__device__ __forceinline__ bool mbar_try(u64* bar, unsigned parity, unsigned ns) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
// non-blocking test
__device__ __forceinline__ bool mbar_test(u64* bar, unsigned parity) {
  unsigned ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 2-D tensor-map copy (cp.async.bulk.tensor, SASS UTMALDG): box (x .. x+w, y .. y+h) of the tensor described by
// `tmap` (a __grid_constant__ CUtensorMap kernel parameter) into dense shared memory; elements outside the tensor
// are zero-filled and the mbarrier is credited with the full box size either way
__device__ __forceinline__ void tma_load_2d(void* dst_smem, const void* tmap, int x, int y, u64* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(smem_u32(bar))
      : "memory");
}

}  // namespace
}  // namespace awx
