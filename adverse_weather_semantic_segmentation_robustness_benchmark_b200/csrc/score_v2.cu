// awx_score, kernel v2 (C == 19): TMA-staged, register-resident, packed-fp32 scoring.
//
// One persistent CTA per SM: warp 15 is the producer, warps 0-14 (480 threads) consume.
//   producer  walks the CTA's tiles (480 consecutive pixels of one image) and, per member, issues 19
//             bulk async copies (cp.async.bulk, one 1.9 KB plane segment each) into the next free unit of
//             a shared-memory ring, completing on that unit's mbarrier.  The ring holds 4-5 units of
//             36 KB, i.e. up to ~180 KB per SM in flight independent of what the consumers do.
//   consumer  thread t owns pixel t of the tile: after the unit's barrier flips it pulls its 19 (or 38)
//             values into registers with immediate-offset LDS and releases the unit at once, so
//             shared memory is only a landing zone.  Element-wise arithmetic runs on float2 PAIRS OF
//             CLASSES (FADD2 / FMUL2 / FFMA2), transcendentals on the MUFU pipe.
// Exactness: ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 even with explicit rounding
// modifiers, so the three-rounding fusion w0*a + w1*b is written as
// add2(fma2(w0,a,-0), fma2(w1,b,-0)) with a runtime -0.0 addend: fma(x,y,-0) == fl(x*y) bit for
// bit, and the final add is a true FADD2.
// Statistics: CTA-private shared histograms with a warp-uniform fast path (__match_all_sync) for
// piecewise-constant real data; ECE bins are private to the warp (32-bit packed words, flushed to
// 64-bit before they can overflow).  Pixels with non-finite sums (NaN / inf logits) and pixels
// within 16 ulp of an ECE edge take the scalar slow paths shared with kernel v1.
#include "score_common.cuh"

namespace awx {
using namespace score_detail;
namespace {

constexpr int kC = 19;
constexpr int kTP = 480;                        // pixels per tile (15 consumer warps; 16 warps total -> 128 regs)
constexpr int kCons = 480;                      // consumer threads (1 px each)
constexpr int kConsWarps = kCons / 32;
constexpr int kV2Threads = kCons + 32;          // + producer warp
constexpr int kUnitFloats = kC * kTP;
constexpr int kUnitBytes = kUnitFloats * 4;     // 36480
constexpr int kMaxUnits = 5;
constexpr unsigned kFlushPixels = 60000;        // per-warp ECE words are flushed before 2^16 pixels

typedef unsigned long long u64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(u64* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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
__device__ __forceinline__ void bulk_load(void* dst_smem, const void* src, unsigned bytes, u64* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2)
__device__ __forceinline__ u64& bits(float2& v) { return *reinterpret_cast<u64*>(&v); }
__device__ __forceinline__ const u64& bits(const float2& v) { return *reinterpret_cast<const u64*>(&v); }
__device__ __forceinline__ float2 mul2(const float2 a, const float2 b) {
  float2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(bits(d)) : "l"(bits(a)), "l"(bits(b)));
  return d;
}
__device__ __forceinline__ float2 add2(const float2 a, const float2 b) {
  float2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(bits(d)) : "l"(bits(a)), "l"(bits(b)));
  return d;
}
__device__ __forceinline__ float2 sub2(const float2 a, const float2 b) {
  float2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(bits(d)) : "l"(bits(a)), "l"(bits(b)));
  return d;
}
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(bits(d)) : "l"(bits(a)), "l"(bits(b)), "l"(bits(c)));
  return d;
}
__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }
__device__ __forceinline__ float2 max2(const float2 a, const float2 b) { return make_float2(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ float2 ex2_2(const float2 a) { return make_float2(ex2_approx(a.x), ex2_approx(a.y)); }
__device__ __forceinline__ float2 lg2_2(const float2 a) { return make_float2(lg2_approx(a.x), lg2_approx(a.y)); }
__device__ __forceinline__ bool finite2(const float2 a) { return isfinite(a.x) && isfinite(a.y); }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// largest float below x (finite x)
__device__ __forceinline__ float float_prev(float x) {
  if (x == 0.f) return -1.401298464e-45f;
  const int b = __float_as_int(x);
  return __int_as_float(x > 0.f ? b - 1 : b + 1);
}
// x / T without branches: q0 = RN(x*y), r = RN(x - q0*T) (exact, FMA), q = RN(q0 + r*y) with
// y = RN(1/T) from the host is the correctly rounded quotient (Markstein) for normal-range operands;
// tiny / huge |x| (where the residual could underflow or q overflow) take __fdiv_rn.
__device__ __forceinline__ float div_by_T(float x, float T, float rT) {
  const float q0 = x * rT;
  const float r = fmaf(-q0, T, x);
  return fmaf(r, rT, q0);
}

// (lo, hi] bin of conf; edges are within an ulp of i/nb so the guess is off by at most one
__device__ __forceinline__ int ece_bin_fast(float conf, const float* e, int nb) {
  int b = min(max((int)ceilf(conf * (float)nb) - 1, 0), nb - 1);
  b -= (b > 0 && !(conf > e[b])) ? 1 : 0;
  b += (b < nb - 1 && conf > e[b + 1]) ? 1 : 0;
  return (conf > e[b] && conf <= e[b + 1]) ? b : -1;
}

// byte offset of the TMA ring inside dynamic shared memory (everything before it is bookkeeping)
__host__ __device__ inline size_t v2_ring_offset(int nb, int NB) {
  size_t o = 2 * kMaxUnits * sizeof(u64) + (8 + 368) * 4 + (AWX_MAX_ECE_BINS + 4) * 4;
  o += (size_t)kConsWarps * nb * 36 + (size_t)2 * NB * 4;
  return (o + 127) & ~(size_t)127;
}

// histogram add with a warp-uniform fast path; key < 0 = nothing to add
__device__ __forceinline__ void hist_add(unsigned* h, int key, int lane) {
  int same;
  __match_all_sync(0xffffffffu, key, &same);
  if (same) {
    if (lane == 0 && key >= 0) atomicAdd(h + key, 32u);
  } else if (key >= 0) {
    atomicAdd(h + key, 1u);
  }
}

// Slow path for one pixel straight from global memory (NaN / inf logits): the scalar v1 code.
template <bool ENS, bool JS>
__device__ __noinline__ void slow_pixel(const ScoreParams& p, const float* s_edges, const float* ga, const float* gb,
                                        PixOut& o) {
  float a[kC], b[ENS ? kC : 1];
  float amax = 0.f, bmax = 0.f;
  for (int c = 0; c < kC; ++c) {
    a[c] = ga[c * p.HW];
    if (ENS) b[ENS ? c : 0] = gb[c * p.HW];
  }
  if (ENS) {
    amax = a[0];
    bmax = b[0];
    for (int c = 1; c < kC; ++c) {
      amax = fmaxf(amax, a[c]);
      bmax = fmaxf(bmax, b[ENS ? c : 0]);
    }
  }
  o.mi = 0.f;
  o.js = 0.f;
  o.mpred = 0;
  score_pixel<kC, ENS, JS>(a, b, kC, p, s_edges, ga, gb, p.w0, p.w1, amax, bmax, o);
}

// MODE: 0 single member, 1 weighted average, 2 mean.
// FAST: 0 generic (runtime label dtype, optional per-pixel maps), 1 uint8 labels and bins only,
//       2 int64 labels and bins only -- the streaming-evaluation configurations, with every
//       map / dtype branch compiled out.
//
// Consumer thread t owns pixel t of the 480-pixel tile.  Its 19 class values are held as 10 float2
// PAIRS OF CLASSES (2i, 2i+1); the 20th slot is a finite "never wins" dummy (-1e30) whose
// exponentials are exactly 0.  Element-wise work (fusion, x - max, scaling, e*d products) runs on
// FADD2/FMUL2/FFMA2 over class pairs; maxima and arg-maxima are scalar (no packed min/max exists).
//
// Shared memory: [mbarriers | counters | confusion | edges | per-warp ECE words | AUROC | ring].
// DIV: -1 runtime p.div_mode (generic kernels), 0 / 1 compile-time (bins-only kernels).
template <int MODE, bool JS, int FAST, int DIV>
__global__ void __launch_bounds__(kV2Threads, 1) score_v2_kernel(const __grid_constant__ ScoreParams p, const int NU,
                                                                  const float negzero) {
  constexpr bool ENS = MODE != 0;
  constexpr int NP = (kC + 1) / 2;  // 10 class pairs
  constexpr float kDummy = -1e30f;
  const int nb = p.nb, NB = p.auroc_bins;
  const bool have_labels = FAST != 0 || p.labels != nullptr;
  extern __shared__ __align__(128) unsigned char smem[];
  u64* full = reinterpret_cast<u64*>(smem);              // [kMaxUnits]
  u64* empty = full + kMaxUnits;                         // [kMaxUnits]
  unsigned* s_cnt = reinterpret_cast<unsigned*>(empty + kMaxUnits);  // [8]
  unsigned* s_conf = s_cnt + 8;                          // [368] (361 used)
  float* s_edges = reinterpret_cast<float*>(s_conf + 368);           // [AWX_MAX_ECE_BINS + 4] (keeps the u64 arrays 8-byte aligned)
  u64* w_cnt64 = reinterpret_cast<u64*>(s_edges + AWX_MAX_ECE_BINS + 4);  // [warps][nb]
  u64* w_cor64 = w_cnt64 + kConsWarps * nb;
  u64* w_sum64 = w_cor64 + kConsWarps * nb;
  unsigned* w_cc = reinterpret_cast<unsigned*>(w_sum64 + kConsWarps * nb);  // [warps][nb] count | correct << 16
  unsigned* w_lo = w_cc + kConsWarps * nb;
  unsigned* w_hi = w_lo + kConsWarps * nb;
  unsigned* s_auroc = w_hi + kConsWarps * nb;            // [2*NB]
  float* units = reinterpret_cast<float*>(smem + v2_ring_offset(nb, NB));
  {
    unsigned* w = s_cnt;
    const int words = 8 + 368;
    for (int i = threadIdx.x; i < words; i += kV2Threads) w[i] = 0u;
    unsigned* w2 = reinterpret_cast<unsigned*>(w_cnt64);
    const int words2 = kConsWarps * nb * 9 + 2 * NB;
    for (int i = threadIdx.x; i < words2; i += kV2Threads) w2[i] = 0u;
    for (int i = threadIdx.x; i <= nb; i += kV2Threads) s_edges[i] = p.edges[i];
    if (threadIdx.x == 0) {
      for (int u = 0; u < NU; ++u) {
        mbar_init(full + u, 1);
        mbar_init(empty + u, kConsWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
  }
  __syncthreads();

  int t;  // read %tid.x once (the compiler otherwise re-reads the special register in the loop)
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  const int warp = t >> 5, lane = t & 31;
  const long long HW = p.HW;
  const long long tpi = (HW + kTP - 1) / kTP;  // tiles per image
  const long long ntiles = p.B * tpi;

  if (warp == kConsWarps) {
    // ------------------------------------------------------------------ producer warp
    unsigned u = 0, ph = 0;
    long long img = blockIdx.x / tpi, tin = blockIdx.x - img * tpi;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long p0 = tin * kTP;
      const unsigned npx = (unsigned)((HW - p0) < kTP ? (HW - p0) : kTP);
#pragma unroll
      for (int m = 0; m < (ENS ? 2 : 1); ++m) {
        mbar_wait(empty + u, ph ^ 1u);
        if (lane == 0) mbar_expect_tx(full + u, kC * npx * 4u);
        __syncwarp();
        if (lane < kC) {
          const float* src = (m == 0 ? p.a : p.b) + (img * kC + lane) * HW + p0;
          bulk_load(units + (size_t)u * kUnitFloats + lane * kTP, src, npx * 4u, full + u);
        }
        if (++u == (unsigned)NU) {
          u = 0;
          ph ^= 1u;
        }
      }
      tin += gridDim.x;
      while (tin >= tpi) {
        tin -= tpi;
        ++img;
      }
    }
    return;
  }

  // -------------------------------------------------------------------- consumer warps
  const float2 nz = splat(negzero);
  const float2 w0 = splat(p.w0), w1 = splat(p.w1), half = splat(0.5f);
  const float2 l2e = splat(kLog2e), kz = splat(p.kz), eps = splat(kEps);
  const float T = p.T;
  const int ignore = p.ignore_index;
  const bool lab_u8 = FAST == 1 || (FAST == 0 && p.label_mode == AWX_LABEL_U8);
  const int div_mode = DIV >= 0 ? DIV : p.div_mode;
  unsigned n_correct = 0, n_bad = 0, n_ambig = 0;
  unsigned u = 0, ph = 0, since_flush = 0;
  unsigned* my_cc = w_cc + warp * nb;
  unsigned* my_lo = w_lo + warp * nb;
  unsigned* my_hi = w_hi + warp * nb;
  const float* my_units = units + t;
  long long img = blockIdx.x / tpi, tin = blockIdx.x - img * tpi;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long p0 = tin * kTP;
    const int npx = (int)((HW - p0) < kTP ? (HW - p0) : kTP);
    const bool act = t < npx;
    const long long li = img * HW + p0 + t;
    // label as a 32-bit int; int64 values outside the int range can only be "bad" labels
    int lab = ignore;
    if (have_labels && act) {
      if (lab_u8) {
        lab = static_cast<const uint8_t*>(p.labels)[li];
      } else {
        const long long l = static_cast<const long long*>(p.labels)[li];
        lab = (l >= -2147483647LL && l <= 2147483647LL) ? (int)l : (ignore == -2 ? -3 : -2);
      }
    }
    // ---- pull the pixel's 19 (+19) values into registers, release the ring units at once
    float2 a[NP], b[ENS ? NP : 1];
    {
      mbar_wait(full + u, ph);
      const float* s = my_units + (size_t)u * kUnitFloats;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        a[i].x = s[(2 * i) * kTP];
        a[i].y = (2 * i + 1 < kC) ? s[(2 * i + 1) * kTP] : kDummy;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + u);
      if (++u == (unsigned)NU) {
        u = 0;
        ph ^= 1u;
      }
    }
    if (ENS) {
      mbar_wait(full + u, ph);
      const float* s = my_units + (size_t)u * kUnitFloats;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        b[ENS ? i : 0].x = s[(2 * i) * kTP];
        b[ENS ? i : 0].y = (2 * i + 1 < kC) ? s[(2 * i + 1) * kTP] : kDummy;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + u);
      if (++u == (unsigned)NU) {
        u = 0;
        ph ^= 1u;
      }
    }
    if (!act) {  // tail tile: stale ring contents stand in for this lane; neutralise them
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        a[i] = make_float2(0.f, (2 * i + 1 < kC) ? 0.f : kDummy);
        if (ENS) b[ENS ? i : 0] = a[i];
      }
    }
    if (FAST == 0 && p.debug_skip) {  // dev: measure the TMA ring alone (generic kernels only)
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) acc += a[i].x + a[i].y + (ENS ? b[ENS ? i : 0].x + b[ENS ? i : 0].y : 0.f);
      if (acc == 1234.5678f) n_bad += 1;
      tin += gridDim.x;
      while (tin >= tpi) {
        tin -= tpi;
        ++img;
      }
      continue;
    }

    // ---- P1: fused logits (exact), maxima, first arg-max
    float2 v[NP];
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (MODE == 1)
        v[i] = add2(fma2(w0, a[i], nz), fma2(w1, b[ENS ? i : 0], nz));  // three roundings, see header
      else if (MODE == 2)
        v[i] = mul2(add2(a[i], b[ENS ? i : 0]), half);
      else
        v[i] = a[i];
    }
    if (div_mode == 2) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i].x = __fdiv_rn(v[i].x, T);
        if (2 * i + 1 < kC) v[i].y = __fdiv_rn(v[i].y, T);
      }
    }
    float vmax = v[0].x, amax = a[0].x, bmax = b[0].x;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      vmax = fmaxf(vmax, fmaxf(v[i].x, v[i].y));
      if (ENS) {
        amax = fmaxf(amax, fmaxf(a[i].x, a[i].y));
        bmax = fmaxf(bmax, fmaxf(b[ENS ? i : 0].x, b[ENS ? i : 0].y));
      }
    }
    // first index attaining the max.  With a division by T > 0 still pending (div_mode 1) the
    // quotient can merge the max with the one or two floats just below it (at most 3 inputs share
    // a quotient); torch's argmax over the divided logits returns the first of those, so compare
    // against the smallest float whose quotient equals the max quotient.
    float vlo = vmax;
    if (div_mode == 1) {
      const float c1 = float_prev(vmax), c2 = float_prev(c1);
      // branch-free exact quotients; |vmax| outside [1e-25, 1e25] is routed to the scalar slow
      // path below (residual underflow / quotient overflow), so this block stays straight-line
      const float zmax = div_by_T(vmax, T, p.rT), z1 = div_by_T(c1, T, p.rT), z2 = div_by_T(c2, T, p.rT);
      vlo = (z1 == zmax) ? ((z2 == zmax) ? c2 : c1) : vmax;
    }
    int arg = 0;
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
      if (2 * i + 1 < kC) arg = (v[i].y >= vlo) ? 2 * i + 1 : arg;
      arg = (v[i].x >= vlo) ? 2 * i : arg;
    }

    // optional fused-logit output (bit exact: div_mode is 0 or 2 whenever it is requested)
    if (FAST == 0 && p.fused != nullptr && act) {
      float* fo = p.fused + img * kC * HW + p0 + t;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        fo[(2 * i) * HW] = v[i].x;
        if (2 * i + 1 < kC) fo[(2 * i + 1) * HW] = v[i].y;
      }
    }

    // ---- P3: softmax denominator of the fused logits (dominant term exactly 1)
    float sz;
    {
      float2 sz2 = splat(0.f);
      const float2 vm2 = splat(vmax);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 x = mul2(sub2(v[i], vm2), kz);
        sz2 = add2(sz2, make_float2(ex2_approx(x.x), (2 * i + 1 < kC) ? ex2_approx(x.y) : 0.f));
      }
      sz = sz2.x + sz2.y;
    }

    // ---- members: softmax sums, entropies, mean probabilities
    float mi = 0.f, js = 0.f, sa = 1.f, sb = 1.f;
    int marg = 0;
    if (ENS) {
      float2 sa2 = splat(0.f), sb2 = splat(0.f), ta2 = splat(0.f), tb2 = splat(0.f), xab2 = splat(0.f), xba2 = splat(0.f);
      const float2 am2 = splat(amax), bm2 = splat(bmax);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 da = sub2(a[i], am2), db = sub2(b[ENS ? i : 0], bm2);
        const float2 xa = mul2(da, l2e), xb = mul2(db, l2e);
        const float2 ea = make_float2(ex2_approx(xa.x), (2 * i + 1 < kC) ? ex2_approx(xa.y) : 0.f);
        const float2 eb = make_float2(ex2_approx(xb.x), (2 * i + 1 < kC) ? ex2_approx(xb.y) : 0.f);
        sa2 = add2(sa2, ea);
        sb2 = add2(sb2, eb);
        ta2 = fma2(ea, da, ta2);
        tb2 = fma2(eb, db, tb2);
        if (JS) {
          xab2 = fma2(ea, db, xab2);
          xba2 = fma2(eb, da, xba2);
        }
        a[i] = ea;
        b[ENS ? i : 0] = eb;
      }
      sa = sa2.x + sa2.y;
      sb = sb2.x + sb2.y;
      const float ta = ta2.x + ta2.y, tb = tb2.x + tb2.y;
      const float ra = rcp_approx(sa), rb = rcp_approx(sb);
      const float2 ka = splat(0.5f * ra), kb = splat(0.5f * rb);
      float2 hm2 = splat(0.f);
      float mlm = 0.f, mmax = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 m = fma2(a[i], ka, mul2(b[ENS ? i : 0], kb));
        const float2 me = add2(m, eps);
        hm2 = fma2(m, make_float2(lg2_approx(me.x), (2 * i + 1 < kC) ? lg2_approx(me.y) : 0.f), hm2);
        if (JS) {
          mlm += m.x > 0.f ? m.x * lg2_approx(m.x) : 0.f;
          if (2 * i + 1 < kC) mlm += m.y > 0.f ? m.y * lg2_approx(m.y) : 0.f;
        }
        a[i] = m;  // keep the mean probabilities for the arg-max below
        mmax = fmaxf(mmax, fmaxf(m.x, m.y));
      }
#pragma unroll
      for (int i = NP - 1; i >= 0; --i) {
        if (2 * i + 1 < kC) marg = (a[i].y == mmax) ? 2 * i + 1 : marg;
        marg = (a[i].x == mmax) ? 2 * i : marg;
      }
      const float lsa = kLn2 * lg2_approx(sa), lsb = kLn2 * lg2_approx(sb);
      const float ceps = (float)kC * kEps;
      const float ha = lsa - ta * ra - ceps;
      const float hb = lsb - tb * rb - ceps;
      mi = -kLn2 * (hm2.x + hm2.y) - 0.5f * (ha + hb);
      if (JS) {
        const float xab = xab2.x + xab2.y, xba = xba2.x + xba2.y;
        const float kas = 0.5f * ra, kbs = 0.5f * rb;
        const float mlp = kas * ta + kbs * xba - lsa;
        const float mlq = kbs * tb + kas * xab - lsb;
        js = kLn2 * mlm - 0.5f * (mlp + mlq);
      }
    }

    // ---- confidence, ECE bin, slow paths
    PixOut o;
    o.pred = arg;
    o.mi = mi;
    o.js = js;
    o.mpred = marg;
    o.ambig = 0;
    {
      const bool range_ok = div_mode != 1 || (fabsf(vmax) > 1e-25f && fabsf(vmax) < 1e25f);
      const bool sane = range_ok && isfinite(sz) && (!ENS || (isfinite(sa) && isfinite(sb) && isfinite(mi)));
      float conf = __frcp_rn(sz);
      int bin = ece_bin_fast(conf, s_edges, nb);
      if (act && !sane) {
        const float* ga = p.a + img * kC * HW + p0 + t;
        slow_pixel<ENS, JS>(p, s_edges, ga, ENS ? p.b + img * kC * HW + p0 + t : nullptr, o);
      } else {
        if (act && bin >= 0) {
          const float tol = conf * 1.9e-6f;  // 16 ulp
          const bool near_lo = bin > 0 && (conf - s_edges[bin]) <= tol;
          const bool near_hi = bin < nb - 1 && (s_edges[bin + 1] - conf) <= tol;
          if (near_lo || near_hi) {
            const float* ga = p.a + img * kC * HW + p0 + t;
            conf = exact_confidence(ga, ENS ? p.b + img * kC * HW + p0 + t : nullptr, HW, kC, MODE == 2, p.w0, p.w1,
                                    div_mode, T, s_edges, nb, &o.ambig);
            bin = ece_bin_fast(conf, s_edges, nb);
          }
        }
        o.conf = conf;
        o.bin = bin;
      }
    }

    if (FAST == 0 && act) {
      if (p.pred) {
        if (p.pred_dtype == AWX_PRED_U8)
          static_cast<uint8_t*>(p.pred)[li] = (uint8_t)o.pred;
        else
          static_cast<long long*>(p.pred)[li] = o.pred;
      }
      if (p.conf) p.conf[li] = o.conf;
      if (ENS && p.mi) p.mi[li] = o.mi;
      if (ENS && JS && p.js) p.js[li] = o.js;
    }

    // ---- statistics
    if (have_labels) {
      if (since_flush + 32u > kFlushPixels) {  // warp-uniform
        __syncwarp();
        for (int i = lane; i < nb; i += 32) {
          const unsigned cc = my_cc[i];
          w_cnt64[warp * nb + i] += cc & 0xffffu;
          w_cor64[warp * nb + i] += cc >> 16;
          w_sum64[warp * nb + i] += ((u64)my_hi[i] << 16) + my_lo[i];
          my_cc[i] = my_lo[i] = my_hi[i] = 0u;
        }
        __syncwarp();
        since_flush = 0;
      }
      since_flush += 32u;
      const bool valid = act && lab != ignore;
      const bool correct = valid && lab == o.pred;
      n_correct += correct;
      int ckey = -1, akey = -1;
      if (valid) {
        // confusion index as torch evaluates targets*C + predictions (uint8 product wraps mod 256)
        const int idx = (lab_u8 ? ((lab * kC) & 0xff) : lab * kC) + o.pred;
        const bool inside = lab_u8 ? (idx < kC * kC) : (lab >= 0 && lab < kC);
        if (inside)
          ckey = idx;
        else
          ++n_bad;
        n_ambig += o.ambig;
        if (ENS && NB > 0) {
          float qv = floorf(o.mi * p.auroc_scale);
          qv = is_nan(qv) ? 0.f : qv;
          akey = (lab != o.mpred ? 0 : NB) + (int)fminf(fmaxf(qv, 0.f), (float)(NB - 1));
        }
      }
      const int bin = valid ? o.bin : -1;
      const unsigned fx = bin >= 0 ? __float2uint_rz(o.conf * 2147483648.f) : 0u;
      // one warp-uniformity test for all three histograms (piecewise-constant real data)
      const int key = (ckey + 1) | ((akey + 1) << 9) | ((bin + 1) << 23);
      int same;
      __match_all_sync(0xffffffffu, key, &same);
      if (same) {
        if (key != 0) {
          const unsigned lo = __reduce_add_sync(0xffffffffu, fx & 0xffffu);
          const unsigned hi = __reduce_add_sync(0xffffffffu, fx >> 16);
          if (lane == 0) {
            if (ckey >= 0) atomicAdd(s_conf + ckey, 32u);
            if (akey >= 0) atomicAdd(s_auroc + akey, 32u);
            if (bin >= 0) {
              atomicAdd(my_cc + bin, 32u | (correct ? (32u << 16) : 0u));
              atomicAdd(my_lo + bin, lo);
              atomicAdd(my_hi + bin, hi);
            }
          }
        }
      } else {
        if (ckey >= 0) atomicAdd(s_conf + ckey, 1u);
        if (akey >= 0) atomicAdd(s_auroc + akey, 1u);
        if (bin >= 0) {
          atomicAdd(my_cc + bin, 1u | (correct ? 0x10000u : 0u));
          atomicAdd(my_lo + bin, fx & 0xffffu);
          atomicAdd(my_hi + bin, fx >> 16);
        }
      }
    }
    tin += gridDim.x;
    while (tin >= tpi) {
      tin -= tpi;
      ++img;
    }
  }

  if (!have_labels) return;
  {
    unsigned vv[3] = {n_correct, n_bad, n_ambig};
    const int slot[3] = {AWX_CNT_CORRECT, AWX_CNT_BAD_LABEL, AWX_CNT_ECE_AMBIG};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const unsigned s = __reduce_add_sync(0xffffffffu, vv[k]);
      if (lane == 0 && s) atomicAdd(&s_cnt[slot[k]], s);
    }
  }
  // consumer-only barrier (the producer warp has already left)
  asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory");
  unsigned long long* bins = p.bins;
  // the remaining counters follow from the histograms: valid = sum(confusion) + bad,
  // ensemble-wrong = sum(auroc_pos), no-bin = valid - sum(ece_count); pixels are added by the host
  unsigned part_conf = 0, part_pos = 0, part_ece = 0;
  for (int i = t; i < kC * kC; i += kCons) {
    const unsigned c = s_conf[i];
    part_conf += c;
    if (c) atomicAdd(bins + p.lay.confusion + i, (u64)c);
  }
  for (int i = t; i < 2 * NB; i += kCons) {
    const unsigned c = s_auroc[i];
    if (i < NB) part_pos += c;
    if (c) atomicAdd(bins + (i < NB ? p.lay.auroc_pos + i : p.lay.auroc_neg + (i - NB)), (u64)c);
  }
  for (int i = t; i < nb; i += kCons) {
    u64 cnt = 0, cor = 0, sum = 0;
    for (int w = 0; w < kConsWarps; ++w) {
      const unsigned cc = w_cc[w * nb + i];
      cnt += w_cnt64[w * nb + i] + (cc & 0xffffu);
      cor += w_cor64[w * nb + i] + (cc >> 16);
      sum += w_sum64[w * nb + i] + ((u64)w_hi[w * nb + i] << 16) + w_lo[w * nb + i];
    }
    part_ece += (unsigned)cnt;
    if (cnt) {
      atomicAdd(bins + p.lay.ece_count + i, cnt);
      if (cor) atomicAdd(bins + p.lay.ece_correct + i, cor);
      atomicAdd(bins + p.lay.ece_conf_hi + i, sum >> 32);
      atomicAdd(bins + p.lay.ece_conf_lo + i, sum & 0xffffffffull);
    }
  }
  {
    const unsigned sc = __reduce_add_sync(0xffffffffu, part_conf);
    const unsigned sp = __reduce_add_sync(0xffffffffu, part_pos);
    const unsigned se = __reduce_add_sync(0xffffffffu, part_ece);
    if (lane == 0) {
      if (sc) atomicAdd(&s_cnt[AWX_CNT_VALID], sc);
      if (sp) atomicAdd(&s_cnt[AWX_CNT_ENS_WRONG], sp);
      if (se) atomicAdd(&s_cnt[AWX_CNT_NO_BIN], se);  // holds sum(ece_count) until the fix-up below
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory");
  if (t == 0) {
    const unsigned bad = s_cnt[AWX_CNT_BAD_LABEL];
    const unsigned valid = s_cnt[AWX_CNT_VALID] + bad;
    const unsigned nobin = valid - s_cnt[AWX_CNT_NO_BIN];
    if (valid) atomicAdd(bins + p.lay.counters + AWX_CNT_VALID, (u64)valid);
    if (s_cnt[AWX_CNT_CORRECT]) atomicAdd(bins + p.lay.counters + AWX_CNT_CORRECT, (u64)s_cnt[AWX_CNT_CORRECT]);
    if (bad) atomicAdd(bins + p.lay.counters + AWX_CNT_BAD_LABEL, (u64)bad);
    if (s_cnt[AWX_CNT_ECE_AMBIG]) atomicAdd(bins + p.lay.counters + AWX_CNT_ECE_AMBIG, (u64)s_cnt[AWX_CNT_ECE_AMBIG]);
    if (s_cnt[AWX_CNT_ENS_WRONG]) atomicAdd(bins + p.lay.counters + AWX_CNT_ENS_WRONG, (u64)s_cnt[AWX_CNT_ENS_WRONG]);
    if (nobin) atomicAdd(bins + p.lay.counters + AWX_CNT_NO_BIN, (u64)nobin);
    if (blockIdx.x == 0) atomicAdd(bins + p.lay.counters + AWX_CNT_PIXELS, (u64)(p.B * p.HW));
  }
}

template <int MODE, bool JS, int FAST, int DIV>
int launch_v2(const ScoreParams& p, cudaStream_t stream) {
  auto kern = score_v2_kernel<MODE, JS, FAST, DIV>;
  int dev = 0, max_smem = 0;
  AWX_CUDA(cudaGetDevice(&dev));
  AWX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = v2_ring_offset(p.nb, p.auroc_bins);
  int nu = (int)(((size_t)max_smem - fixed) / kUnitBytes);
  if (nu > kMaxUnits) nu = kMaxUnits;
  AWX_REQUIRE(nu >= 2, AWX_E_UNSUPPORTED, "awx_score v2: histograms leave no room for the TMA ring");
  const size_t smem = (size_t)nu * kUnitBytes + fixed;
  AWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = p.B * ((p.HW + kTP - 1) / kTP);
  long long blocks = sm_count();
  if (blocks > ntiles) blocks = ntiles;
  kern<<<(unsigned)blocks, kV2Threads, smem, stream>>>(p, nu, -0.0f);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

template <int MODE, int FAST>
int launch_v2_fast(const ScoreParams& p, cudaStream_t stream) {
  if (p.div_mode == 0) return launch_v2<MODE, false, FAST, 0>(p, stream);
  if (p.div_mode == 1) return launch_v2<MODE, false, FAST, 1>(p, stream);
  return launch_v2<MODE, false, 0, -1>(p, stream);  // exact division everywhere (T <= 0): generic kernel
}

template <int MODE>
int launch_v2_mode(const ScoreParams& p, bool js, cudaStream_t stream) {
  const bool maps = p.pred || p.fused || p.conf || p.mi || p.js;
  if (!maps && p.labels != nullptr && !js && !p.debug_skip)
    return p.label_mode == AWX_LABEL_U8 ? launch_v2_fast<MODE, 1>(p, stream) : launch_v2_fast<MODE, 2>(p, stream);
  if (MODE != 0 && js) return launch_v2<MODE, (MODE != 0), 0, -1>(p, stream);
  return launch_v2<MODE, false, 0, -1>(p, stream);
}

}  // namespace

bool score_v2_supported(const ScoreParams& p) {
  if (p.C != kC || p.strategy == AWX_FUSE_MAXCONF) return false;
  if (p.HW % 4 != 0) return false;
  if (p.B * p.HW >= (1LL << 32)) return false;  // CTA-level counters are 32 bit
  if (((uintptr_t)p.a & 15) != 0 || ((uintptr_t)p.b & 15) != 0) return false;
  if (p.labels && p.label_mode == AWX_LABEL_I64 && ((uintptr_t)p.labels & 7) != 0) return false;
  return true;
}

int launch_score_v2(const ScoreParams& p, bool ens, bool js, cudaStream_t stream) {
  if (!ens) return launch_v2_mode<0>(p, false, stream);
  if (p.strategy == AWX_FUSE_MEAN) return launch_v2_mode<2>(p, js, stream);
  return launch_v2_mode<1>(p, js, stream);
}

}  // namespace awx
