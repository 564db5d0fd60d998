// awx_score, kernel v2 (C == 19): TMA-staged, register-resident, packed-fp32 scoring.
//
// One persistent CTA per SM: the last warp is the producer, warps 0..CW-1 (480 or 608 threads) consume.
//   producer  walks the CTA's tiles (32*CW consecutive pixels of one image) and, per tile, issues 19 bulk
//             async copies per member (cp.async.bulk, one 1.9 KB plane segment each) into the next free
//             unit of a shared-memory ring, completing on that unit's mbarrier.  A unit is one tile of every
//             member (36 / 46 KB for one member, 73 KB for two); the ring holds 4 (one member) or 2 (two
//             members) of them, i.e. ~146-185 KB per SM in flight independent of what the consumers do.
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
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstring>

#include "score_common.cuh"
#include "tma_ring.cuh"

#ifndef AWX_V2_GROUPS
#define AWX_V2_GROUPS 3
#endif
#ifndef AWX_V2_TMAP
#define AWX_V2_TMAP 0   // 2-D tensor-map copies: built and measured (1-2 % slower than the bulk copies), kept as a build option
#endif

namespace awx {
using namespace score_detail;
namespace {

constexpr int kC = 19;
// Geometry of one kernel variant: CW consumer warps (one pixel per consumer thread) + 1 producer warp.
//   CW = 15: 512 threads, 128 registers per thread, 36 KB ring units (the generic kernels)
//   CW = 19: 640 threads,  96 registers per thread, 46 KB ring units (bins-only kernels: more warps to
//            blend the MUFU / FMA / ALU phases of different warps on each scheduler)
template <int CW>
struct Geo {
  static constexpr int kTP = 32 * CW;            // pixels per tile
  static constexpr int kCons = kTP;              // consumer threads
  static constexpr int kThreads = kCons + 32;    // + producer warp
  static constexpr int kUnitFloats = kC * kTP;
  static constexpr int kUnitBytes = kUnitFloats * 4;
  // Consumer warps form kGroups GROUPS, each waiting for its own CHUNK (pixel range) of a ring unit on its own pair
  // of barriers.  The producer issues the chunks of a tile one after the other, so they land a third of a tile's
  // transfer time apart and the groups run out of phase: while one pulls its pixels out of shared memory (LDS),
  // another is in its exponentials (MUFU) and the third in the packed arithmetic.  With ONE barrier per unit all
  // warps started every tile together and queued for the same pipe (measured: forcing lock-step with a named
  // barrier costs 12 %; mio_throttle is the top stall either way).
  static constexpr int kGroups = (AWX_V2_GROUPS <= CW) ? AWX_V2_GROUPS : 1;
  __host__ __device__ static constexpr int group_first_warp(int g) {   // groups differ by at most one warp
    return g * (CW / kGroups) + (g < CW % kGroups ? g : CW % kGroups);
  }
  __host__ __device__ static constexpr int group_warps(int g) { return CW / kGroups + (g < CW % kGroups ? 1 : 0); }
  // Build option AWX_V2_TMAP=1: equal groups of at most 256 pixels are fetched with ONE 2-D tensor-map copy per member
  // and chunk (box = kBoxW pixels x 19 planes, SASS UTMALDG) instead of 19 bulk copies -- 2 * kGroups copy
  // instructions per tile instead of 38 * kGroups, and (every completed copy wakes every warp sleeping on an mbarrier
  // of the CTA) ~150 fewer SYNCS / NANOSLEEP / BRA per pixel.  Measured on B200 it is nevertheless 1-2 % SLOWER in
  // the bench step than the bulk copies (profiles/r2h_ab_bench.md), so the default is off.  The box lands densely: a
  // unit is then [group][member][plane][kBoxW] instead of [member][plane][kTP].
  static constexpr bool kTensorMap = AWX_V2_TMAP != 0 && CW % kGroups == 0 && kTP / kGroups <= 256;
  static constexpr int kBoxW = kTP / kGroups;
  __host__ __device__ static constexpr int group_of_warp(int w) {
    int g = 0;
    for (int i = 1; i < kGroups; ++i) g += (w >= group_first_warp(i)) ? 1 : 0;
    return g;
  }
};
constexpr int kMaxGroups = 5;
constexpr int kMaxUnits = 4;                    // ring depth for one member (2 units of both members for an ensemble):
                                                // 1.5, 2 and 2.5 tiles measure the same; the shared memory is better
                                                // spent on conflict-free statistics
constexpr int kFastEceBins = 15, kFastAurocBins = 4096;  // the streaming evaluator's configuration
constexpr int kSingleWarps = 19;                // consumer warps of the bins-only single-member kernels
constexpr unsigned kFlushPixels = 60000;        // per-warp confidence sums are flushed before 2^16 pixels
constexpr int kEceRep = 16;                     // replicas (lane & 15) of the per-warp confidence-sum words

// shared-memory reduction on a 32-bit shared address (no generic -> shared conversion at the use site)
__device__ __forceinline__ void red_add(uint32_t addr, unsigned v) {
  asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// the same, predicated on key >= 0 inside the instruction (no branch / reconvergence around it)
__device__ __forceinline__ void red_add_if(int key, uint32_t addr, unsigned v) {
  asm volatile("{\n .reg .pred p;\n setp.ge.s32 p, %0, 0;\n @p red.shared.add.u32 [%1], %2;\n}" ::"r"(key), "r"(addr), "r"(v)
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
__device__ __forceinline__ float2 fma2(const float2 a, const float2 b, const float2 c) {
  float2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(bits(d)) : "l"(bits(a)), "l"(bits(b)), "l"(bits(c)));
  return d;
}
__device__ __forceinline__ float2 splat(float x) { return make_float2(x, x); }

__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// Arg-max of 19 values held as class pairs, with the runner-up test for free.  Bit (18 - c) of the result is
// set iff x_c < thr: the packed subtraction x - thr handles two classes per instruction and one funnel shift
// per class collects the sign bits (x == thr gives +0: not below).  With thr a little under the maximum,
// the complement is the set of classes that may attain it: one bit set -> that class is the arg-max beyond
// doubt (count-leading-zeros gives its index); several -> resolve_ties (score_common.cuh).  19 + 10 + ~6
// instructions instead of the 38 of a compare / select chain, which cannot see near ties at all.
template <int NP>
__device__ __forceinline__ unsigned below_mask(const float2 (&x)[NP], float thr) {
  const float2 nt = splat(-thr);
  unsigned m0 = 0u, m1 = 0u;  // classes 0..9 and 10..18: two independent shift chains
#pragma unroll
  for (int i = 0; i < NP; ++i) {
    const float2 d = add2(x[i], nt);
    if (i < 5) {
      m0 = __funnelshift_l(__float_as_uint(d.x), m0, 1);
      m0 = __funnelshift_l(__float_as_uint(d.y), m0, 1);
    } else {
      m1 = __funnelshift_l(__float_as_uint(d.x), m1, 1);
      if (2 * i + 1 < kC) m1 = __funnelshift_l(__float_as_uint(d.y), m1, 1);
    }
  }
  return (m0 << 9) | m1;
}
constexpr unsigned kClassBits = (1u << kC) - 1u;

// (lo, hi] bin of conf; edges are within an ulp of i/nb so the guess is off by at most one
__device__ __forceinline__ int ece_bin_fast(float conf, const float* e, int nb) {
  int b = min(max((int)ceilf(conf * (float)nb) - 1, 0), nb - 1);
  b -= (b > 0 && !(conf > e[b])) ? 1 : 0;
  b += (b < nb - 1 && conf > e[b + 1]) ? 1 : 0;
  return (conf > e[b] && conf <= e[b + 1]) ? b : -1;
}

// byte offset of the TMA ring inside dynamic shared memory (everything before it is bookkeeping)
__host__ __device__ inline size_t v2_ring_offset(int cons_warps, int nb, int NB) {
  size_t o = 2 * kMaxUnits * kMaxGroups * sizeof(u64) + (AWX_NUM_COUNTERS + 368) * 4 + (AWX_MAX_ECE_BINS + 4) * 4;
  o += (size_t)cons_warps * nb * (2 + 2 * kEceRep) * 4 + (size_t)2 * NB * 4;
  return (o + 127) & ~(size_t)127;
}

// Slow path for one pixel straight from global memory (NaN / inf logits): the scalar v1 code.
template <bool ENS, bool JS>
__device__ __noinline__ void slow_pixel(const ScoreParams& p, const float* s_edges, const float* ga, const float* gb,
                                        long long lab, PixOut& o) {
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
  float w0 = p.w0, w1 = p.w1;
  if (ENS && p.strategy == AWX_FUSE_MAXCONF) {  // member with the larger max-softmax, strict > (model.py:449-455)
    float sa = 0.f, sb = 0.f;
    for (int c = 0; c < kC; ++c) {
      sa += ex2_approx((a[c] - amax) * kLog2e);
      sb += ex2_approx((b[ENS ? c : 0] - bmax) * kLog2e);
    }
    w0 = __frcp_rn(sa) > __frcp_rn(sb) ? 1.f : 0.f;
    w1 = 1.f - w0;
  }
  score_pixel<kC, ENS, JS>(a, b, kC, p, s_edges, ga, gb, w0, w1, amax, bmax, lab, o);
}

// MODE: 0 single member, 1 weighted average, 2 mean, 3 max-confidence (weighted average with per-pixel weights
//       (1,0) / (0,1): 1*a + 0*b is the reference's pick*l1 + (1-pick)*l2 operation for operation).
// FAST: 0 generic (runtime label dtype, optional per-pixel maps, any edges), 1 uint8 labels and bins
//       only, 2 int64 labels and bins only -- the streaming-evaluation configurations, with every
//       map / dtype branch compiled out and the ECE edges known to be linspace(0,1,nb+1).
// DIV:  -1 runtime p.div_mode (generic kernels), 0 / 1 compile-time (bins-only kernels).
// CW:   consumer warps (Geo<CW>).
//
// Consumer thread t owns pixel t of the tile.  Its 19 class values are held as 10 float2 PAIRS OF
// CLASSES (2i, 2i+1); the 20th slot is a finite "never wins" dummy (-1e30) whose exponentials are
// exactly 0.  Element-wise work runs on FADD2/FMUL2/FFMA2 over class pairs; maxima and arg-maxima are
// scalar (no packed min/max exists).
//
// Shifted exponents.  Every softmax is evaluated as e'_c = ex2(fma(x_c, k, -fl(xmax*k))): ONE packed
// FMA per class pair instead of subtract + multiply.  The rounded product fl(xmax*k) shifts all
// exponents of a pixel by the same delta = xmax*k - fl(xmax*k) (|delta| <= ulp/2), i.e. e'_c =
// 2^delta e_c and S' = 2^delta S:
//   * probabilities e'_c / S' and log2 p_c = t_c - lg2 S' do not see delta at all, so the entropies
//     H = ln2 (lg2 S' - sum e'_c t_c / S') and the mean probabilities are unchanged;
//   * the confidence 1/S needs S itself: delta = fma(xmax, k, -fl(xmax*k)) is exact, and
//     conf = rcp(S') (1 + delta ln2)   (2^delta to first order; delta^2 < 1e-11).
//
// Shared memory: [mbarriers | counters | confusion | edges | per-warp ECE words | AUROC | ring].
template <int MODE, bool JS, int FAST, int DIV, int CW>
__global__ void __launch_bounds__(Geo<CW>::kThreads, 1)
    score_v2_kernel(const __grid_constant__ ScoreParams p, const int NU, const float negzero,
                    const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b) {
  using G = Geo<CW>;
  constexpr int kTP = G::kTP, kCons = G::kCons, kConsWarps = CW, kV2Threads = G::kThreads;
  constexpr bool ENS = MODE != 0;
  // one ring unit = one tile of every member (19 or 38 planes): consumers wait once and release once per tile
  constexpr int kUnitFloats = (ENS ? 2 : 1) * G::kUnitFloats;
  constexpr uint32_t kUnitBytes = 4u * kUnitFloats;
  constexpr bool TM = G::kTensorMap;
  constexpr int kM = ENS ? 2 : 1;
  constexpr int kPS = TM ? G::kBoxW : kTP;      // plane stride inside a unit (floats)
  constexpr int kMO = kC * kPS;                 // offset of the second member's planes (floats)
  constexpr int NP = (kC + 1) / 2;  // 10 class pairs
  constexpr float kDummy = -1e30f;
  // bins-only kernels: 15 ECE bins and 4096 (ensemble) / 0 (single) AUROC bins are compile-time constants, so
  // every shared-memory offset below is an immediate (launch_v2_mode checks the configuration)
  const int nb = FAST != 0 ? kFastEceBins : p.nb;
  const int NB = FAST != 0 ? (ENS ? kFastAurocBins : 0) : p.auroc_bins;
  const bool have_labels = FAST != 0 || p.labels != nullptr;
  extern __shared__ __align__(128) unsigned char smem[];
  u64* full = reinterpret_cast<u64*>(smem);              // [kMaxUnits][kMaxGroups]
  u64* empty = full + kMaxUnits * kMaxGroups;            // [kMaxUnits][kMaxGroups]
  unsigned* s_cnt = reinterpret_cast<unsigned*>(empty + kMaxUnits * kMaxGroups);  // [AWX_NUM_COUNTERS]
  unsigned* s_conf = s_cnt + AWX_NUM_COUNTERS;                          // [368] (361 used)
  float* s_edges = reinterpret_cast<float*>(s_conf + 368);           // [AWX_MAX_ECE_BINS + 4] (keeps the u64 arrays 8-byte aligned)
  // per-warp ECE words.  Counts take the value 1 (the hardware aggregates lanes hitting the same word:
  // ATOMS.POPC.INC, ~3 wavefronts); the confidence sums carry per-lane values, and a shared-memory atomic add
  // whose lanes collide on one address is serialised lane by lane (measured: 28 wavefronts per instruction with
  // ~10 lanes per bin), so every (warp, bin) sum has kEceRep replicas selected by lane & 15.
  unsigned* w_cnt = reinterpret_cast<unsigned*>(s_edges + AWX_MAX_ECE_BINS + 4);  // [warps][nb]
  unsigned* w_cor = w_cnt + kConsWarps * nb;                                      // [warps][nb]
  unsigned* w_lo = w_cor + kConsWarps * nb;                                       // [warps][nb][kEceRep] low 16 bits of conf * 2^31
  unsigned* w_hi = w_lo + kConsWarps * nb * kEceRep;                              // [warps][nb][kEceRep] high 15 bits
  unsigned* s_auroc = w_hi + kConsWarps * nb * kEceRep;  // [2*NB]
  float* units = reinterpret_cast<float*>(smem + v2_ring_offset(kConsWarps, nb, NB));
  {
    unsigned* w = s_cnt;
    const int words = AWX_NUM_COUNTERS + 368;
    for (int i = threadIdx.x; i < words; i += kV2Threads) w[i] = 0u;
    unsigned* w2 = w_cnt;
    const int words2 = kConsWarps * nb * (2 + 2 * kEceRep) + 2 * NB;
    for (int i = threadIdx.x; i < words2; i += kV2Threads) w2[i] = 0u;
    for (int i = threadIdx.x; i <= nb; i += kV2Threads) s_edges[i] = p.edges[i];
    if (threadIdx.x == 0) {
      for (int u = 0; u < NU; ++u) {
        for (int g = 0; g < G::kGroups; ++g) {
          mbar_init(full + u * kMaxGroups + g, 1);
          mbar_init(empty + u * kMaxGroups + g, G::group_warps(g));
        }
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

  if (threadIdx.x >= kCons) {  // (threadIdx.x, not the opaque copy of %tid.x: the compiler must see a warp-uniform branch)
    // ------------------------------------------------------------------ producer warp
    unsigned u = 0, ph = 0;
    long long img = blockIdx.x / tpi, tin = blockIdx.x - img * tpi;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long p0 = tin * kTP;
      const int npx = (int)((HW - p0) < kTP ? (HW - p0) : kTP);
#pragma unroll
      for (int g = 0; g < G::kGroups; ++g) {
        // chunk g of the unit: the pixels of consumer group g (nothing for a group past the end of the image, whose
        // barrier still completes: expect_tx of 0 bytes is a plain arrival)
        const int c0 = 32 * G::group_first_warp(g), cn = 32 * G::group_warps(g);
        const int n = min(max(npx - c0, 0), cn);
        u64* fb = full + u * kMaxGroups + g;
        mbar_wait(empty + u * kMaxGroups + g, ph ^ 1u);
        // ONE thread issues the 19 (38) copies of the chunk from warp-uniform operands: the copy instruction takes
        // uniform registers, and per-lane addresses would make the compiler broadcast every lane's operands one
        // after the other (ELECT / 4 x R2UR / UBLKCP per copy: ~190 instructions per unit, which kept this warp
        // busy 70 % of the time)
        if (elect_one()) {
          if (TM) {
            // out-of-image pixels of a tail tile are zero-filled by the copy engine and still count as bytes
            mbar_expect_tx(fb, n > 0 ? (unsigned)(kM * kC * G::kBoxW * 4) : 0u);
            if (n > 0) {
              float* dst = units + (size_t)u * kUnitFloats + (size_t)g * (kM * kMO);
              tma_load_2d(dst, &tmap_a, (int)(p0 + c0), (int)(img * kC), fb);
              if (ENS) tma_load_2d(dst + kMO, &tmap_b, (int)(p0 + c0), (int)(img * kC), fb);
            }
          } else {
            mbar_expect_tx(fb, (ENS ? 2u : 1u) * kC * (unsigned)n * 4u);
            if (n > 0) {
              float* dst = units + (size_t)u * kUnitFloats + c0;
#pragma unroll
              for (int m = 0; m < (ENS ? 2 : 1); ++m) {
                const float* src = (m == 0 ? p.a : p.b) + img * kC * HW + p0 + c0;
#pragma unroll
                for (int c = 0; c < kC; ++c) bulk_load(dst + (m * kC + c) * kTP, src + c * HW, (unsigned)n * 4u, fb);
              }
            }
          }
        }
        __syncwarp();
      }
      if (++u == (unsigned)NU) {
        u = 0;
        ph ^= 1u;
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
  unsigned n_correct = 0, n_bad = 0, n_ambig = 0, n_pick = 0;
  unsigned u = 0, ph = 0, since_flush = 0;
  unsigned* my_lo = w_lo + warp * nb * kEceRep;
  unsigned* my_hi = w_hi + warp * nb * kEceRep;
  const uint32_t conf_sa = smem_u32(s_conf), auroc_sa = smem_u32(s_auroc);
  const uint32_t cnt_sa = smem_u32(w_cnt + warp * nb), cor_sa = smem_u32(w_cor + warp * nb);
  const uint32_t lo_sa = smem_u32(my_lo) + 4u * (lane & (kEceRep - 1)), hi_sa = smem_u32(my_hi) + 4u * (lane & (kEceRep - 1));
  // this warp's group: full[u][grp] at my_bar0 + 8 kMaxGroups u, empty[u][grp] kMaxUnits * kMaxGroups words further
  const uint32_t sbase = smem_u32(smem);
  const uint32_t my_bar0 = sbase + 8u * (uint32_t)G::group_of_warp(warp);
  const int grp = G::group_of_warp(warp);
  const uint32_t my_unit0 = sbase + (uint32_t)v2_ring_offset(kConsWarps, nb, NB) +
                            4u * (uint32_t)(TM ? grp * (kM * kMO) + (t - 32 * G::group_first_warp(grp)) : t);
  // B * HW < 2^32 (score_v2_supported): all pixel and tile indices of the consumers are 32 bit
  const unsigned HWu = (unsigned)HW, tpiu = (unsigned)tpi, ntu = (unsigned)ntiles;
  unsigned img = blockIdx.x / tpiu, tin = blockIdx.x - img * tpiu;
  const unsigned step_img = gridDim.x / tpiu, step_tin = gridDim.x - step_img * tpiu;

  for (unsigned tile = blockIdx.x; tile < ntu; tile += gridDim.x) {
    const unsigned p0 = tin * kTP;
    const unsigned rem = HWu - p0;
    const bool act = (unsigned)t < rem;  // rem >= kTP except in the tail tile of an image
    const unsigned li = img * HWu + p0 + t;
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
    // ---- pull the pixel's 19 (+19) values into registers, release the ring unit at once
    float2 a[NP], b[ENS ? NP : 1];
    uint32_t held_unit, held_bar;
    {
      mbar_wait_a(my_bar0 + 8u * kMaxGroups * u, ph);
      const uint32_t s = my_unit0 + u * kUnitBytes;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        a[i].x = lds_f32(s + (2 * i) * kPS * 4);
        a[i].y = (2 * i + 1 < kC) ? lds_f32(s + (2 * i + 1) * kPS * 4) : kDummy;
      }
      if (ENS) {
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          b[ENS ? i : 0].x = lds_f32(s + (kMO + (2 * i) * kPS) * 4);
          b[ENS ? i : 0].y = (2 * i + 1 < kC) ? lds_f32(s + (kMO + (2 * i + 1) * kPS) * 4) : kDummy;
        }
      }
      // max-confidence kernels keep the unit until the winning member's logits have been re-read (see P0)
      held_unit = s;
      held_bar = my_bar0 + 8u * kMaxGroups * (kMaxUnits + u);
      if (MODE != 3) {
        __syncwarp();
        if (lane == 0) mbar_arrive_a(held_bar);
      }
      if (++u == (unsigned)NU) {
        u = 0;
        ph ^= 1u;
      }
    }
    // Tail tile: lanes past the end of the image compute on stale ring contents; nothing they produce
    // is stored or counted (every store, slow path and histogram update below is guarded by act).
    if (FAST == 0 && p.debug_skip) {  // dev: measure the TMA ring alone (generic kernels only)
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) acc += a[i].x + a[i].y + (ENS ? b[ENS ? i : 0].x + b[ENS ? i : 0].y : 0.f);
      if (acc == 1234.5678f) n_bad += 1;
      if (MODE == 3) {
        __syncwarp();
        if (lane == 0) mbar_arrive_a(held_bar);
      }
      tin += gridDim.x;
      while (tin >= tpiu) {
        tin -= tpiu;
        ++img;
      }
      continue;
    }

    // ---- members: softmax sums, entropies, mean probabilities (log2 domain, shifted exponents).  Consumes a[] / b[]
    // (overwritten in place by the exponentials, then by the unnormalised mean probabilities).  Runs after the
    // fused-logit phases, except in the max-confidence kernels, where its sums decide which member is scored.
    float mi = 0.f, js = 0.f, sa = 1.f, sb = 1.f;
    float amax = a[0].x, bmax = b[0].x;
    int marg = 0;
    bool tie_u = false;  // several classes may attain the maximum of the mean probabilities
    auto members_phase = [&]() {
    if (ENS) {
      float2 sa2 = splat(0.f), sb2 = splat(0.f), ta2 = splat(0.f), tb2 = splat(0.f), xab2 = splat(0.f), xba2 = splat(0.f);
      const float2 ca2 = splat(-(amax * kLog2e)), cb2 = splat(-(bmax * kLog2e));
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 ta = fma2(a[i], l2e, ca2), tb = fma2(b[ENS ? i : 0], l2e, cb2);
        const float2 ea = make_float2(ex2_approx(ta.x), (2 * i + 1 < kC) ? ex2_approx(ta.y) : 0.f);
        const float2 eb = make_float2(ex2_approx(tb.x), (2 * i + 1 < kC) ? ex2_approx(tb.y) : 0.f);
        sa2 = add2(sa2, ea);
        sb2 = add2(sb2, eb);
        ta2 = fma2(ea, ta, ta2);  // sum e'_c t_c
        tb2 = fma2(eb, tb, tb2);
        if (JS) {
          xab2 = fma2(ea, tb, xab2);  // sum e'^a_c t^b_c
          xba2 = fma2(eb, ta, xba2);
        }
        a[i] = ea;
        b[ENS ? i : 0] = eb;
      }
      sa = sa2.x + sa2.y;
      sb = sb2.x + sb2.y;
      const float tsa = ta2.x + ta2.y, tsb = tb2.x + tb2.y;
      // one reciprocal for both member sums: 1/Sa = Sb r, 1/Sb = Sa r with r = rcp(Sa Sb) (tying 1/Sz in as
      // well was measured slower: it makes the log phase wait for the fused-logit phase)
      const float sab = sa * sb;
      const float rab = rcp_approx(sab);
      const float ra = sb * rab, rb = sa * rab;
      // mean probabilities m_c = (e^a_c/Sa + e^b_c/Sb)/2 = ka * u_c with u_c = e^a_c + rho e^b_c
      const float kas = 0.5f * ra, kbs = 0.5f * rb;
      const float2 ka = splat(kas), rho = splat(sa * rb);
      float2 hm2 = splat(0.f);
      float mlm = 0.f, umax = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 uu = fma2(b[ENS ? i : 0], rho, a[i]);
        const float2 me = fma2(uu, ka, eps);
        hm2 = fma2(uu, make_float2(lg2_approx(me.x), (2 * i + 1 < kC) ? lg2_approx(me.y) : 0.f), hm2);
        if (JS) {
          const float2 m = mul2(uu, ka);
          mlm += m.x > 0.f ? m.x * lg2_approx(m.x) : 0.f;
          if (2 * i + 1 < kC) mlm += m.y > 0.f ? m.y * lg2_approx(m.y) : 0.f;
        }
        a[i] = uu;  // keep the (unnormalised) mean probabilities for the arg-max below
        umax = fmaxf(umax, fmaxf(uu.x, uu.y));
      }
      if (NB > 0) {
        // arg-max of the mean probabilities (the AUROC's positive flag): classes within kMargBand of the maximum
        const unsigned cand = ~below_mask<NP>(a, umax * (1.f - kMargBand)) & kClassBits;
        marg = __clz(cand) - (32 - kC);
        tie_u = __popc(cand) != 1;
      }
      // entropies in bits: H2(p) = lg2 S' - (sum e'_c t_c)/S'; the reference's log(p + eps) adds
      // -C*eps nats (p >> eps)
      // only lg2 Sa' + lg2 Sb' = lg2(Sa' Sb') enters the mutual information: one logarithm
      const float lsab = lg2_approx(sab);
      const float ceps = (float)kC * kEps;
      // explicit roundings: every instantiation of this kernel (bins-only / generic, all strategies) must produce the
      // same MI bit pattern for a pixel, or the AUROC histograms of two code paths differ at bin edges
      const float hsum = __fmaf_rn(-tsb, rb, __fmaf_rn(-tsa, ra, lsab));
      mi = __fmaf_rn(kLn2, __fmaf_rn(-0.5f, hsum, __fmul_rn(-kas, __fadd_rn(hm2.x, hm2.y))), ceps);
      if (JS) {
        // sum_c m_c lg2 p_c = ka*Ta + kb*Xba - lg2 Sa' ; sum_c m_c lg2 q_c = kb*Tb + ka*Xab - lg2 Sb'
        const float xab = xab2.x + xab2.y, xba = xba2.x + xba2.y;
        const float lsa = lg2_approx(sa), lsb = lg2_approx(sb);
        const float mlp = kas * tsa + kbs * xba - lsa;
        const float mlq = kbs * tsb + kas * xab - lsb;
        js = kLn2 * (mlm - 0.5f * (mlp + mlq));
      }
    }
    };

    // ---- P0 (max-confidence only): the member with the larger max-softmax supplies the logits.  The members
    // phase runs first -- its softmax sums ARE the two confidences -- and the winner's 19 logits are then read
    // again from the ring unit, which is released only now (one set of member exponentials instead of two).
    float2 v[NP];
    float w0s = w0.x, w1s = w1.x;  // per-pixel fusion weights (the exact slow path re-fuses from global memory)
    if (MODE == 3) {
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        amax = fmaxf(amax, fmaxf(a[i].x, a[i].y));
        bmax = fmaxf(bmax, fmaxf(b[ENS ? i : 0].x, b[ENS ? i : 0].y));
      }
      members_phase();
      // confidences 1/S with the shifted-exponent correction (see header)
      const float ca0 = -(amax * kLog2e), cb0 = -(bmax * kLog2e);
      const float ra = rcp_approx(sa), rb = rcp_approx(sb);
      const float ca = fmaf(ra, fmaf(amax, kLog2e, ca0) * kLn2, ra), cb = fmaf(rb, fmaf(bmax, kLog2e, cb0) * kLn2, rb);
      const bool pick_a = ca > cb;
      n_pick += act && fabsf(ca - cb) <= 4.8e-7f * fmaxf(ca, cb);
      w0s = pick_a ? 1.f : 0.f;
      w1s = pick_a ? 0.f : 1.f;
      const uint32_t sv = held_unit + (pick_a ? 0u : (uint32_t)(kMO * 4));
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        v[i].x = lds_f32(sv + (2 * i) * kPS * 4);
        v[i].y = (2 * i + 1 < kC) ? lds_f32(sv + (2 * i + 1) * kPS * 4) : kDummy;
      }
      if (FAST == 0 && p.fused != nullptr) {
        // the fused-logit MAP is the reference's expression as written, mask*l1 + (1-mask)*l2: the other member
        // enters as 0 * x (a signed zero, or NaN for an infinite logit)
        const uint32_t so = held_unit + (pick_a ? (uint32_t)(kMO * 4) : 0u);
        const float2 zero = splat(0.f);
#pragma unroll
        for (int i = 0; i < NP; ++i) {
          float2 other;
          other.x = lds_f32(so + (2 * i) * kPS * 4);
          other.y = (2 * i + 1 < kC) ? lds_f32(so + (2 * i + 1) * kPS * 4) : 0.f;
          v[i] = add2(v[i], fma2(zero, other, nz));
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive_a(held_bar);
    }

    // ---- P1: fused logits (exact), maxima, first arg-max
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      if (MODE == 3)
        break;  // v already holds the winning member's logits
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
    float vmax = v[0].x;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
      vmax = fmaxf(vmax, fmaxf(v[i].x, v[i].y));
      if (ENS && MODE != 3) {
        amax = fmaxf(amax, fmaxf(a[i].x, a[i].y));
        bmax = fmaxf(bmax, fmaxf(b[ENS ? i : 0].x, b[ENS ? i : 0].y));
      }
    }
    // Arg-max of the fused logits.  Classes within the band of the maximum are candidates: a pending division by
    // T > 0 (div_mode 1) can merge the maximum only with values a few ulp below it (torch's argmax over the divided
    // logits then returns the first of them), and the ECE's arg-max over fp32 probabilities can tie only within
    // 2.5e-7 of the divided maximum; a single candidate is the arg-max of both, anything else is resolved exactly.
    const unsigned cand_v = ~below_mask<NP>(v, vmax - fmaf(fabsf(vmax), kPredBandRel, p.band_abs)) & kClassBits;
    const int arg = __clz(cand_v) - (32 - kC);
    const bool tie_v = __popc(cand_v) != 1;

    // optional fused-logit output (bit exact: div_mode is 0 or 2 whenever it is requested)
    if (FAST == 0 && p.fused != nullptr && act) {
      float* fo = p.fused + ((long long)img * kC * HW + p0 + t);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        fo[(2 * i) * HW] = v[i].x;
        if (2 * i + 1 < kC) fo[(2 * i + 1) * HW] = v[i].y;
      }
    }

    // ---- P3: softmax denominator of the fused logits (shifted exponents, see header)
    float sz, zdelta;
    {
      const float cz = -(vmax * p.kz);
      zdelta = fmaf(vmax, p.kz, cz) * kLn2;  // exact residual of the rounded product, in nats
      float2 sz2 = splat(0.f);
      const float2 cz2 = splat(cz);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 x = fma2(v[i], kz, cz2);
        sz2 = add2(sz2, make_float2(ex2_approx(x.x), (2 * i + 1 < kC) ? ex2_approx(x.y) : 0.f));
      }
      sz = sz2.x + sz2.y;
    }

    if (MODE != 3) members_phase();

    // ---- confidence, ECE bin, slow paths
    int pred = arg, epred = arg, bin, ambig = 0;
    float conf;
    {
      // one finiteness test: every term is >= 1 or tiny, so the sum is non-finite iff a term is
      const float chk = ENS ? (sz + sa) + (sb + mi) : sz;
      const bool sane = fabsf(chk) < 3e38f;
      const float r0 = rcp_approx(sz);
      const float r = fmaf(r0, fmaf(-sz, r0, 1.f), r0);  // one Newton step: <= 1 ulp
      conf = fminf(fmaf(r, zdelta, r), 1.f);
      // relative half-width of the "near an edge" band: 16 ulp for this kernel's own error, plus -- when the
      // reference divides the logits by T in fp32 before its softmax (div_mode 1: this kernel keeps the exact
      // quotient in the exponent instead) -- the reference's own rounding of z = v/T, ~ulp(zmax) relative
      const float band = div_mode == 1 ? fmaf(fabsf(vmax * p.rT), 1.5e-7f, 2.4e-6f) : 2.4e-6f;
      bool near;
      if (FAST != 0) {
        // edges are linspace(0,1,nb+1) (checked by the host): bin = ceil(conf*nb) - 1 unless conf is
        // within ~18 ulp of an edge, and those pixels take the fp64 path against the real edges
        const float s = conf * p.nbf;
        const float tt = s + 12582912.f;  // 1.5 * 2^23: round to nearest integer
        const int j = __float_as_int(tt) - 0x4b400000;
        const float d = s - (tt - 12582912.f);
        bin = min(j - (d < 0.f ? 1 : 0), nb - 1);
        near = fabsf(d) <= s * band && j < nb;
      } else {
        bin = ece_bin_fast(conf, s_edges, nb);
        near = false;
        if (bin >= 0) {
          const float tol = conf * band;
          near = (bin > 0 && (conf - s_edges[bin]) <= tol) || (bin < nb - 1 && (s_edges[bin + 1] - conf) <= tol);
        }
      }
      const bool tie = tie_v || (ENS && tie_u);
      if (act && (!sane || near || tie)) {
        const long long go = (long long)img * kC * HW + p0 + t;
        const float* ga = p.a + go;
        const float* gb = ENS ? p.b + go : nullptr;
        int eamb = 0, mamb = 0;
        if (!sane) {
          PixOut so;
          slow_pixel<ENS, JS>(p, s_edges, ga, gb, lab, so);
          pred = so.pred;
          epred = so.epred;
          conf = so.conf;
          bin = so.bin;
          ambig = so.ambig;
          mi = so.mi;
          js = so.js;
          marg = so.mpred;
          eamb = so.eamb;
          mamb = so.mamb;
        } else {
          if (near) {
            int amb = 0;
            conf = exact_confidence(ga, gb, HW, kC, MODE == 2, w0s, w1s, div_mode, T, s_edges, nb, &amb);
            ambig = amb;
            bin = ece_bin(conf, s_edges, nb);
          }
          if (tie) {
            TieOut to;
            resolve_ties(ga, gb, HW, kC, MODE == 2, w0s, w1s, div_mode, T, lab, tie_v, ENS && tie_u, to);
            if (tie_v) {
              pred = to.pred;
              epred = to.epred;
              eamb = to.eamb;
            }
            if (ENS && tie_u) {
              marg = to.marg;
              mamb = to.mamb;
            }
          }
        }
        // bookkeeping that only these rare pixels can need: the ambiguity counts, and AWX_CNT_CORRECT (arg-max of
        // the LOGITS == label), which every other pixel contributes through its ECE bin's packed `correct` word
        // (arg-max of the PROBABILITIES == label): add the difference, or the whole term for a pixel in no bin
        if (have_labels && lab != ignore) {
          n_ambig += ambig;
          if (eamb) atomicAdd(&s_cnt[AWX_CNT_EPRED_AMBIG], 1u);
          if (mamb && NB > 0) atomicAdd(&s_cnt[AWX_CNT_MARG_AMBIG], 1u);
          n_correct += (unsigned)((int)(lab == pred) - (bin >= 0 ? (int)(lab == epred) : 0));
        }
      }
    }

    if (FAST == 0 && act) {
      if (p.pred) {
        if (p.pred_dtype == AWX_PRED_U8)
          static_cast<uint8_t*>(p.pred)[li] = (uint8_t)pred;
        else
          static_cast<long long*>(p.pred)[li] = pred;
      }
      if (p.conf) p.conf[li] = conf;
      if (ENS && p.mi) p.mi[li] = mi;
      if (ENS && JS && p.js) p.js[li] = js;
    }

    // ---- statistics
    if (have_labels) {
      if (since_flush + 32u > kFlushPixels) {  // warp-uniform, rare: move the 32-bit sums to the global bins
        __syncwarp();
        for (int i = lane; i < nb; i += 32) {
          u64 sum = 0;
          for (int r = 0; r < kEceRep; ++r) {
            sum += ((u64)my_hi[i * kEceRep + r] << 16) + my_lo[i * kEceRep + r];
            my_hi[i * kEceRep + r] = my_lo[i * kEceRep + r] = 0u;
          }
          if (sum) {
            atomicAdd(p.bins + p.lay.ece_conf_hi + i, sum >> 32);
            atomicAdd(p.bins + p.lay.ece_conf_lo + i, sum & 0xffffffffull);
          }
        }
        __syncwarp();
        since_flush = 0;
      }
      since_flush += 32u;
      const bool valid = act && lab != ignore;
      const bool correct = valid && lab == epred;  // the ECE's accuracy term (arg-max of the probabilities)
      int ckey = -1, akey = -1;
      if (valid) {
        // confusion index as torch evaluates targets*C + predictions (uint8 product wraps mod 256)
        const int idx = (lab_u8 ? ((lab * kC) & 0xff) : lab * kC) + pred;
        const bool inside = lab_u8 ? (idx < kC * kC) : (lab >= 0 && lab < kC);
        if (inside)
          ckey = idx;
        else
          ++n_bad;
        if (ENS && NB > 0) {
          // floor(mi * scale) clamped to [0, NB-1]; NaN -> 0 (fmaxf returns the non-NaN operand)
          const float qv = fminf(fmaxf(mi * p.auroc_scale, 0.f), p.auroc_top);
          // floor of 0 <= qv < 2^23 without the conversion pipe: round-down add of 2^23, low mantissa bits
          akey = (lab != marg ? 0 : NB) + (__float_as_int(__fadd_rd(qv, 8388608.f)) & 0x7fffff);
        }
      }
      if (!valid) bin = -1;
      // conf * 2^31 as an integer without the conversion pipe: 1/19 <= conf <= 1, so the product is >= 2^23 and
      // integer valued: its 24-bit significand shifted left by (exponent - 23)
      const unsigned cb = __float_as_uint(conf);
      const unsigned fx = bin >= 0 ? (((cb & 0x7fffffu) | 0x800000u) << ((cb >> 23) - 119u)) : 0u;
      // one warp-uniformity test for all three histograms (piecewise-constant real data)
      const int key = (ckey + 1) | ((akey + 1) << 9) | ((bin + 1) << 23);
      int same;
      __match_all_sync(0xffffffffu, key, &same);
      if (same) {
        if (key != 0) {
          const unsigned lo = __reduce_add_sync(0xffffffffu, fx & 0xffffu);
          const unsigned hi = __reduce_add_sync(0xffffffffu, fx >> 16);
          if (lane == 0) {
            if (ckey >= 0) red_add(conf_sa + 4u * ckey, 32u);
            if (akey >= 0) red_add(auroc_sa + 4u * akey, 32u);
            if (bin >= 0) {
              red_add(cnt_sa + 4u * bin, 32u);
              if (correct) red_add(cor_sa + 4u * bin, 32u);
              red_add(lo_sa + 4u * kEceRep * bin, lo);
              red_add(hi_sa + 4u * kEceRep * bin, hi);
            }
          }
        }
      } else if (ckey >= 0 && bin >= 0 && (akey >= 0 || !(ENS && NB > 0))) {
        // the common case under ONE branch: five reductions, no reconvergence point between them
        red_add(conf_sa + 4u * ckey, 1u);
        if (ENS && NB > 0) red_add(auroc_sa + 4u * akey, 1u);
        red_add(cnt_sa + 4u * bin, 1u);
        red_add_if(correct ? 0 : -1, cor_sa + 4u * bin, 1u);
        red_add(lo_sa + 4u * kEceRep * bin, fx & 0xffffu);
        red_add(hi_sa + 4u * kEceRep * bin, fx >> 16);
      } else {
        red_add_if(ckey, conf_sa + 4u * ckey, 1u);
        red_add_if(akey, auroc_sa + 4u * akey, 1u);
        red_add_if(bin, cnt_sa + 4u * bin, 1u);
        red_add_if(correct ? bin : -1, cor_sa + 4u * bin, 1u);
        red_add_if(bin, lo_sa + 4u * kEceRep * bin, fx & 0xffffu);
        red_add_if(bin, hi_sa + 4u * kEceRep * bin, fx >> 16);
      }
    }
    tin += step_tin;
    img += step_img;
    if (tin >= tpiu) {
      tin -= tpiu;
      ++img;
    }
  }

  if (!have_labels) return;
  {
    unsigned vv[4] = {n_correct, n_bad, n_ambig, n_pick};
    const int slot[4] = {AWX_CNT_CORRECT, AWX_CNT_BAD_LABEL, AWX_CNT_ECE_AMBIG, AWX_CNT_PICK_AMBIG};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned s = __reduce_add_sync(0xffffffffu, vv[k]);
      if (lane == 0 && s) atomicAdd(&s_cnt[slot[k]], s);
    }
  }
  // consumer-only barrier (the producer warp has already left)
  asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory");
  unsigned long long* bins = p.bins;
  // the remaining counters follow from the histograms: valid = sum(confusion) + bad,
  // ensemble-wrong = sum(auroc_pos), no-bin = valid - sum(ece_count); pixels are added by the host
  unsigned part_conf = 0, part_pos = 0, part_ece = 0, part_cor = 0;
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
      cnt += w_cnt[w * nb + i];
      cor += w_cor[w * nb + i];
      for (int r = 0; r < kEceRep; ++r)
        sum += ((u64)w_hi[(w * nb + i) * kEceRep + r] << 16) + w_lo[(w * nb + i) * kEceRep + r];
    }
    part_ece += (unsigned)cnt;
    part_cor += (unsigned)cor;
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
    const unsigned sr = __reduce_add_sync(0xffffffffu, part_cor);
    if (lane == 0) {
      if (sr) atomicAdd(&s_cnt[AWX_CNT_CORRECT], sr);  // correct pixels counted through their ECE bins
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
    if (s_cnt[AWX_CNT_PICK_AMBIG]) atomicAdd(bins + p.lay.counters + AWX_CNT_PICK_AMBIG, (u64)s_cnt[AWX_CNT_PICK_AMBIG]);
    if (s_cnt[AWX_CNT_MARG_AMBIG]) atomicAdd(bins + p.lay.counters + AWX_CNT_MARG_AMBIG, (u64)s_cnt[AWX_CNT_MARG_AMBIG]);
    if (s_cnt[AWX_CNT_EPRED_AMBIG]) atomicAdd(bins + p.lay.counters + AWX_CNT_EPRED_AMBIG, (u64)s_cnt[AWX_CNT_EPRED_AMBIG]);
    if (nobin) atomicAdd(bins + p.lay.counters + AWX_CNT_NO_BIN, (u64)nobin);
    if (blockIdx.x == 0) atomicAdd(bins + p.lay.counters + AWX_CNT_PIXELS, (u64)(p.B * p.HW));
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time dependency on libcuda)
PFN_cuTensorMapEncodeTiled_v12000 tensor_map_encoder() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(f);
  }();
  return fn;
}

// logits [B, 19, HW] seen as a 2-D tensor [B * 19 planes][HW pixels]; box = box_w pixels x 19 planes
int make_plane_map(CUtensorMap* tm, const float* base, long long planes, long long HW, int box_w) {
  auto enc = tensor_map_encoder();
  AWX_REQUIRE(enc != nullptr, AWX_E_UNSUPPORTED, "awx_score v2: cuTensorMapEncodeTiled is not available in this driver");
  const cuuint64_t gdim[2] = {(cuuint64_t)HW, (cuuint64_t)planes};
  const cuuint64_t gstride[1] = {(cuuint64_t)HW * 4};
  const cuuint32_t box[2] = {(cuuint32_t)box_w, (cuuint32_t)kC};
  const cuuint32_t estr[2] = {1, 1};
  static const CUtensorMapL2promotion promo = [] {   // dev knob AWX_V2_L2PROMO=0|1|2|3: none / 64 / 128 / 256 bytes
    const char* e = getenv("AWX_V2_L2PROMO");
    const int v = e ? atoi(e) : 0;
    return v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
         : v == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE;
  }();
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstride, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AWX_REQUIRE(r == CUDA_SUCCESS, AWX_E_UNSUPPORTED, "awx_score v2: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return AWX_OK;
}

template <int MODE, bool JS, int FAST, int DIV, int CW>
int launch_v2(const ScoreParams& p, cudaStream_t stream) {
  using G = Geo<CW>;
  auto kern = score_v2_kernel<MODE, JS, FAST, DIV, CW>;
  int dev = 0, max_smem = 0;
  AWX_CUDA(cudaGetDevice(&dev));
  AWX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = v2_ring_offset(CW, p.nb, p.auroc_bins);
  const size_t unit_bytes = (size_t)(MODE != 0 ? 2 : 1) * G::kUnitBytes;
  int nu = (int)(((size_t)max_smem - fixed) / unit_bytes);
  if (nu > kMaxUnits / (MODE != 0 ? 2 : 1)) nu = kMaxUnits / (MODE != 0 ? 2 : 1);
  if (const char* e = getenv("AWX_V2_UNITS")) {  // dev knob: ring depth sensitivity
    const int want = atoi(e);
    if (want >= 2 && want < nu) nu = want;
  }
  AWX_REQUIRE(nu >= 2, AWX_E_UNSUPPORTED, "awx_score v2: histograms leave no room for the TMA ring");
  const size_t smem = (size_t)nu * unit_bytes + fixed;
  AWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const long long ntiles = p.B * ((p.HW + G::kTP - 1) / G::kTP);
  long long blocks = sm_count();
  if (blocks > ntiles) blocks = ntiles;
  CUtensorMap tma, tmb;
  memset(&tma, 0, sizeof(tma));
  memset(&tmb, 0, sizeof(tmb));
  if (G::kTensorMap) {
    int rc = make_plane_map(&tma, p.a, p.B * kC, p.HW, G::kBoxW);
    if (rc != AWX_OK) return rc;
    if (MODE != 0) {
      rc = make_plane_map(&tmb, p.b, p.B * kC, p.HW, G::kBoxW);
      if (rc != AWX_OK) return rc;
    }
  }
  kern<<<(unsigned)blocks, G::kThreads, smem, stream>>>(p, nu, -0.0f, tma, tmb);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

// bins-only kernels: compile-time division mode; the single-member kernels (fewer live registers) run
// 19 consumer warps.  AWX_V2_WARPS=15|19 overrides for A/B measurements.
template <int MODE, int FAST>
int launch_v2_fast(const ScoreParams& p, cudaStream_t stream) {
  static const int cw = [] {
    const char* e = getenv("AWX_V2_WARPS");
    return (e && atoi(e) == 15) ? 15 : (e && atoi(e) == 19) ? 19 : kSingleWarps;
  }();
  if (p.div_mode == 2) return launch_v2<MODE, false, 0, -1, 15>(p, stream);  // exact division everywhere (T <= 0)
  if constexpr (MODE == 0) {
    if (cw == 19)
      return p.div_mode == 0 ? launch_v2<MODE, false, FAST, 0, 19>(p, stream) : launch_v2<MODE, false, FAST, 1, 19>(p, stream);
  }
  return p.div_mode == 0 ? launch_v2<MODE, false, FAST, 0, 15>(p, stream) : launch_v2<MODE, false, FAST, 1, 15>(p, stream);
}

// edges within 2 ulp of j/nb (what torch.linspace(0, 1, nb + 1) produces): the bins-only kernels
// classify without looking the edges up
bool uniform_edges(const ScoreParams& p) {
  if (p.edges[0] != 0.f || p.edges[p.nb] != 1.f) return false;
  for (int j = 1; j < p.nb; ++j) {
    const double want = (double)j / (double)p.nb;
    if (fabs((double)p.edges[j] - want) > 2.4e-7 * want) return false;
  }
  return true;
}

template <int MODE>
int launch_v2_mode(const ScoreParams& p, bool js, cudaStream_t stream) {
  const bool maps = p.pred || p.fused || p.conf || p.mi || p.js;
  if (!maps && p.labels != nullptr && !js && !p.debug_skip && uniform_edges(p) && p.nb == kFastEceBins &&
      p.auroc_bins == (MODE != 0 ? kFastAurocBins : 0))
    return p.label_mode == AWX_LABEL_U8 ? launch_v2_fast<MODE, 1>(p, stream) : launch_v2_fast<MODE, 2>(p, stream);
  if (MODE != 0 && js) return launch_v2<MODE, (MODE != 0), 0, -1, 15>(p, stream);
  return launch_v2<MODE, false, 0, -1, 15>(p, stream);
}

}  // namespace

bool score_v2_supported(const ScoreParams& p) {
  if (p.C != kC) return false;
  if (p.HW % 4 != 0) return false;
  if (p.B * p.HW >= (1LL << 32)) return false;  // CTA-level counters are 32 bit
  if (((uintptr_t)p.a & 15) != 0 || ((uintptr_t)p.b & 15) != 0) return false;
  if (p.labels && p.label_mode == AWX_LABEL_I64 && ((uintptr_t)p.labels & 7) != 0) return false;
  // two ring units next to the histograms (large ECE x AUROC configurations go to the register-resident kernel)
  const size_t ring = 2 * (size_t)(p.b ? 2 : 1) * Geo<15>::kUnitBytes;
  if (v2_ring_offset(15, p.nb, p.auroc_bins) + ring > (size_t)227 * 1024) return false;
  return true;
}

int launch_score_v2(const ScoreParams& p, bool ens, bool js, cudaStream_t stream) {
  if (!ens) return launch_v2_mode<0>(p, false, stream);
  if (p.strategy == AWX_FUSE_MEAN) return launch_v2_mode<2>(p, js, stream);
  if (p.strategy == AWX_FUSE_MAXCONF) return launch_v2_mode<3>(p, js, stream);
  return launch_v2_mode<1>(p, js, stream);
}

}  // namespace awx
