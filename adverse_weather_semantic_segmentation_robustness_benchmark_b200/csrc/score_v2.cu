// awx_score, kernel v2 (C == 19): TMA-staged, register-resident, packed-fp32 scoring.
//
// One persistent CTA per SM: warp 16 is the producer, warps 0-15 (512 threads) consume.
//   producer  walks the CTA's tiles (512 consecutive pixels of one image) and, per member, issues 19
//             bulk async copies (cp.async.bulk, one 2 KB plane segment each) into the next free unit of
//             a shared-memory ring, completing on that unit's mbarrier.  The ring holds 4-5 units of
//             38 KB, i.e. up to ~190 KB per SM in flight independent of what the consumers do.
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
constexpr int kTP = 512;                        // pixels per tile
constexpr int kCons = 512;                      // consumer threads (1 px each)
constexpr int kConsWarps = kCons / 32;
constexpr int kV2Threads = kCons + 32;          // + producer warp
constexpr int kUnitFloats = kC * kTP;
constexpr int kUnitBytes = kUnitFloats * 4;     // 38912
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
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
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

// MODE: 0 single member, 1 weighted average, 2 mean
//
// Consumer thread t owns pixel t of the 512-pixel tile.  Its 19 class values are held as 10 float2
// PAIRS OF CLASSES (2i, 2i+1); the 20th slot is a finite "never wins" dummy (-1e30) whose
// exponentials are exactly 0.  Element-wise work (fusion, x - max, scaling, e*d products) runs on
// FADD2/FMUL2/FFMA2 over class pairs; maxima and arg-maxima are scalar (no packed min/max exists).
template <int MODE, bool JS>
__global__ void __launch_bounds__(kV2Threads, 1) score_v2_kernel(const __grid_constant__ ScoreParams p, const int NU,
                                                                  const float negzero) {
  constexpr bool ENS = MODE != 0;
  constexpr int NP = (kC + 1) / 2;  // 10 class pairs
  constexpr float kDummy = -1e30f;
  const int nb = p.nb, NB = p.auroc_bins;
  const bool have_labels = p.labels != nullptr;
  extern __shared__ __align__(128) unsigned char smem[];
  float* units = reinterpret_cast<float*>(smem);
  unsigned char* q8 = smem + (size_t)NU * kUnitBytes;
  u64* full = reinterpret_cast<u64*>(q8);
  u64* empty = full + kMaxUnits;
  u64* w_cnt64 = empty + kMaxUnits;                     // [warps][nb]
  u64* w_cor64 = w_cnt64 + kConsWarps * nb;
  u64* w_sum64 = w_cor64 + kConsWarps * nb;
  unsigned* w_cc = reinterpret_cast<unsigned*>(w_sum64 + kConsWarps * nb);  // [warps][nb] count | correct << 16
  unsigned* w_lo = w_cc + kConsWarps * nb;
  unsigned* w_hi = w_lo + kConsWarps * nb;
  unsigned* s_conf = w_hi + kConsWarps * nb;            // [C*C]
  unsigned* s_auroc = s_conf + kC * kC;                 // [2*NB]
  unsigned* s_cnt = s_auroc + 2 * NB;                   // [8]
  float* s_edges = reinterpret_cast<float*>(s_cnt + 8); // [nb+1]
  {
    unsigned* w = reinterpret_cast<unsigned*>(w_cnt64);
    const int words = kConsWarps * nb * 9 + kC * kC + 2 * NB + 8;
    for (int i = threadIdx.x; i < words; i += kV2Threads) w[i] = 0u;
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

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long HW = p.HW;
  const long long tpi = (HW + kTP - 1) / kTP;  // tiles per image
  const long long ntiles = p.B * tpi;

  if (warp == kConsWarps) {
    // ------------------------------------------------------------------ producer warp
    unsigned uc = 0;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long img = tile / tpi;
      const long long p0 = (tile - img * tpi) * kTP;
      const unsigned npx = (unsigned)((HW - p0) < kTP ? (HW - p0) : kTP);
#pragma unroll
      for (int m = 0; m < (ENS ? 2 : 1); ++m) {
        const unsigned u = uc % NU, ph = (uc / NU) & 1u;
        ++uc;
        mbar_wait(empty + u, ph ^ 1u);
        if (lane == 0) mbar_expect_tx(full + u, kC * npx * 4u);
        __syncwarp();
        if (lane < kC) {
          const float* src = (m == 0 ? p.a : p.b) + (img * kC + lane) * HW + p0;
          bulk_load(units + (size_t)u * kUnitFloats + lane * kTP, src, npx * 4u, full + u);
        }
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
  const bool lab_u8 = p.label_mode == AWX_LABEL_U8;
  unsigned n_valid = 0, n_correct = 0, n_bad = 0, n_ambig = 0, n_wrong = 0, n_nobin = 0, n_pix = 0;
  unsigned uc = 0, since_flush = 0;
  unsigned* my_cc = w_cc + warp * nb;
  unsigned* my_lo = w_lo + warp * nb;
  unsigned* my_hi = w_hi + warp * nb;
  const int t = threadIdx.x;

  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long img = tile / tpi;
    const long long p0 = (tile - img * tpi) * kTP;
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
      const unsigned u = uc % NU, ph = (uc / NU) & 1u;
      ++uc;
      mbar_wait(full + u, ph);
      const float* s = units + (size_t)u * kUnitFloats + t;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        a[i].x = act ? s[(2 * i) * kTP] : 0.f;
        a[i].y = (2 * i + 1 < kC) ? (act ? s[(2 * i + 1) * kTP] : 0.f) : kDummy;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + u);
    }
    if (ENS) {
      const unsigned u = uc % NU, ph = (uc / NU) & 1u;
      ++uc;
      mbar_wait(full + u, ph);
      const float* s = units + (size_t)u * kUnitFloats + t;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        b[ENS ? i : 0].x = act ? s[(2 * i) * kTP] : 0.f;
        b[ENS ? i : 0].y = (2 * i + 1 < kC) ? (act ? s[(2 * i + 1) * kTP] : 0.f) : kDummy;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty + u);
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
    if (p.div_mode == 2) {
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
    int arg = 0;
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {  // first index attaining the max
      if (2 * i + 1 < kC) arg = (v[i].y == vmax) ? 2 * i + 1 : arg;
      arg = (v[i].x == vmax) ? 2 * i : arg;
    }
    if (p.div_mode == 1) {
      // division by T > 0 is monotone but can merge the max with an earlier value within ~2 ulp
      float cur = v[0].x, second = -INFINITY;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        if (i > 0) {
          second = fmaxf(second, fminf(v[i].x, cur));
          cur = fmaxf(cur, v[i].x);
        }
        if (2 * i + 1 < kC) {
          second = fmaxf(second, fminf(v[i].y, cur));
          cur = fmaxf(cur, v[i].y);
        }
      }
      const float tol = fmaxf(fabsf(vmax) * 4.8e-7f, 1e-30f);
      if (act && second >= vmax - tol) {
        const float zmax = __fdiv_rn(vmax, T);
        const float* ga = p.a + img * kC * HW + p0 + t;
        const float* gb = ENS ? p.b + img * kC * HW + p0 + t : nullptr;
        for (int c = 0; c < arg; ++c) {
          const float x = ENS ? fuse_one(ga[c * HW], gb[c * HW], MODE == 2, p.w0, p.w1) : ga[c * HW];
          if (__fdiv_rn(x, T) == zmax) {
            arg = c;
            break;
          }
        }
      }
    }

    // optional fused-logit output (bit exact: div_mode is 0 or 2 whenever it is requested)
    if (p.fused != nullptr && act) {
      float* fo = p.fused + img * kC * HW + p0 + t;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        fo[(2 * i) * HW] = v[i].x;
        if (2 * i + 1 < kC) fo[(2 * i + 1) * HW] = v[i].y;
      }
    }

    // ---- P3: softmax denominator of the fused logits (dominant term exactly 1)
    float2 sz2 = splat(0.f);
    {
      const float2 vm2 = splat(vmax);
#pragma unroll
      for (int i = 0; i < NP; ++i) sz2 = add2(sz2, ex2_2(mul2(sub2(v[i], vm2), kz)));
    }
    const float sz = sz2.x + sz2.y;

    // ---- members: softmax sums, entropies, mean probabilities
    float mi = 0.f, js = 0.f, sa = 1.f, sb = 1.f;
    int marg = 0;
    if (ENS) {
      float2 sa2 = splat(0.f), sb2 = splat(0.f), ta2 = splat(0.f), tb2 = splat(0.f), xab2 = splat(0.f), xba2 = splat(0.f);
      const float2 am2 = splat(amax), bm2 = splat(bmax);
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 da = sub2(a[i], am2), db = sub2(b[ENS ? i : 0], bm2);
        const float2 ea = ex2_2(mul2(da, l2e)), eb = ex2_2(mul2(db, l2e));
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
      const float ra = __frcp_rn(sa), rb = __frcp_rn(sb);
      const float2 ka = splat(0.5f * ra), kb = splat(0.5f * rb);
      float2 hm2 = splat(0.f);
      float mlm = 0.f, mmax = 0.f;
#pragma unroll
      for (int i = 0; i < NP; ++i) {
        const float2 m = fma2(a[i], ka, mul2(b[ENS ? i : 0], kb));
        hm2 = fma2(m, lg2_2(add2(m, eps)), hm2);
        if (JS) {
          const float2 l = lg2_2(m);
          mlm += m.x > 0.f ? m.x * l.x : 0.f;
          mlm += m.y > 0.f ? m.y * l.y : 0.f;
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

    // ---- per-pixel epilogue
    PixOut o;
    o.pred = arg;
    o.mi = mi;
    o.js = js;
    o.mpred = marg;
    o.ambig = 0;
    {
      const bool sane = isfinite(sz) && (!ENS || (isfinite(sa) && isfinite(sb) && isfinite(mi)));
      float conf = __frcp_rn(sz);
      int bin = ece_bin(conf, s_edges, nb);
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
                                    p.div_mode, T, s_edges, nb, &o.ambig);
            bin = ece_bin(conf, s_edges, nb);
          }
        }
        o.conf = conf;
        o.bin = bin;
      }
    }

    if (act) {
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
      n_pix += act;
      n_valid += valid;
      n_correct += correct;
      int ckey = -1;
      if (valid) {
        // confusion index as torch evaluates targets*C + predictions (uint8 product wraps mod 256)
        const int idx = (lab_u8 ? ((lab * kC) & 0xff) : lab * kC) + o.pred;
        const bool inside = lab_u8 ? (idx < kC * kC) : (lab >= 0 && lab < kC);
        if (inside)
          ckey = idx;
        else
          ++n_bad;
        n_ambig += o.ambig;
        n_nobin += o.bin < 0;
      }
      hist_add(s_conf, ckey, lane);
      if (ENS && NB > 0) {
        int akey = -1;
        if (valid) {
          const bool wrong = lab != o.mpred;
          n_wrong += wrong;
          float qv = floorf(o.mi * p.auroc_scale);
          qv = is_nan(qv) ? 0.f : qv;
          akey = (wrong ? 0 : NB) + (int)fminf(fmaxf(qv, 0.f), (float)(NB - 1));
        }
        hist_add(s_auroc, akey, lane);
      }
      // ECE: warp-private packed words; uniform fast path via match_all
      const int bin = valid ? o.bin : -1;
      const unsigned fx = bin >= 0 ? __float2uint_rz(o.conf * 2147483648.f) : 0u;
      int same;
      __match_all_sync(0xffffffffu, bin, &same);
      if (same) {
        if (bin >= 0) {
          const unsigned ncor = __popc(__ballot_sync(0xffffffffu, correct));
          const unsigned lo = __reduce_add_sync(0xffffffffu, fx & 0xffffu);
          const unsigned hi = __reduce_add_sync(0xffffffffu, fx >> 16);
          if (lane == 0) {
            atomicAdd(my_cc + bin, 32u | (ncor << 16));
            atomicAdd(my_lo + bin, lo);
            atomicAdd(my_hi + bin, hi);
          }
        }
      } else if (bin >= 0) {
        atomicAdd(my_cc + bin, 1u | (correct ? 0x10000u : 0u));
        atomicAdd(my_lo + bin, fx & 0xffffu);
        atomicAdd(my_hi + bin, fx >> 16);
      }
    }
  }

  if (!have_labels) return;
  {
    unsigned vv[8] = {n_valid, n_correct, n_bad, n_ambig, n_wrong, 0u, n_nobin, n_pix};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const unsigned s = __reduce_add_sync(0xffffffffu, vv[k]);
      if (lane == 0 && s) atomicAdd(&s_cnt[k], s);
    }
  }
  // consumer-only barrier (the producer warp has already left)
  asm volatile("bar.sync 1, %0;" ::"n"(kCons) : "memory");
  unsigned long long* bins = p.bins;
  for (int i = threadIdx.x; i < kC * kC; i += kCons)
    if (s_conf[i]) atomicAdd(bins + p.lay.confusion + i, (u64)s_conf[i]);
  for (int i = threadIdx.x; i < 2 * NB; i += kCons)
    if (s_auroc[i]) atomicAdd(bins + (i < NB ? p.lay.auroc_pos + i : p.lay.auroc_neg + (i - NB)), (u64)s_auroc[i]);
  for (int i = threadIdx.x; i < nb; i += kCons) {
    u64 cnt = 0, cor = 0, sum = 0;
    for (int w = 0; w < kConsWarps; ++w) {
      const unsigned cc = w_cc[w * nb + i];
      cnt += w_cnt64[w * nb + i] + (cc & 0xffffu);
      cor += w_cor64[w * nb + i] + (cc >> 16);
      sum += w_sum64[w * nb + i] + ((u64)w_hi[w * nb + i] << 16) + w_lo[w * nb + i];
    }
    if (cnt) {
      atomicAdd(bins + p.lay.ece_count + i, cnt);
      if (cor) atomicAdd(bins + p.lay.ece_correct + i, cor);
      atomicAdd(bins + p.lay.ece_conf_hi + i, sum >> 32);
      atomicAdd(bins + p.lay.ece_conf_lo + i, sum & 0xffffffffull);
    }
  }
  if (threadIdx.x < 8 && s_cnt[threadIdx.x]) atomicAdd(bins + p.lay.counters + threadIdx.x, (u64)s_cnt[threadIdx.x]);
}

size_t v2_fixed_smem(int nb, int NB) {
  return 2 * kMaxUnits * sizeof(u64) + (size_t)kConsWarps * nb * 36 + ((size_t)kC * kC + 2 * (size_t)NB + 8) * 4 +
         (size_t)(nb + 1) * 4;
}

template <int MODE, bool JS>
int launch_v2(const ScoreParams& p, cudaStream_t stream) {
  auto kern = score_v2_kernel<MODE, JS>;
  int dev = 0, max_smem = 0;
  AWX_CUDA(cudaGetDevice(&dev));
  AWX_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t fixed = v2_fixed_smem(p.nb, p.auroc_bins);
  int nu = (int)(((size_t)max_smem - fixed - 256) / kUnitBytes);
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

}  // namespace

bool score_v2_supported(const ScoreParams& p) {
  if (p.C != kC || p.strategy == AWX_FUSE_MAXCONF) return false;
  if (p.HW % 4 != 0) return false;
  if (((uintptr_t)p.a & 15) != 0 || ((uintptr_t)p.b & 15) != 0) return false;
  if (p.labels && p.label_mode == AWX_LABEL_I64 && ((uintptr_t)p.labels & 7) != 0) return false;
  if (p.labels && p.label_mode == AWX_LABEL_U8 && ((uintptr_t)p.labels & 1) != 0) return false;
  if (p.fused && ((uintptr_t)p.fused & 7) != 0) return false;
  if ((p.conf && ((uintptr_t)p.conf & 7)) || (p.mi && ((uintptr_t)p.mi & 7)) || (p.js && ((uintptr_t)p.js & 7))) return false;
  if (p.pred && p.pred_dtype == AWX_PRED_I64 && ((uintptr_t)p.pred & 15) != 0) return false;
  if (p.pred && p.pred_dtype == AWX_PRED_U8 && ((uintptr_t)p.pred & 1) != 0) return false;
  return true;
}

int launch_score_v2(const ScoreParams& p, bool ens, bool js, cudaStream_t stream) {
  if (!ens) return launch_v2<0, false>(p, stream);
  if (p.strategy == AWX_FUSE_MEAN) return js ? launch_v2<2, true>(p, stream) : launch_v2<2, false>(p, stream);
  return js ? launch_v2<1, true>(p, stream) : launch_v2<1, false>(p, stream);
}

}  // namespace awx
