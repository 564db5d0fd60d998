// Shared device code of the two awx_score kernels (score.cu: register-resident generic kernel;
// score_v2.cu: TMA-staged, packed-math kernel for C == 19).
#pragma once
#include "awx_internal.cuh"

namespace awx {
namespace score_detail {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr float kEps = 1e-8f;  // the reference's log(p + 1e-8), metrics.py:360-363

struct ScoreParams {
  const float* a;
  const float* b;
  const void* labels;
  long long B, HW;
  int C;
  int strategy;
  float w0, w1, T;
  int div_mode;  // 0: none, 1: T>0 (division only where it can change the argmax), 2: exact everywhere
  float kz;      // log2(e)/T for div_mode 1, log2(e) otherwise
  float rT;      // fl(1/T), correctly rounded by the host (branch-free exact division, score_v2.cu)
  int label_mode;
  int ignore_index;
  int nb;
  float nbf;     // (float)nb
  int auroc_bins;
  float auroc_scale;
  float auroc_top;  // (float)(auroc_bins - 1)
  unsigned long long* bins;
  AwxBinsLayout lay;
  void* pred;
  int pred_dtype;
  float* fused;
  float* conf;
  float* mi;
  float* js;
  float edges[AWX_MAX_ECE_BINS + 1];
  int debug_skip;  // AWX_DEBUG_SKIP_MATH=1: consumers only drain the ring (data-movement ceiling; dev only)
};

struct PixOut {
  int pred;
  float conf;
  int bin;  // -1: in no bin
  int ambig;
  float mi;
  int mpred;
  float js;
};

__device__ __forceinline__ float fuse_one(float x, float y, bool mean, float w0, float w1) {
  // weighted / max-confidence: w0*x + w1*y as three roundings; mean: (x+y)/2 (model.py:445-458)
  return mean ? __fmul_rn(__fadd_rn(x, y), 0.5f) : __fadd_rn(__fmul_rn(w0, x), __fmul_rn(w1, y));
}

__device__ __forceinline__ int ece_bin(float conf, const float* e, int nb) {
  int b = (int)ceilf(conf * (float)nb) - 1;
  b = min(max(b, 0), nb - 1);
  while (b > 0 && !(conf > e[b])) --b;
  while (b < nb - 1 && conf > e[b + 1]) ++b;
  return (conf > e[b] && conf <= e[b + 1]) ? b : -1;
}

// Rare path: confidence of the fused logits in fp64, from global memory, with the exact
// fp32 fusion / division / subtraction the reference performs before its exp.
static __device__ __noinline__ float exact_confidence(const float* ga, const float* gb, long long HW, int C, bool mean,
                                               float w0, float w1, int div_mode, float T, const float* edges, int nb,
                                               int* ambig) {
  float zmax = 0.f;
  for (int c = 0; c < C; ++c) {
    float z = gb ? fuse_one(ga[c * HW], gb[c * HW], mean, w0, w1) : ga[c * HW];
    if (div_mode) z = __fdiv_rn(z, T);
    if (c == 0 || beats(z, zmax)) zmax = z;
  }
  double s = 0.0;
  for (int c = 0; c < C; ++c) {
    float z = gb ? fuse_one(ga[c * HW], gb[c * HW], mean, w0, w1) : ga[c * HW];
    if (div_mode) z = __fdiv_rn(z, T);
    s += exp((double)__fsub_rn(z, zmax));
  }
  const double cd = 1.0 / s;
  const float cf = (float)cd;
  const double ulp = (double)(__int_as_float(__float_as_int(cf) + 1) - cf);
  for (int j = 1; j < nb; ++j)
    if (fabs(cd - (double)edges[j]) <= 3.0 * ulp) *ambig = 1;
  return cf;
}

template <int CS, bool ENS, bool JS>
__device__ __forceinline__ void score_pixel(float (&a)[CS > 0 ? CS : AWX_MAX_CLASSES],
                                            float (&b)[ENS ? (CS > 0 ? CS : AWX_MAX_CLASSES) : 1], const int C,
                                            const ScoreParams& p, const float* s_edges, const float* ga,
                                            const float* gb, float w0, float w1, float amax, float bmax, PixOut& o) {
  const bool mean = ENS && p.strategy == AWX_FUSE_MEAN;
  const float T = p.T;
  // ---- pass 1: max / argmax of the fused logits (first index wins ties, NaN wins)
  float vmax = 0.f;
  int arg = 0;
  if (p.div_mode == 1) {
    float second = -INFINITY;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float v = ENS ? fuse_one(a[c], b[ENS ? c : 0], mean, w0, w1) : a[c];
      if (c == 0) {
        vmax = v;
      } else if (beats(v, vmax)) {
        second = vmax;
        vmax = v;
        arg = c;
      } else {
        second = fmaxf(second, v);
      }
    }
    // division by T>0 is monotone, but rounding can merge vmax with an earlier, slightly
    // smaller value; torch's argmax over the divided logits would then return that index.
    const float tol = fmaxf(fabsf(vmax) * 4.8e-7f, 1e-30f);
    if (second >= vmax - tol) {
      const float zmax = __fdiv_rn(vmax, T);
      for (int c = 0; c < arg; ++c) {
        const float v = ga ? (gb ? fuse_one(ga[c * p.HW], gb[c * p.HW], mean, w0, w1) : ga[c * p.HW]) : vmax;
        if (__fdiv_rn(v, T) == zmax) {
          arg = c;
          break;
        }
      }
    }
  } else {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float v = ENS ? fuse_one(a[c], b[ENS ? c : 0], mean, w0, w1) : a[c];
      if (p.div_mode == 2) v = __fdiv_rn(v, T);
      if (c == 0 || beats(v, vmax)) {
        vmax = v;
        arg = c;
      }
    }
  }
  o.pred = arg;

  // ---- pass 2a: softmax denominator of the fused logits -> confidence -> ECE bin
  float sz = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = ENS ? fuse_one(a[c], b[ENS ? c : 0], mean, w0, w1) : a[c];
    if (p.div_mode == 2) v = __fdiv_rn(v, T);
    sz += ex2_approx((v - vmax) * p.kz);
  }
  float conf = __frcp_rn(sz);
  int bin = ece_bin(conf, s_edges, p.nb);
  o.ambig = 0;
  if (bin >= 0) {
    // 16 ulp, plus the reference's rounding of z = v/T when it divides before the softmax (see score_v2.cu)
    const float tol = conf * (p.div_mode == 1 ? fmaf(fabsf(vmax / T), 1.5e-7f, 2.4e-6f) : 1.9e-6f);
    const bool near_lo = bin > 0 && (conf - s_edges[bin]) <= tol;
    const bool near_hi = bin < p.nb - 1 && (s_edges[bin + 1] - conf) <= tol;
    if (near_lo || near_hi) {
      conf = exact_confidence(ga, gb, p.HW, C, mean, w0, w1, p.div_mode, T, s_edges, p.nb, &o.ambig);
      bin = ece_bin(conf, s_edges, p.nb);
    }
  }
  o.conf = conf;
  o.bin = bin;

  // ---- members: softmax sums, entropies, mean-probability argmax, MI (and reverse-KL "JS")
  if (ENS) {
    float sa = 0.f, sb = 0.f, ta = 0.f, tb = 0.f, xab = 0.f, xba = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float da = a[c] - amax;
      const float db = b[ENS ? c : 0] - bmax;
      const float ea = ex2_approx(da * kLog2e);
      const float eb = ex2_approx(db * kLog2e);
      sa += ea;
      sb += eb;
      // guard 0 * -inf for fully underflowed classes
      ta = fmaf(ea, ea > 0.f ? da : 0.f, ta);
      tb = fmaf(eb, eb > 0.f ? db : 0.f, tb);
      if (JS) {
        xab = fmaf(ea, db, xab);  // sum_c e^a_c * d^b_c
        xba = fmaf(eb, da, xba);
      }
      a[c] = ea;
      b[ENS ? c : 0] = eb;
    }
    const float ra = __frcp_rn(sa), rb = __frcp_rn(sb);
    const float ka = 0.5f * ra, kb = 0.5f * rb;
    float hm2 = 0.f, mlm2 = 0.f, mbest = 0.f;
    int marg = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float m = fmaf(a[c], ka, b[ENS ? c : 0] * kb);
      hm2 = fmaf(m, lg2_approx(m + kEps), hm2);
      if (JS) mlm2 += (m > 0.f) ? m * lg2_approx(m) : 0.f;
      if (c == 0 || beats(m, mbest)) {
        mbest = m;
        marg = c;
      }
    }
    const float lsa = kLn2 * lg2_approx(sa), lsb = kLn2 * lg2_approx(sb);
    // H(p) with the reference's eps: -sum p ln(p+eps) ~= ln S - T/S - C*eps (p >> eps)
    const float ceps = (float)C * kEps;
    const float ha = lsa - ta * ra - ceps;
    const float hb = lsb - tb * rb - ceps;
    o.mi = -kLn2 * hm2 - 0.5f * (ha + hb);
    o.mpred = marg;
    if (JS) {
      // sum_c m ln p = ka*Ta + kb*Xba - ln Sa ; sum_c m ln q = kb*Tb + ka*Xab - ln Sb
      const float mlp = ka * ta + kb * xba - lsa;
      const float mlq = kb * tb + ka * xab - lsb;
      o.js = kLn2 * mlm2 - 0.5f * (mlp + mlq);
    }
  }
}


}  // namespace score_detail
}  // namespace awx
