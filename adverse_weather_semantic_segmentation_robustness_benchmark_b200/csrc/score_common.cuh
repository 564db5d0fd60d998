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
  float rT;      // fl(1/T), correctly rounded by the host
  float band_abs;  // absolute half-width of the "top classes may tie" band on the UNdivided fused logits: 4e-7 * T in
                   // div_mode 1 (a gap of 2.5e-7 between the divided logits is where exp() rounds to 1 - 4 ulp), else 4e-7
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
  int pred;   // arg-max of the fused (divided) logits: confusion matrix, pixel accuracy, the prediction map
  int epred;  // arg-max of the fused fp32 probabilities: the ECE's accuracy term (metrics.py:161-162)
  float conf;
  int bin;  // -1: in no bin
  int ambig;
  float mi;
  int mpred;  // arg-max of the mean member probabilities (metrics.py:414-416)
  float js;
  int eamb, mamb;  // the label is one of several classes whose fp32 probabilities may tie for epred / mpred
};

// relative half-width of the band inside which the approximate mean probabilities of the hot loops (ex2.approx /
// rcp.approx: ~1e-6 relative) cannot order two classes: such pixels go to resolve_ties
constexpr float kMargBand = 4e-6f;
// |vmax| * this + band_abs: two fused logits this close can share a quotient by T, or have fp32 probabilities
// that tie for the maximum
constexpr float kPredBandRel = 4.8e-7f;

struct TieOut {
  int pred, epred, eamb, marg, mamb;
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

// Rare path (a few pixels per million; every pixel of a region with constant logits): the three arg-maxima of a
// pixel whose top classes the hot loop could not separate, from global memory.
//   pred   first index of the maximum of the fused logits exactly as the reference forms them (three roundings,
//          true division): torch.argmax of the logits (evaluate.py:179, metrics.py:50-51).
//   epred  the ECE takes torch.max over the fp32 softmax instead (metrics.py:161-162): classes whose divided
//          logits lie within 2.5e-7 of the maximum have exp(z - zmax) >= 1 - 4 ulp and may come out of torch's
//          fp32 exp / division EQUAL to the maximum probability, in which case the first of them wins.  Exactly
//          equal logits tie for certain (first index, no ambiguity).  Otherwise the probabilities are formed in
//          fp64, rounded to fp32, and the first class equal to the maximum is taken; torch's own result depends
//          on its vectorised exp, so when the label is one of the candidates the pixel is reported (eamb).
//   marg   arg-max of (softmax(a) + softmax(b)) / 2 (metrics.py:414-416): candidates are the classes within
//          8e-7 relative of the fp64 maximum; identical member logits tie for certain; otherwise the fp32 sum of
//          the fp32-rounded fp64 probabilities decides, reported (mamb) when the label is a candidate.
static __device__ __noinline__ void resolve_ties(const float* ga, const float* gb, long long HW, int C, bool mean, float w0,
                                                 float w1, int div_mode, float T, long long lab, bool want_pred,
                                                 bool want_marg, TieOut& o) {
  o.pred = o.epred = o.marg = 0;
  o.eamb = o.mamb = 0;
  if (want_pred) {
    float zmax = 0.f;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      float z = gb ? fuse_one(ga[c * HW], gb[c * HW], mean, w0, w1) : ga[c * HW];
      if (div_mode) z = __fdiv_rn(z, T);
      if (c == 0 || beats(z, zmax)) {
        zmax = z;
        arg = c;
      }
    }
    o.pred = o.epred = arg;
    if (!is_nan(zmax)) {
      double s = 0.0;
      int ncand = 0, nexact = 0, lab_in = 0;
      for (int c = 0; c < C; ++c) {
        float z = gb ? fuse_one(ga[c * HW], gb[c * HW], mean, w0, w1) : ga[c * HW];
        if (div_mode) z = __fdiv_rn(z, T);
        const float d = __fsub_rn(z, zmax);
        s += exp((double)d);
        if (d >= -2.5e-7f) {
          ++ncand;
          nexact += d == 0.f;
          lab_in |= (long long)c == lab;
        }
      }
      if (ncand > nexact) {
        const float pmax = (float)(1.0 / s);
        int first = -1;
        for (int c = 0; c < C && first < 0; ++c) {
          float z = gb ? fuse_one(ga[c * HW], gb[c * HW], mean, w0, w1) : ga[c * HW];
          if (div_mode) z = __fdiv_rn(z, T);
          const float d = __fsub_rn(z, zmax);
          if (d >= -2.5e-7f && (float)(exp((double)d) / s) == pmax) first = c;
        }
        o.epred = first < 0 ? arg : first;
        o.eamb = lab_in;
      }
    }
  }
  if (want_marg && gb) {
    float amax = ga[0], bmax = gb[0];
    for (int c = 1; c < C; ++c) {
      amax = fmaxf(amax, ga[c * HW]);
      bmax = fmaxf(bmax, gb[c * HW]);
    }
    double sa = 0.0, sb = 0.0;
    for (int c = 0; c < C; ++c) {
      sa += exp((double)__fsub_rn(ga[c * HW], amax));
      sb += exp((double)__fsub_rn(gb[c * HW], bmax));
    }
    double smax = -1.0;
    int arg = 0;
    for (int c = 0; c < C; ++c) {
      const double m = exp((double)__fsub_rn(ga[c * HW], amax)) / sa + exp((double)__fsub_rn(gb[c * HW], bmax)) / sb;
      if (m > smax) {
        smax = m;
        arg = c;
      }
    }
    o.marg = arg;
    if (smax == smax && smax > 0.0) {
      const double thr = smax * (1.0 - 8e-7);
      int ncand = 0, nsame = 0, lab_in = 0, first = -1;
      float a0 = 0.f, b0 = 0.f, mbest = -1.f;
      int best = arg;
      for (int c = 0; c < C; ++c) {
        const float av = ga[c * HW], bv = gb[c * HW];
        const double pa = exp((double)__fsub_rn(av, amax)) / sa, pb = exp((double)__fsub_rn(bv, bmax)) / sb;
        if (pa + pb >= thr) {
          if (first < 0) {
            first = c;
            a0 = av;
            b0 = bv;
          }
          ++ncand;
          nsame += (av == a0 && bv == b0);
          lab_in |= (long long)c == lab;
          const float m32 = __fadd_rn((float)pa, (float)pb);  // torch: stack(...).mean(0) = fl(p + q) / 2
          if (m32 > mbest) {
            mbest = m32;
            best = c;
          }
        }
      }
      if (ncand > 1) {
        o.marg = (nsame == ncand) ? first : best;
        o.mamb = (nsame == ncand) ? 0 : lab_in;
      }
    }
  }
}

template <int CS, bool ENS, bool JS>
__device__ __forceinline__ void score_pixel(float (&a)[CS > 0 ? CS : AWX_MAX_CLASSES],
                                            float (&b)[ENS ? (CS > 0 ? CS : AWX_MAX_CLASSES) : 1], const int C,
                                            const ScoreParams& p, const float* s_edges, const float* ga,
                                            const float* gb, float w0, float w1, float amax, float bmax, long long lab, PixOut& o) {
  const bool mean = ENS && p.strategy == AWX_FUSE_MEAN;
  const float T = p.T;
  // ---- pass 1: max / argmax of the fused logits (first index wins ties, NaN wins).  The runner-up is tracked
  // too: when it lies within the band of the maximum, a division by T may merge the two, or their fp32
  // probabilities may tie (the ECE's arg-max), and resolve_ties settles the pixel from global memory
  float vmax = 0.f, second = -INFINITY;
  int arg = 0;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    float v = ENS ? fuse_one(a[c], b[ENS ? c : 0], mean, w0, w1) : a[c];
    if (p.div_mode == 2) v = __fdiv_rn(v, T);
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
  o.pred = o.epred = arg;
  o.eamb = o.mamb = 0;
  bool tie_pred = C > 1 && second >= vmax - fmaf(fabsf(vmax), kPredBandRel, p.band_abs);
  bool tie_marg = false;

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
    float hm2 = 0.f, mlm2 = 0.f, mbest = 0.f, msecond = 0.f;
    int marg = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float m = fmaf(a[c], ka, b[ENS ? c : 0] * kb);
      hm2 = fmaf(m, lg2_approx(m + kEps), hm2);
      if (JS) mlm2 += (m > 0.f) ? m * lg2_approx(m) : 0.f;
      if (c == 0) {
        mbest = m;
      } else if (beats(m, mbest)) {
        msecond = mbest;
        mbest = m;
        marg = c;
      } else {
        msecond = fmaxf(msecond, m);
      }
    }
    tie_marg = C > 1 && msecond >= mbest * (1.f - kMargBand);
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
  if ((tie_pred || tie_marg) && ga != nullptr) {
    TieOut t;
    resolve_ties(ga, gb, p.HW, C, mean, w0, w1, p.div_mode, T, lab, tie_pred, tie_marg, t);
    if (tie_pred) {
      o.pred = t.pred;
      o.epred = t.epred;
      o.eamb = t.eamb;
    }
    if (tie_marg) {
      o.mpred = t.marg;
      o.mamb = t.mamb;
    }
  }
}


}  // namespace score_detail
}  // namespace awx
