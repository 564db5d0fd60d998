// awx_members_n: EnsembleDisagreementMetrics for ANY number of members N >= 2 (the two-member case of the
// reference's SegFormer + DeepLabV3+ pair runs inside awx_score; this is the general list form of
// evaluation/metrics.py:336-438):
//   mi[p]      = H(mean_k p_k) - mean_k H(p_k), logs of (p + 1e-8)                           :353-367
//   var[c,p]   = unbiased variance over members of p_k[c]   (torch.var: mean first, then squares)   :384-391
//   AUROC bins = histogram of mi over [0, hi) for pixels whose argmax(mean p) != / == label       :410-428
// A thread owns one pixel; members are streamed one after the other (C strided, coalesced loads each), the
// mean probabilities live in registers (C == 19) or local memory (generic C).  The variance needs the mean
// first, so its members are read twice (second pass from L2 for small frames).  Not a tuned kernel: the
// reference never builds ensembles of more than two members; this closes the API.
#include "awx_internal.cuh"

namespace awx {
namespace {

constexpr int kMaxMembers = 8;
constexpr float kEpsM = 1e-8f;

struct MembersParams {
  const float* m[kMaxMembers];
  int n;
  const void* labels;
  int label_mode, ignore_index;
  long long B, HW;
  int C;
  int NB;
  float scale, top;
  unsigned long long* pos;   // [NB]
  unsigned long long* neg;   // [NB]
  unsigned long long* counters;  // AWX_CNT_* [AWX_NUM_COUNTERS]
  float* mi;
  float* var;
};

// Rare path: arg-max of the mean member probabilities for a pixel whose two best classes the fp32 loop cannot
// separate (within 4e-6 relative).  fp64 softmaxes; classes within 8e-7 of the fp64 maximum are candidates;
// identical member logits tie for certain (first index); otherwise the fp32 mean of the fp32-rounded
// probabilities (torch: stack(...).mean(0), metrics.py:414-416) decides and, when the label is a candidate, the
// pixel is counted as ambiguous (AWX_CNT_MARG_AMBIG).  Same rule as resolve_ties (score_common.cuh).
static __device__ __noinline__ int resolve_marg_n(const MembersParams& p, int C, long long img, long long px, long long lab,
                                                  int* ambig) {
  double mean[AWX_MAX_CLASSES];
  for (int c = 0; c < C; ++c) mean[c] = 0.0;
  for (int k = 0; k < p.n; ++k) {
    const float* g = p.m[k] + img * C * p.HW + px;
    float mx = g[0];
    for (int c = 1; c < C; ++c) mx = fmaxf(mx, g[(long long)c * p.HW]);
    double s = 0.0;
    for (int c = 0; c < C; ++c) s += exp((double)__fsub_rn(g[(long long)c * p.HW], mx));
    for (int c = 0; c < C; ++c) mean[c] += exp((double)__fsub_rn(g[(long long)c * p.HW], mx)) / s;
  }
  int arg = 0;
  for (int c = 1; c < C; ++c)
    if (mean[c] > mean[arg]) arg = c;
  const double thr = mean[arg] * (1.0 - 8e-7);
  int ncand = 0, nsame = 0, lab_in = 0, first = -1, best = arg;
  float mbest = -1.f;
  for (int c = 0; c < C; ++c) {
    if (!(mean[c] >= thr)) continue;
    if (first < 0) first = c;
    ++ncand;
    lab_in |= (long long)c == lab;
    bool same = true;
    float acc = 0.f;
    for (int k = 0; k < p.n; ++k) {
      const float* g = p.m[k] + img * C * p.HW + px;
      same = same && g[(long long)c * p.HW] == g[(long long)first * p.HW];
      float mx = g[0];
      for (int j = 1; j < C; ++j) mx = fmaxf(mx, g[(long long)j * p.HW]);
      double s = 0.0;
      for (int j = 0; j < C; ++j) s += exp((double)__fsub_rn(g[(long long)j * p.HW], mx));
      acc = __fadd_rn(acc, (float)(exp((double)__fsub_rn(g[(long long)c * p.HW], mx)) / s));
    }
    nsame += same;
    const float m32 = __fdiv_rn(acc, (float)p.n);
    if (m32 > mbest) {
      mbest = m32;
      best = c;
    }
  }
  if (ncand <= 1) return arg;
  if (nsame == ncand) return first;
  *ambig = lab_in;
  return best;
}

template <int CS>
__global__ void __launch_bounds__(256) members_kernel(const __grid_constant__ MembersParams p) {
  constexpr int CA = CS > 0 ? CS : AWX_MAX_CLASSES;
  const int C = CS > 0 ? CS : p.C;
  const long long total = p.B * p.HW;
  const float inv_n = __fdiv_rn(1.0f, (float)p.n);
  unsigned n_valid = 0, n_wrong = 0, n_mamb = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / p.HW, px = i - img * p.HW;
    float mean[CA];
#pragma unroll
    for (int c = 0; c < C; ++c) mean[c] = 0.f;
    float h_each = 0.f;
    for (int k = 0; k < p.n; ++k) {
      const float* g = p.m[k] + img * C * p.HW + px;
      float x[CA];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        x[c] = ld_stream(g + (long long)c * p.HW);
        mx = fmaxf(mx, x[c]);
      }
      float s = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        x[c] = ex2_approx((x[c] - mx) * kLog2e);
        s += x[c];
      }
      const float r = __frcp_rn(s);
      float h = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float pr = x[c] * r;
        mean[c] += pr;
        h = fmaf(pr, lg2_approx(pr + kEpsM), h);
      }
      h_each += h;
    }
    float hm = 0.f, best = -1.f, second = -1.f;
    int arg = 0;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      mean[c] *= inv_n;
      hm = fmaf(mean[c], lg2_approx(mean[c] + kEpsM), hm);
      if (mean[c] > best) {
        second = best;
        best = mean[c];
        arg = c;
      } else {
        second = fmaxf(second, mean[c]);
      }
    }
    const bool tie = C > 1 && second >= best * (1.f - 4e-6f);
    const float mi = kLn2 * (h_each * inv_n - hm);  // H(mean) - mean_k H_k with H = -sum p ln(p+eps)
    if (p.mi) p.mi[i] = mi;
    if (p.var) {
      // second pass: deviations from the mean (torch.var's two-pass form), unbiased 1/(N-1)
      float acc[CA];
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = 0.f;
      for (int k = 0; k < p.n; ++k) {
        const float* g = p.m[k] + img * C * p.HW + px;
        float x[CA];
        float mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          x[c] = g[(long long)c * p.HW];
          mx = fmaxf(mx, x[c]);
        }
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          x[c] = ex2_approx((x[c] - mx) * kLog2e);
          s += x[c];
        }
        const float r = __frcp_rn(s);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float d = x[c] * r - mean[c];
          acc[c] = fmaf(d, d, acc[c]);
        }
      }
      const float k1 = __fdiv_rn(1.0f, (float)(p.n - 1));
      float* vo = p.var + img * C * p.HW + px;
#pragma unroll
      for (int c = 0; c < C; ++c) vo[(long long)c * p.HW] = acc[c] * k1;
    }
    if (p.labels) {
      long long y;
      if (p.label_mode == AWX_LABEL_U8)
        y = static_cast<const uint8_t*>(p.labels)[i];
      else
        y = static_cast<const long long*>(p.labels)[i];
      if (y != p.ignore_index) {
        ++n_valid;
        if (tie) {
          int amb = 0;
          arg = resolve_marg_n(p, C, img, px, y, &amb);
          n_mamb += amb;
        }
        const bool wrong = y != arg;
        n_wrong += wrong;
        if (p.NB > 0) {
          const float qv = fminf(fmaxf(mi * p.scale, 0.f), p.top);
          const int b = __float_as_int(__fadd_rd(qv, 8388608.f)) & 0x7fffff;
          atomicAdd((wrong ? p.pos : p.neg) + b, 1ull);
        }
      }
    }
  }
  if (p.labels) {
    const unsigned v = __reduce_add_sync(0xffffffffu, n_valid), w = __reduce_add_sync(0xffffffffu, n_wrong);
    const unsigned m = __reduce_add_sync(0xffffffffu, n_mamb);
    if ((threadIdx.x & 31) == 0) {
      if (v) atomicAdd(p.counters + AWX_CNT_VALID, (unsigned long long)v);
      if (w) atomicAdd(p.counters + AWX_CNT_ENS_WRONG, (unsigned long long)w);
      if (m) atomicAdd(p.counters + AWX_CNT_MARG_AMBIG, (unsigned long long)m);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(p.counters + AWX_CNT_PIXELS, (unsigned long long)total);
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" int awx_members_n(const float* const* members, int32_t n_members, const void* labels, int32_t label_dtype,
                             int64_t batch, int32_t C, int64_t pixels_per_image, int32_t ignore_index, int32_t auroc_bins,
                             float auroc_hi, int64_t* auroc_pos, int64_t* auroc_neg, int64_t* counters, float* mi_out,
                             float* var_out, void* stream) {
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0, AWX_E_ARG, "awx_members_n: negative size");
  AWX_REQUIRE(n_members >= 2, AWX_E_ARG, "awx_members_n: need at least 2 members");
  AWX_REQUIRE(n_members <= kMaxMembers, AWX_E_UNSUPPORTED, "awx_members_n: %d members > %d", n_members, kMaxMembers);
  AWX_REQUIRE(C >= 1 && C <= AWX_MAX_CLASSES, AWX_E_UNSUPPORTED, "awx_members_n: num_classes %d outside 1..%d", C, AWX_MAX_CLASSES);
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;
  AWX_REQUIRE(members != nullptr, AWX_E_ARG, "awx_members_n: members is NULL");
  AWX_REQUIRE(labels == nullptr || (counters != nullptr && (auroc_bins == 0 || (auroc_pos && auroc_neg))), AWX_E_ARG,
              "awx_members_n: labels given but counters / histograms are NULL");
  AWX_REQUIRE(auroc_bins >= 0 && auroc_bins <= AWX_MAX_AUROC_BINS, AWX_E_UNSUPPORTED, "awx_members_n: auroc_bins outside 0..%d", AWX_MAX_AUROC_BINS);
  AWX_REQUIRE(label_dtype == AWX_LABEL_U8 || label_dtype == AWX_LABEL_I64, AWX_E_ARG, "awx_members_n: unknown label dtype");
  MembersParams p{};
  for (int k = 0; k < n_members; ++k) {
    AWX_REQUIRE(members[k] != nullptr, AWX_E_ARG, "awx_members_n: member %d is NULL", k);
    p.m[k] = members[k];
  }
  p.n = n_members;
  p.labels = labels;
  p.label_mode = label_dtype;
  p.ignore_index = ignore_index;
  p.B = batch;
  p.HW = pixels_per_image;
  p.C = C;
  p.NB = labels ? auroc_bins : 0;
  p.scale = p.NB > 0 ? (float)p.NB / auroc_hi : 0.f;
  p.top = p.NB > 0 ? (float)(p.NB - 1) : 0.f;
  p.pos = reinterpret_cast<unsigned long long*>(auroc_pos);
  p.neg = reinterpret_cast<unsigned long long*>(auroc_neg);
  p.counters = reinterpret_cast<unsigned long long*>(counters);
  p.mi = mi_out;
  p.var = var_out;
  const long long total = batch * pixels_per_image;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 19)
    members_kernel<19><<<(unsigned)blocks, 256, 0, s>>>(p);
  else
    members_kernel<0><<<(unsigned)blocks, 256, 0, s>>>(p);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}
