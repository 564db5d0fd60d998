// awx_score: fuse -> softmax -> argmax -> confusion / ECE / AUROC bins in one pass over HBM.
//
// Data layout: logits fp32 [B,C,HW] (NCHW).  A thread owns PX consecutive pixels of one image
// and reads, for every class plane, PX consecutive floats (64-bit loads for PX=2): a warp
// covers 256 contiguous bytes per plane, all 2*C loads are issued before any arithmetic so
// that ~2*C*8 B per thread are in flight.  Statistics go to shared-memory histograms private
// to the CTA (ECE bins private to the warp, no atomics), flushed once per CTA with 64-bit
// global atomics; the grid is persistent (SM count x resident CTAs).
//
// Arithmetic notes (see DESIGN.md "numerics"):
//  * fused logits use separately rounded fp32 ops (__fmul_rn/__fadd_rn/__fdiv_rn): argmax and
//    the optional `fused` output are bit-exact w.r.t. torch eager.
//  * softmax sums use ex2.approx; a pixel whose confidence lands within 16 ulp of an interior
//    ECE edge is recomputed in fp64 (exact_confidence) and flagged ambiguous when it is still
//    within 3 ulp, i.e. inside the reference's own fp32 rounding noise.
//  * entropies use H(p) = ln S - (sum e_c d_c)/S (d_c = x_c - max) instead of C logarithms.
#include <cstdlib>

#include <cstring>

#include "score_common.cuh"

namespace awx {
using namespace score_detail;

// score_v2.cu
bool score_v2_supported(const ScoreParams& p);
int launch_score_v2(const ScoreParams& p, bool ens, bool js, cudaStream_t stream);

namespace {
template <int CS, int PX, bool ENS, bool JS>
__global__ void __launch_bounds__(kThreads, CS > 0 ? 2 : 1) score_kernel(const __grid_constant__ ScoreParams p) {
  constexpr int CA = CS > 0 ? CS : AWX_MAX_CLASSES;
  constexpr int CB = ENS ? CA : 1;
  const int C = CS > 0 ? CS : p.C;
  const int nb = p.nb;
  const int NB = p.auroc_bins;
  const bool have_labels = p.labels != nullptr;

  extern __shared__ __align__(16) unsigned char smem[];
  unsigned long long* s_ece_sum = reinterpret_cast<unsigned long long*>(smem);            // [kWarps][nb]
  unsigned* s_ece_cnt = reinterpret_cast<unsigned*>(s_ece_sum + kWarps * nb);              // [kWarps][nb]
  unsigned* s_ece_cor = s_ece_cnt + kWarps * nb;                                          // [kWarps][nb]
  unsigned* s_conf = s_ece_cor + kWarps * nb;                                             // [C*C]
  unsigned* s_auroc = s_conf + C * C;                                                     // [2*NB]
  unsigned* s_cnt = s_auroc + 2 * NB;                                                     // [AWX_NUM_COUNTERS]
  float* s_edges = reinterpret_cast<float*>(s_cnt + AWX_NUM_COUNTERS);                                   // [nb+1]
  {
    const int words = kWarps * nb * 4 + C * C + 2 * NB + AWX_NUM_COUNTERS;  // u64 counts as two words
    unsigned* w = reinterpret_cast<unsigned*>(smem);
    for (int i = threadIdx.x; i < words; i += kThreads) w[i] = 0u;
    for (int i = threadIdx.x; i <= nb; i += kThreads) s_edges[i] = p.edges[i];
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long HW = p.HW;
  const long long gpi = HW / PX;
  const long long total = p.B * gpi;
  const long long stride = (long long)gridDim.x * kThreads;
  const bool mean = ENS && p.strategy == AWX_FUSE_MEAN;
  unsigned n_valid = 0, n_correct = 0, n_bad = 0, n_ambig = 0, n_wrong = 0, n_pick = 0, n_nobin = 0, n_pix = 0;
  unsigned n_mamb = 0, n_eamb = 0;

  for (long long g0 = (long long)blockIdx.x * kThreads + warp * 32; g0 < total; g0 += stride) {
    const long long g = g0 + lane;
    const bool act = g < total;
    const long long gg = act ? g : total - 1;
    const long long img = gg / gpi;
    const long long px = (gg - img * gpi) * PX;
    const float* ga = p.a + img * C * HW + px;
    const float* gb = ENS ? p.b + img * C * HW + px : nullptr;

    float a[PX][CA];
    float b[PX][CB];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (PX == 2) {
        const float2 t = ld_stream2(ga + c * HW);
        a[0][c] = t.x;
        a[PX - 1][c] = t.y;
      } else {
        a[0][c] = ld_stream(ga + c * HW);
      }
    }
    if (ENS) {
#pragma unroll
      for (int c = 0; c < C; ++c) {
        if (PX == 2) {
          const float2 t = ld_stream2(gb + c * HW);
          b[0][ENS ? c : 0] = t.x;
          b[PX - 1][ENS ? c : 0] = t.y;
        } else {
          b[0][ENS ? c : 0] = ld_stream(gb + c * HW);
        }
      }
    }
    long long lab[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) lab[j] = p.ignore_index;
    if (have_labels) {
      const long long li = img * HW + px;
      if (p.label_mode == AWX_LABEL_U8) {
        const uint8_t* l8 = static_cast<const uint8_t*>(p.labels) + li;
#pragma unroll
        for (int j = 0; j < PX; ++j) lab[j] = l8[j];
      } else {
        const long long* l64 = static_cast<const long long*>(p.labels) + li;
#pragma unroll
        for (int j = 0; j < PX; ++j) lab[j] = l64[j];
      }
    }

    PixOut o[PX];
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      float w0 = p.w0, w1 = p.w1, amax = 0.f, bmax = 0.f;
      int pick_ambig = 0;
      if (ENS) {
        amax = a[j][0];
        bmax = b[j][0];
#pragma unroll
        for (int c = 1; c < C; ++c) {
          amax = fmaxf(amax, a[j][c]);
          bmax = fmaxf(bmax, b[j][ENS ? c : 0]);
        }
        if (p.strategy == AWX_FUSE_MAXCONF) {
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int c = 0; c < C; ++c) {
            sa += ex2_approx((a[j][c] - amax) * kLog2e);
            sb += ex2_approx((b[j][ENS ? c : 0] - bmax) * kLog2e);
          }
          const float ca = __frcp_rn(sa), cb = __frcp_rn(sb);
          w0 = ca > cb ? 1.f : 0.f;
          w1 = 1.f - w0;
          pick_ambig = fabsf(ca - cb) <= 4.8e-7f * fmaxf(ca, cb);
        }
      }
      if (p.fused != nullptr && act) {
        float* fo = p.fused + img * C * HW + px + j;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float v = ENS ? fuse_one(a[j][c], b[j][ENS ? c : 0], mean, w0, w1) : a[j][c];
          if (p.div_mode) v = __fdiv_rn(v, p.T);
          fo[c * HW] = v;
        }
      }
      o[j].mi = 0.f;
      o[j].js = 0.f;
      o[j].mpred = 0;
      score_pixel<CS, ENS, JS>(a[j], b[j], C, p, s_edges, ga + j, ENS ? gb + j : nullptr, w0, w1, amax, bmax, lab[j], o[j]);
      if (act) n_pick += pick_ambig;
    }

    // ---- optional maps
    if (act) {
      const long long li = img * HW + px;
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        if (p.pred) {
          if (p.pred_dtype == AWX_PRED_U8)
            static_cast<uint8_t*>(p.pred)[li + j] = (uint8_t)o[j].pred;
          else
            static_cast<long long*>(p.pred)[li + j] = o[j].pred;
        }
        if (p.conf) p.conf[li + j] = o[j].conf;
        if (ENS && p.mi) p.mi[li + j] = o[j].mi;
        if (ENS && JS && p.js) p.js[li + j] = o[j].js;
      }
    }

    // ---- statistics
    if (have_labels) {
#pragma unroll
      for (int j = 0; j < PX; ++j) {
        const long long t = lab[j];
        const bool valid = act && t != (long long)p.ignore_index;
        const int pred = o[j].pred;
        // the ECE's accuracy term compares the arg-max of the PROBABILITIES (metrics.py:161-170); pixel accuracy
        // and the confusion matrix use the arg-max of the logits
        const bool correct = valid && t == (long long)o[j].epred;
        n_pix += act;
        n_valid += valid;
        n_correct += valid && t == (long long)pred;
        if (valid) {
          // confusion index exactly as torch evaluates targets*C + predictions (metrics.py:68)
          const long long idx = (p.label_mode == AWX_LABEL_U8 ? ((t * C) & 0xff) : t * C) + pred;
          if (idx >= 0 && idx < (long long)C * C)
            atomicAdd(&s_conf[idx], 1u);
          else
            ++n_bad;
          n_ambig += o[j].ambig;
          n_eamb += o[j].eamb;
          n_nobin += o[j].bin < 0;
          if (ENS && NB > 0) {
            const bool wrong = t != (long long)o[j].mpred;
            n_wrong += wrong;
            n_mamb += o[j].mamb;
            float q = floorf(o[j].mi * p.auroc_scale);
            q = is_nan(q) ? 0.f : q;
            const int mb = (int)fminf(fmaxf(q, 0.f), (float)(NB - 1));
            atomicAdd(&s_auroc[(wrong ? 0 : NB) + mb], 1u);
          }
        }
        // ECE: warp-aggregated, warp-private bins (no atomics).  conf is accumulated in
        // 2^-31 fixed point: exact for conf >= 2^-7, order independent.
        const int bin = valid ? o[j].bin : -1;
        const unsigned fx = bin >= 0 ? __float2uint_rz(o[j].conf * 2147483648.f) : 0u;
        unsigned todo = __ballot_sync(0xffffffffu, bin >= 0);
        while (todo) {
          const int leader = __ffs(todo) - 1;
          const int bl = __shfl_sync(0xffffffffu, bin, leader);
          const bool mine = bin == bl;
          const unsigned peers = __ballot_sync(0xffffffffu, mine);
          const unsigned ncor = __popc(__ballot_sync(0xffffffffu, mine && correct));
          const unsigned lo = __reduce_add_sync(0xffffffffu, mine ? (fx & 0xffffu) : 0u);
          const unsigned hi = __reduce_add_sync(0xffffffffu, mine ? (fx >> 16) : 0u);
          if (lane == leader) {
            s_ece_cnt[warp * nb + bl] += __popc(peers);
            s_ece_cor[warp * nb + bl] += ncor;
            s_ece_sum[warp * nb + bl] += ((unsigned long long)hi << 16) + lo;
          }
          todo &= ~peers;
        }
        __syncwarp();
      }
    }
  }

  if (!have_labels) return;
  // ---- per-thread counters -> CTA
  {
    unsigned v[10] = {n_valid, n_correct, n_bad, n_ambig, n_wrong, n_pick, n_nobin, n_pix, n_mamb, n_eamb};
#pragma unroll
    for (int k = 0; k < 10; ++k) {
      const unsigned s = __reduce_add_sync(0xffffffffu, v[k]);
      if (lane == 0 && s) atomicAdd(&s_cnt[k], s);
    }
  }
  __syncthreads();
  // ---- flush CTA histograms with 64-bit global atomics (non-zero bins only)
  unsigned long long* bins = p.bins;
  for (int i = threadIdx.x; i < C * C; i += kThreads)
    if (s_conf[i]) atomicAdd(bins + p.lay.confusion + i, (unsigned long long)s_conf[i]);
  for (int i = threadIdx.x; i < 2 * NB; i += kThreads)
    if (s_auroc[i]) atomicAdd(bins + (i < NB ? p.lay.auroc_pos + i : p.lay.auroc_neg + (i - NB)), (unsigned long long)s_auroc[i]);
  for (int i = threadIdx.x; i < nb; i += kThreads) {
    unsigned long long cnt = 0, cor = 0, sum_hi = 0, sum_lo = 0;
    for (int w = 0; w < kWarps; ++w) {
      cnt += s_ece_cnt[w * nb + i];
      cor += s_ece_cor[w * nb + i];
      const unsigned long long s = s_ece_sum[w * nb + i];
      sum_hi += s >> 32;
      sum_lo += s & 0xffffffffull;
    }
    if (cnt) {
      atomicAdd(bins + p.lay.ece_count + i, cnt);
      if (cor) atomicAdd(bins + p.lay.ece_correct + i, cor);
      atomicAdd(bins + p.lay.ece_conf_hi + i, sum_hi);
      atomicAdd(bins + p.lay.ece_conf_lo + i, sum_lo);
    }
  }
  if (threadIdx.x < AWX_NUM_COUNTERS && s_cnt[threadIdx.x]) atomicAdd(bins + p.lay.counters + threadIdx.x, (unsigned long long)s_cnt[threadIdx.x]);
}

size_t score_smem_bytes(int C, int nb, int NB) {
  return (size_t)kWarps * nb * 16 + ((size_t)C * C + 2 * (size_t)NB + AWX_NUM_COUNTERS) * 4 + (size_t)(nb + 1) * 4;
}

template <int CS, int PX, bool ENS, bool JS>
int launch_score(const ScoreParams& p, cudaStream_t stream) {
  auto kern = score_kernel<CS, PX, ENS, JS>;
  const size_t smem = score_smem_bytes(p.C, p.nb, p.auroc_bins);
  AWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int occ = 0;
  AWX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, smem));
  AWX_REQUIRE(occ > 0, AWX_E_UNSUPPORTED, "awx_score: kernel does not fit on an SM (smem %zu B)", smem);
  const long long groups = p.B * (p.HW / PX);
  long long blocks = (groups + kThreads - 1) / kThreads;
  const long long resident = (long long)sm_count() * occ;
  if (blocks > resident) blocks = resident;
  if (blocks < 1) blocks = 1;
  kern<<<(unsigned)blocks, kThreads, smem, stream>>>(p);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

template <int CS, int PX>
int dispatch_mode(const ScoreParams& p, bool ens, bool js, cudaStream_t s) {
  if (!ens) return launch_score<CS, PX, false, false>(p, s);
  if (js) return launch_score<CS, PX, true, true>(p, s);
  return launch_score<CS, PX, true, false>(p, s);
}

// ------------------------------------------------------------------ member variance map
__global__ void __launch_bounds__(256) variance_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                        float* __restrict__ out, long long B, int C, long long HW) {
  const long long total = B * HW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long img = i / HW, px = i - img * HW;
    const float* ga = a + img * C * HW + px;
    const float* gb = b + img * C * HW + px;
    float amax = ga[0], bmax = gb[0];
    for (int c = 1; c < C; ++c) {
      amax = fmaxf(amax, ga[c * HW]);
      bmax = fmaxf(bmax, gb[c * HW]);
    }
    float sa = 0.f, sb = 0.f;
    for (int c = 0; c < C; ++c) {
      sa += expf(ga[c * HW] - amax);
      sb += expf(gb[c * HW] - bmax);
    }
    float* go = out + img * C * HW + px;
    for (int c = 0; c < C; ++c) {
      const float pa = expf(ga[c * HW] - amax) / sa;
      const float pb = expf(gb[c * HW] - bmax) / sb;
      // torch.var over 2 samples, unbiased: sum (x - mean)^2 / (2 - 1)
      const float m = (pa + pb) * 0.5f;
      const float d0 = pa - m, d1 = pb - m;
      go[c * HW] = d0 * d0 + d1 * d1;
    }
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

// The branch-free quotient x * RN(1/T) corrected by one FMA residual step is the correctly rounded x / T for
// every divisor except those whose significand is all ones (Markstein); such a T takes the true division.
static bool all_ones_significand(float t) {
  uint32_t u;
  memcpy(&u, &t, sizeof(u));
  return (u & 0x7fffffu) == 0x7fffffu;
}

extern "C" int awx_bins_layout(int32_t C, int32_t nb, int32_t NB, AwxBinsLayout* out) {
  AWX_REQUIRE(out != nullptr, AWX_E_ARG, "awx_bins_layout: out is NULL");
  AWX_REQUIRE(C >= 1 && C <= AWX_MAX_CLASSES, AWX_E_UNSUPPORTED, "num_classes %d outside 1..%d", C, AWX_MAX_CLASSES);
  AWX_REQUIRE(nb >= 1 && nb <= AWX_MAX_ECE_BINS, AWX_E_UNSUPPORTED, "ece_bins %d outside 1..%d", nb, AWX_MAX_ECE_BINS);
  AWX_REQUIRE(NB >= 0 && NB <= AWX_MAX_AUROC_BINS, AWX_E_UNSUPPORTED, "auroc_bins %d outside 0..%d", NB, AWX_MAX_AUROC_BINS);
  int64_t o = 0;
  out->confusion = o; o += (int64_t)C * C;
  out->ece_count = o; o += nb;
  out->ece_correct = o; o += nb;
  out->ece_conf_hi = o; o += nb;
  out->ece_conf_lo = o; o += nb;
  out->counters = o; o += AWX_NUM_COUNTERS;
  // the AUROC histograms come last: a buffer laid out for NB > 0 is a valid target for launches without them
  // (single-member scoring), which makes one buffer size fit every rank of a sharded evaluation
  out->auroc_pos = o; o += NB;
  out->auroc_neg = o; o += NB;
  out->total_words = o;
  return AWX_OK;
}

extern "C" int awx_score(const float* logits_a, const float* logits_b, const void* labels, int64_t batch,
                         int64_t pixels_per_image, const AwxScoreConfig* cfg, int64_t* bins, const AwxScoreMaps* maps,
                         void* stream) {
  AWX_REQUIRE(cfg != nullptr, AWX_E_ARG, "awx_score: cfg is NULL");
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0, AWX_E_ARG, "awx_score: negative size");
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;  // empty input: nothing to add to the bins
  AWX_REQUIRE(logits_a != nullptr, AWX_E_ARG, "awx_score: logits_a is NULL");
  const bool ens = cfg->strategy != AWX_FUSE_SINGLE;
  AWX_REQUIRE(cfg->strategy >= AWX_FUSE_SINGLE && cfg->strategy <= AWX_FUSE_MEAN, AWX_E_ARG, "awx_score: unknown strategy %d", cfg->strategy);
  AWX_REQUIRE(!ens || logits_b != nullptr, AWX_E_ARG, "awx_score: ensemble strategy needs logits_b");
  AWX_REQUIRE(labels == nullptr || bins != nullptr, AWX_E_ARG, "awx_score: labels given but bins is NULL");
  AWX_REQUIRE(cfg->label_dtype == AWX_LABEL_U8 || cfg->label_dtype == AWX_LABEL_I64, AWX_E_ARG, "awx_score: unknown label dtype %d", cfg->label_dtype);
  ScoreParams p{};
  int rc = awx_bins_layout(cfg->num_classes, cfg->ece_bins, ens ? cfg->auroc_bins : 0, &p.lay);
  if (rc != AWX_OK) return rc;
  AWX_REQUIRE(((uintptr_t)logits_a & 3) == 0 && ((uintptr_t)logits_b & 3) == 0, AWX_E_ALIGN, "awx_score: logits must be 4-byte aligned");
  AWX_REQUIRE(cfg->label_dtype != AWX_LABEL_I64 || ((uintptr_t)labels & 7) == 0, AWX_E_ALIGN, "awx_score: int64 labels must be 8-byte aligned");

  p.a = logits_a;
  p.b = ens ? logits_b : nullptr;
  p.labels = labels;
  p.B = batch;
  p.HW = pixels_per_image;
  p.C = cfg->num_classes;
  p.strategy = cfg->strategy;
  p.w0 = cfg->w0;
  p.w1 = cfg->w1;
  p.T = cfg->temperature;
  const bool want_fused = maps && maps->fused;
  if (!cfg->use_temperature || cfg->temperature == 1.0f)
    p.div_mode = 0;
  else if (cfg->temperature > 0.f && std::isfinite(cfg->temperature) && !want_fused && !all_ones_significand(cfg->temperature))
    p.div_mode = 1;
  else
    p.div_mode = 2;
  p.rT = (float)(1.0 / (double)cfg->temperature);
  p.band_abs = p.div_mode == 1 ? 4e-7f * cfg->temperature : 4e-7f;
  p.kz = p.div_mode == 1 ? (float)(1.4426950408889634 / (double)cfg->temperature) : kLog2e;
  p.label_mode = cfg->label_dtype;
  p.ignore_index = cfg->ignore_index;
  p.nb = cfg->ece_bins;
  p.nbf = (float)cfg->ece_bins;
  p.auroc_bins = ens ? cfg->auroc_bins : 0;
  p.auroc_scale = p.auroc_bins > 0 ? (float)p.auroc_bins / cfg->auroc_hi : 0.f;
  p.auroc_top = p.auroc_bins > 0 ? (float)(p.auroc_bins - 1) : 0.f;
  p.bins = reinterpret_cast<unsigned long long*>(bins);
  for (int i = 0; i <= cfg->ece_bins; ++i) p.edges[i] = cfg->ece_edges[i];
  bool js = false;
  if (maps) {
    p.pred = maps->pred;
    p.pred_dtype = maps->pred_dtype;
    p.fused = maps->fused;
    p.conf = maps->conf;
    p.mi = ens ? maps->mi : nullptr;
    p.js = ens ? maps->js : nullptr;
    js = p.js != nullptr;
    AWX_REQUIRE(maps->pred == nullptr || maps->pred_dtype == AWX_PRED_U8 || maps->pred_dtype == AWX_PRED_I64, AWX_E_ARG, "awx_score: unknown pred dtype");
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  // kernel v2 (TMA ring + packed math) whenever its layout requirements hold; AWX_SCORE_KERNEL=v1
  // forces the generic register-resident kernel (A/B measurements, parity tests of both paths)
  const char* dbg = getenv("AWX_DEBUG_SKIP_MATH");
  p.debug_skip = (dbg && dbg[0] == '1') ? 1 : 0;
  const char* force = getenv("AWX_SCORE_KERNEL");
  const bool allow_v2 = !(force && force[0] == 'v' && force[1] == '1');
  if (allow_v2 && score_v2_supported(p)) return launch_score_v2(p, ens, js, s);
  const bool vec2 = (pixels_per_image % 2 == 0) && ((uintptr_t)logits_a & 7) == 0 && ((uintptr_t)logits_b & 7) == 0;
  if (p.C == 19) return vec2 ? dispatch_mode<19, 2>(p, ens, js, s) : dispatch_mode<19, 1>(p, ens, js, s);
  return dispatch_mode<0, 1>(p, ens, js, s);
}

extern "C" int awx_member_variance(const float* a, const float* b, float* out, int64_t batch, int32_t C,
                                   int64_t pixels_per_image, void* stream) {
  AWX_REQUIRE(a && b && out, AWX_E_ARG, "awx_member_variance: NULL pointer");
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0 && C >= 1, AWX_E_ARG, "awx_member_variance: bad size");
  const long long total = batch * pixels_per_image;
  if (total == 0) return AWX_OK;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  variance_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, out, batch, C, pixels_per_image);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}
