/*
 * awx.h -- C ABI of libawx.so, the B200 (sm_100a) implementation of the
 * adverse-weather robustness-evaluation hot path.
 *
 * The reference (A-SHOJAEI/adverse-weather-semantic-segmentation-robustness-benchmark)
 * is pure Python and has no FFI of its own; its boundary for this path is its public
 * class API.  Each entry point below therefore names the reference method(s) whose
 * per-pixel arithmetic it replaces (paths relative to
 * src/adverse_weather_semantic_segmentation_robustness_benchmark/).  INTEGRATION.md
 * shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every data pointer is a DEVICE pointer unless its comment says HOST.
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on it.
 *   - kernels never allocate: the caller owns inputs, outputs, bins and workspaces.
 *   - return value: 0 = ok; <0 = argument error (AWX_E_*); >0 = cudaError_t.
 *     awx_last_error() returns the thread-local message of the last failure.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef AWX_H_
#define AWX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AWX_VERSION 101 /* major*100 + minor */

#define AWX_OK 0
#define AWX_E_ARG (-1)         /* null / negative / inconsistent argument       */
#define AWX_E_UNSUPPORTED (-2) /* e.g. num_classes > AWX_MAX_CLASSES            */
#define AWX_E_ALIGN (-3)       /* pointer not aligned as documented             */

#define AWX_MAX_CLASSES 64
#define AWX_MAX_ECE_BINS 64
#define AWX_MAX_AUROC_BINS 8192
#define AWX_NUM_COUNTERS 16 /* words of the `counters` block of a bins buffer (AWX_CNT_*) */

int awx_version(void);
const char* awx_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
int64_t awx_launch_count(void);
/* SM count / compute capability of the current device (for grid sizing by callers). */
int awx_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------
 * Scoring: fuse -> softmax -> argmax -> confusion / ECE / AUROC bins, one pass over HBM.
 *
 * Replaces the per-pixel arithmetic of
 *   EnsembleModel.forward fusion            models/model.py:442-462
 *   EnsembleModel.get_ensemble_disagreement models/model.py:498-513
 *   IoUMetrics.compute_iou / compute_pixel_accuracy   evaluation/metrics.py:34-123
 *   ConfidenceCalibration.compute_ece                 evaluation/metrics.py:143-226
 *   EnsembleDisagreementMetrics.compute_disagreement_map / _auroc / _jensen_shannon_divergence
 *                                                     evaluation/metrics.py:336-369, 393-467
 * ---------------------------------------------------------------------------------- */

enum {
  AWX_FUSE_SINGLE = 0,   /* one member: logits_b ignored                                    */
  AWX_FUSE_WEIGHTED = 1, /* w0*a + w1*b, three separately rounded fp32 ops (model.py:445)   */
  AWX_FUSE_MAXCONF = 2,  /* member with the larger max-softmax, strict > (model.py:449-455) */
  AWX_FUSE_MEAN = 3      /* (a + b) / 2 (model.py:457)                                      */
};

enum {
  AWX_LABEL_U8 = 0, /* uint8 labels; confusion index (t*C mod 256)+pred, the reference's
                       uint8 wrap (metrics.py:68)                              */
  AWX_LABEL_I64 = 1 /* int64 labels; index t*C+pred                                       */
};

enum { AWX_PRED_U8 = 0, AWX_PRED_I64 = 1 };

typedef struct AwxScoreConfig {
  int32_t num_classes;     /* C, 1..AWX_MAX_CLASSES (C == 19 has a register-resident kernel) */
  int32_t strategy;        /* AWX_FUSE_*                                                     */
  float w0, w1;            /* softmax(ensemble_weights) computed by the host in fp32         */
  float temperature;       /* divisor of the fused logits                                    */
  int32_t use_temperature; /* 0: no division (temperature_scaling=False)                     */
  int32_t label_dtype;     /* AWX_LABEL_*                                                    */
  int32_t ignore_index;    /* 255                                                            */
  int32_t ece_bins;        /* 1..AWX_MAX_ECE_BINS                                            */
  int32_t auroc_bins;      /* 0 = no AUROC histogram; else 1..AWX_MAX_AUROC_BINS             */
  float auroc_hi;          /* MI histogram covers [0, auroc_hi) linearly (ln 2 for 2 members)*/
  float ece_edges[AWX_MAX_ECE_BINS + 1]; /* torch.linspace(0,1,bins+1) as fp32, from the host */
} AwxScoreConfig;

/* Optional per-pixel outputs; any pointer may be NULL. */
typedef struct AwxScoreMaps {
  void* pred;         /* [B,HW] argmax of the fused (temperature-scaled) logits            */
  int32_t pred_dtype; /* AWX_PRED_*                                                        */
  int32_t reserved;
  float* fused; /* [B,C,HW] fused logits, bit-exact (true fp32 division by temperature)    */
  float* conf;  /* [B,HW] max softmax probability of the fused logits                      */
  float* mi;    /* [B,HW] mutual-information disagreement (metrics.py:358-367)             */
  float* js;    /* [B,HW] 0.5*[KL(m||p)+KL(m||q)] (model.py:505-511)                       */
} AwxScoreMaps;

/* Word offsets into the int64 `bins` buffer (all counts; summing buffers of several
 * ranks/launches with a plain integer add is exact and order-independent).
 *   confusion  [C*C]      rows = target, cols = prediction
 *   ece_count  [nb]       pixels with edges[b] < conf <= edges[b+1]
 *   ece_correct[nb]       of those, pred == target
 *   ece_conf_hi[nb], ece_conf_lo[nb]   sum of conf in 2^-31 fixed point = hi*2^32 + lo; both words are plain
 *                          accumulators (lo may exceed 2^32), so equal sums need not be equal word by word
 *   counters   [AWX_NUM_COUNTERS]  AWX_CNT_*
 *   auroc_pos  [NB], auroc_neg[NB]     MI histogram of wrong / right ensemble pixels; last, so that the layout
 *                          for NB = 0 is a prefix of the layout for any NB                 */
typedef struct AwxBinsLayout { /* word offsets; the field order of this struct is not the memory order */
  int64_t confusion, ece_count, ece_correct, ece_conf_hi, ece_conf_lo;
  int64_t auroc_pos, auroc_neg, counters, total_words;
} AwxBinsLayout;

enum {
  AWX_CNT_VALID = 0,     /* label != ignore_index                                          */
  AWX_CNT_CORRECT = 1,   /* valid and argmax(fused LOGITS) == label (pixel accuracy, :110-123) */
  AWX_CNT_BAD_LABEL = 2, /* valid but confusion index outside [0,C*C): reference raises    */
  AWX_CNT_ECE_AMBIG = 3, /* valid pixels whose confidence is within 3 fp32 ulp of an
                            interior bin edge (the reference's own rounding noise)          */
  AWX_CNT_ENS_WRONG = 4, /* valid and argmax(mean member prob) != label (AUROC positives)  */
  AWX_CNT_PICK_AMBIG = 5,/* MAXCONF: member confidences within 4 ulp of each other         */
  AWX_CNT_NO_BIN = 6,    /* valid pixels whose confidence fell in no bin (0, NaN)          */
  AWX_CNT_PIXELS = 7,    /* all pixels seen                                                */
  /* The reference takes two of its arg-maxima over fp32 PROBABILITIES (torch softmax, not reproducible bit for
   * bit across builds / thread counts).  A pixel whose top classes are closer than that arithmetic's rounding
   * noise is re-evaluated in fp64 (resolve_ties, csrc/score_common.cuh); when the label is one of the tied
   * classes -- the only case in which an integer output can depend on the outcome -- it is counted here, so
   * |bins - reference bins| is bounded by these counters and nothing else. */
  AWX_CNT_MARG_AMBIG = 8,  /* arg-max of the mean member probabilities (AUROC positives, :414-416)  */
  AWX_CNT_EPRED_AMBIG = 9  /* arg-max of the fused probabilities (ECE accuracy term, :161-162)      */
};

int awx_bins_layout(int32_t num_classes, int32_t ece_bins, int32_t auroc_bins, AwxBinsLayout* out /*HOST*/);

/* logits_a/logits_b: fp32 [B,C,HW] contiguous (NCHW with HW = H*W).  labels: [B,HW] of
 * cfg->label_dtype, or NULL (then no bins are touched and `bins` may be NULL).
 * `bins` must be zeroed by the caller before the first call; calls accumulate. */
int awx_score(const float* logits_a, const float* logits_b, const void* labels,
              int64_t batch, int64_t pixels_per_image, const AwxScoreConfig* cfg /*HOST*/,
              int64_t* bins, const AwxScoreMaps* maps /*HOST, may be NULL*/, void* stream);

/* Confusion matrix from a prediction map (IoUMetrics.compute_iou / compute_pixel_accuracy with
 * [B,H,W] predictions, evaluation/metrics.py:53-71, 110-123).  pred: AWX_PRED_* dtype, labels:
 * AWX_LABEL_* dtype, n elements each.  The index is formed with torch's type promotion:
 * uint8*C wraps mod 256, and uint8 labels + uint8 predictions wrap as a whole.
 * confusion: int64 [C*C]; counters: int64 [AWX_NUM_COUNTERS] (AWX_CNT_VALID / _CORRECT / _BAD_LABEL / _PIXELS). */
int awx_confusion(const void* pred, int32_t pred_dtype, const void* labels, int32_t label_dtype, int64_t n,
                  int32_t num_classes, int32_t ignore_index, int64_t* confusion, int64_t* counters, void* stream);

/* Unbiased variance over the two members of the class probabilities, [B,C,HW]
 * (EnsembleDisagreementMetrics.compute_variance_map, evaluation/metrics.py:384-391). */
int awx_member_variance(const float* logits_a, const float* logits_b, float* out,
                        int64_t batch, int32_t num_classes, int64_t pixels_per_image, void* stream);

/* EnsembleDisagreementMetrics for a LIST of N >= 2 members (evaluation/metrics.py:336-438; two members run inside
 * awx_score): mi_out [B,HW] = H(mean_k p_k) - mean_k H(p_k); var_out [B,C,HW] = unbiased variance over members;
 * with labels: auroc_pos / auroc_neg int64 [auroc_bins] histograms of mi over [0, auroc_hi) for pixels whose
 * argmax(mean p) != / == label, counters int64 [AWX_NUM_COUNTERS] (AWX_CNT_VALID, AWX_CNT_ENS_WRONG, AWX_CNT_MARG_AMBIG,
 * AWX_CNT_PIXELS); all accumulate.
 * members: HOST array of n_members device pointers to fp32 [B,C,HW]; every output pointer may be NULL. */
int awx_members_n(const float* const* members /*HOST*/, int32_t n_members, const void* labels, int32_t label_dtype,
                  int64_t batch, int32_t num_classes, int64_t pixels_per_image, int32_t ignore_index,
                  int32_t auroc_bins, float auroc_hi, int64_t* auroc_pos, int64_t* auroc_neg, int64_t* counters,
                  float* mi_out, float* var_out, void* stream);

/* ------------------------------------------------------------------------------------
 * Weather corruption (WeatherDegradationTransforms, data/preprocessing.py:61-248).
 * Stochastic parameters are drawn on the host with the reference's RNG and passed in.
 * ---------------------------------------------------------------------------------- */

enum { AWX_CLEAN = 0, AWX_FOG = 1, AWX_RAIN = 2, AWX_SNOW = 3, AWX_NIGHT = 4 };
enum { AWX_F32 = 0, AWX_F64 = 1, AWX_BF16 = 2, AWX_U8 = 3 };

/* One per image (HOST array; awx_corrupt copies it into the head of its workspace).
 * Field/overlay offsets are in ELEMENTS of the arrays passed to awx_corrupt. */
typedef struct AwxCorruptParams {
  int32_t kind;        /* AWX_CLEAN..AWX_NIGHT                                              */
  int32_t blur_k;      /* rain: 3; snow: 3 or 7                                             */
  double d0;           /* fog: beta            night: intensity (noise*intensity*0.5, fp64) */
  double d1;           /* fog: airlight A rounded to fp32 (the reference's A*ones_like(fp32)) */
  float f0;            /* rain: fp32(1-haze)   snow: fp32(0.2*I)   night: fp32(1 - I*u)     */
  float f1;            /* rain: fp32(haze*0.7)                                              */
  float taps[4];       /* rain/snow: half Gaussian kernel, taps[0] = centre                 */
  int64_t field_offset;/* fog: depth[H*W]; night: noise[H*W*3]; element offset into `field` */
  int32_t item_begin;  /* rain: drops; snow: flakes -- range into `items`                   */
  int32_t item_count;
} AwxCorruptParams;

/* items: int32 [n,5] (device).  rain: x0,y0,x1,y1,thickness (cv2.line, preprocessing.py:160);
 * snow: x,y,radius,0,0 (cv2.circle filled, :194).  Scan conversion happens on the device and
 * reproduces OpenCV's footprints bit for bit (csrc/raster.cuh, tests/test_raster_cpu.py).
 * field (device): depth (fog, :227-248 output) or Gaussian noise (night, :222), fp32 or fp64.
 * workspace (device): awx_corrupt_workspace_bytes() bytes = params copy + 1 bit/pixel overlay
 * mask; always required; the call fills it itself.
 * img / out: uint8 [B,H,W,3] (device); `out` of an AWX_CLEAN image is a copy of `img`. */
size_t awx_corrupt_workspace_bytes(int64_t batch, int32_t height, int32_t width);

int awx_corrupt(const uint8_t* img, uint8_t* out, int64_t batch, int32_t height, int32_t width,
                const AwxCorruptParams* params /*HOST*/, const void* field, int32_t field_dtype,
                const int32_t* items, int64_t n_items, void* workspace, void* stream);

/* awx_corrupt with the dataset's Normalize(mean, std) + ToTensorV2 (data/loader.py:196-199) fused into the
 * corruption kernels' epilogue: norm_out [B,3,H,W] fp32 (AWX_F32) or bf16 (AWX_BF16) = (u8 - mean255[c]) * rdenom[c]
 * of the corrupted frame, the tensor the backbones consume, without a second pass over HBM.
 * `out` (uint8 HWC) may be NULL when only the normalised tensor is wanted.  mean255 / rdenom: HOST float[3]. */
int awx_corrupt_normalized(const uint8_t* img, uint8_t* out /*nullable*/, void* norm_out, int32_t norm_dtype,
                           const float* mean255 /*HOST*/, const float* rdenom /*HOST*/,
                           int64_t batch, int32_t height, int32_t width,
                           const AwxCorruptParams* params /*HOST*/, const void* field, int32_t field_dtype,
                           const int32_t* items, int64_t n_items, void* workspace, void* stream);

/* awx_corrupt followed by awx_score of the same batch (pixels_per_image = height * width) on one stream:
 * one call per weather condition of the evaluation sweep (scripts/evaluate.py:177-212 per batch). */
int awx_corrupt_score(const uint8_t* img, uint8_t* out, int32_t height, int32_t width,
                      const AwxCorruptParams* params /*HOST*/, const void* field, int32_t field_dtype,
                      const int32_t* items, int64_t n_items, void* workspace,
                      const float* logits_a, const float* logits_b, const void* labels, int64_t batch,
                      const AwxScoreConfig* cfg /*HOST*/, int64_t* bins, const AwxScoreMaps* maps /*HOST, may be NULL*/,
                      void* stream);

/* depth = max(gaussian_filter(ramp + noise, sigma, mode=reflect), 1) in fp64, axis 0 then axis 1,
 * with scipy's symmetric accumulation order (_generate_synthetic_depth, preprocessing.py:235-246).
 * noise: fp64 [B,H,W] drawn by the host; out: fp64 or fp32 [B,H,W]; tmp: fp64 [B,H,W] workspace;
 * weights: HOST fp64 [2*radius+1] normalised Gaussian taps (sigma=2, truncate=4 -> radius 8). */
int awx_synth_depth(const double* noise, void* out, int32_t out_dtype, double* tmp,
                    int64_t batch, int32_t height, int32_t width, double depth_scale,
                    const double* weights /*HOST*/, int32_t radius, void* stream);

/* ------------------------------------------------------------------------------------
 * Fog-density-aware loss, forward + unscaled gradients in one pass
 * (FogDensityAwareLoss.forward, models/model.py:560-617; focal :619-642).
 * sums[0] = sum_i w_i * loss_i, sums[1] = sum_i (depth_pred-depth_tgt)^2 (fp64, device).
 * dlogits (nullable) = w_i * dloss_i/dlogit / N ; ddepth (nullable) = 2*(pred-tgt)/N.
 * dfog (nullable) = s * base_loss_i / N, the gradient w.r.t. fog_density (model.py:593-597 path).
 * sums accumulates (zero it first); per-CTA partials go through `workspace`
 * (awx_fogloss_workspace_bytes() bytes) and are reduced in a fixed order: bit-reproducible.
 * bad_labels (device int64, nullable): count of labels outside [0,C) -- torch raises on those.
 * ---------------------------------------------------------------------------------- */
size_t awx_fogloss_workspace_bytes(void);

int awx_fogloss(const float* logits, const void* labels, int32_t label_dtype,
                const float* fog_density /*nullable*/, const float* depth_pred /*nullable*/,
                const float* depth_tgt /*nullable*/, float fog_sensitivity, int32_t focal,
                int64_t batch, int32_t num_classes, int64_t pixels_per_image,
                double* sums, float* dlogits, float* ddepth, float* dfog /*nullable*/,
                int64_t* bad_labels, void* workspace, void* stream);

/* EnsembleModel.forward's fused logits alone (models/model.py:443-462), without the statistics awx_score computes
 * next to them -- what a training step wants: fused = (w0*a + w1*b | (a+b)/2 | a if max softmax(a) > max softmax(b)
 * else b) [/ temperature].  Arithmetic as torch eager (the weighted sum rounds three times, max_confidence is
 * mask*l1 + (1-mask)*l2, the division is a true division): equal to awx_score's fused map bit for bit.
 * a, b, fused: device fp32 [batch, C, pixels_per_image] contiguous. */
int awx_fuse_forward(const float* logits_a, const float* logits_b, float* fused, int64_t batch, int32_t num_classes,
                     int64_t pixels_per_image, int32_t strategy, float w0, float w1, float temperature,
                     int32_t use_temperature, void* stream);

/* Gradient of EnsembleModel's fusion for training (models/model.py:442-462): with gs = grad_fused / T,
 * grad_a = gs*w0 | gs/2 | gs*[member a picked], grad_b likewise (either may be NULL), and
 * dots (device fp64 [3], overwritten) = {sum gs*a, sum gs*b, sum grad_fused*fused}, from which the caller forms the
 * gradients of the raw ensemble weights (softmax Jacobian) and of the temperature.
 * workspace: awx_fuse_backward_workspace_bytes() bytes. */
size_t awx_fuse_backward_workspace_bytes(void);
int awx_fuse_backward(const float* grad_fused, const float* logits_a, const float* logits_b, float* grad_a, float* grad_b,
                      int64_t batch, int32_t num_classes, int64_t pixels_per_image, int32_t strategy, float w0, float w1,
                      float temperature, int32_t use_temperature, double* dots, void* workspace, void* stream);

/* FogDensityAwareLoss._estimate_fog_density_from_depth (models/model.py:644-677) and its gradient:
 * density = clamp(0.7 * (d - min d)/(max d - min d + 1e-8) - 0.3 * [|grad d| > mean |grad d|], 0, 1), min / max /
 * mean over the whole [B,H,W] tensor.  bwd: grad_depth = d(sum grad_density * density)/d depth as autograd
 * derives it (min / max terms spread evenly over the attaining elements).  workspace:
 * awx_depth_density_workspace_bytes() bytes (need not persist between the two calls). */
size_t awx_depth_density_workspace_bytes(void);
int awx_depth_density_fwd(const float* depth, float* density, int64_t batch, int32_t height, int32_t width,
                          void* workspace, void* stream);
int awx_depth_density_bwd(const float* depth, const float* grad_density, float* grad_depth, int64_t batch,
                          int32_t height, int32_t width, void* workspace, void* stream);

/* x[i] *= *scale (device scalar) -- backward of a mean-reduced loss with grad_output != 1. */
int awx_scale_inplace(float* x, int64_t n, const float* scale, void* stream);

/* ------------------------------------------------------------------------------------
 * Callers and producers either side of the hot path.
 * ---------------------------------------------------------------------------------- */

/* albumentations Normalize(mean, std, max_pixel_value=255) + ToTensorV2 (data/loader.py:196-199):
 * uint8 [B,H,W,3] -> fp32 (AWX_F32) or bf16 (AWX_BF16) [B,3,H,W];
 * out = (x - mean255[c]) * rdenom[c] as two separately rounded fp32 operations.
 * mean255 / rdenom: HOST float[3] (mean*255 and 1/(std*255), formed by the caller in fp32). */
int awx_normalize_chw(const uint8_t* img, void* out, int32_t out_dtype, int64_t batch, int32_t height, int32_t width,
                      const float* mean255 /*HOST*/, const float* rdenom /*HOST*/, void* stream);

/* WeatherAugmentationPipeline._apply_style_transfer (data/loader.py:364-385):
 * out = cv2.convertScaleAbs(img, alpha, beta); if has_gain: out[...,2] = trunc(clip(out[...,2] * blue_gain, 0, 255)).
 * img / out: uint8, n_pixels * 3 bytes (HWC). */
int awx_style_transfer(const uint8_t* img, uint8_t* out, int64_t n_pixels, float alpha, float beta, double blue_gain,
                       int32_t has_gain, void* stream);

/* ConfidenceCalibration.optimize_temperature (evaluation/metrics.py:283-321), the whole grid in one pass:
 * sums[t] += sum over valid rows of cross_entropy(row / temperatures[t], label); sums[n] += valid rows;
 * sums[n+1] += labels outside [0,C) (torch raises on those).  rows are C consecutive floats of `logits`
 * (the reference's logits.view(-1, C)); labels: one per row; temperatures: HOST float[n_temps], all > 0.
 * sums: device fp64 [n_temps + 2], accumulates (zero it first).  workspace: awx_temperature_workspace_bytes(). */
size_t awx_temperature_workspace_bytes(int32_t n_temps);
int awx_temperature_nll(const float* logits, const void* labels, int32_t label_dtype, int64_t rows, int32_t num_classes,
                        int32_t ignore_index, const float* temperatures /*HOST*/, int32_t n_temps, double* sums,
                        void* workspace, void* stream);

/* WeatherDegradationTransforms.get_fog_density_map (data/preprocessing.py:250-288) in two calls around the
 * host's percentile lerp:
 *   awx_local_contrast: img [B,H,W,3] (AWX_U8, or AWX_F32 / AWX_F64 in [0,1], converted as (img*255).astype(uint8))
 *       -> contrast fp32 [B,H,W] = sqrt(box5((gray - box5(gray))^2)), plus, per image, the order statistics of
 *       ranks rank_lo and rank_hi (0-based, ascending) of its contrast values: order_stats fp32 [B,2].
 *   awx_fog_density_finish: out = clip((1 - contrast / denom[b]) * (0.3 + 0.7 * depth / max(depth_b)), 0, 1);
 *       denom: device fp32 [B] = max_contrast + 1e-8; depth / out: fp64 or fp32 [B,HW].
 * workspace: awx_fog_density_workspace_bytes(batch) bytes, the same buffer for both calls. */
size_t awx_fog_density_workspace_bytes(int64_t batch);
int awx_local_contrast(const void* img, int32_t img_dtype, float* contrast, int64_t batch, int32_t height, int32_t width,
                       int64_t rank_lo, int64_t rank_hi, float* order_stats, void* workspace, void* stream);
int awx_fog_density_finish(const float* contrast, const void* depth, int32_t depth_dtype, const float* denom, void* out,
                           int64_t batch, int64_t pixels_per_image, void* workspace, void* stream);

/* DepthEstimationPreprocessor._geometric_depth_estimation (data/preprocessing.py:332-367):
 * img uint8 [B,H,W,3] -> out fp64 [B,H,W]: perspective ramp with sky / road bands, minus 0.3 * |Laplacian(gray)| /
 * (max + 1e-8), clipped to [0,1], then scipy's gaussian_filter(sigma=2) (weights: HOST fp64 [2*radius+1]).
 * tmp: fp64 [B,H,W]; amax_workspace: int32 [B]. */
int awx_estimate_depth(const uint8_t* img, double* out, double* tmp, int64_t batch, int32_t height, int32_t width,
                       const double* weights /*HOST*/, int32_t radius, int32_t* amax_workspace, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AWX_H_ */
