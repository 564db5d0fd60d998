// libawx: version, error reporting, device queries.
#include <atomic>

#include "awx_internal.cuh"

namespace awx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* where) {
  set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
  return (int)e;
}

static std::atomic<long long> g_launches{0};
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 148;
    cached = prop.multiProcessorCount;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace awx

extern "C" int64_t awx_launch_count(void) { return awx::launches(); }

extern "C" int awx_version(void) { return AWX_VERSION; }

extern "C" const char* awx_last_error(void) { return awx::g_err; }

extern "C" int awx_device_info(int* sms, int* major, int* minor) {
  int dev = 0;
  AWX_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  AWX_CUDA(cudaGetDeviceProperties(&prop, dev));
  if (sms) *sms = prop.multiProcessorCount;
  if (major) *major = prop.major;
  if (minor) *minor = prop.minor;
  return AWX_OK;
}

/* One call per condition: the corruption kernels of a batch and the fused scoring pass of the logits the
 * caller's model produced, enqueued back to back on `stream` (SURVEY.md 8b "awx_corrupt_score").  The two stages
 * share no per-pixel data (the corrupted frames feed the backbone, the score reads its logits), so a single
 * kernel would only interleave two independent streams; see DESIGN.md section 4. */
extern "C" int awx_corrupt_score(const uint8_t* img, uint8_t* out, int32_t H, int32_t W, const AwxCorruptParams* params,
                                 const void* field, int32_t field_dtype, const int32_t* items, int64_t n_items,
                                 void* workspace, const float* logits_a, const float* logits_b, const void* labels,
                                 int64_t batch, const AwxScoreConfig* cfg, int64_t* bins, const AwxScoreMaps* maps,
                                 void* stream) {
  int rc = awx_corrupt(img, out, batch, H, W, params, field, field_dtype, items, n_items, workspace, stream);
  if (rc != AWX_OK) return rc;
  return awx_score(logits_a, logits_b, labels, batch, (int64_t)H * W, cfg, bins, maps, stream);
}
