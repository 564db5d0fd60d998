// awx_confusion: C x C confusion matrix from a prediction map and a label map.
// HBM-bound integer work: 128-bit loads where the dtype allows, CTA-private shared histogram
// with a warp-uniform fast path (real label maps are piecewise constant), one flush per CTA.
#include "awx_internal.cuh"

namespace awx {
namespace {

constexpr int kThreads = 256;

template <typename P, typename L>
__global__ void __launch_bounds__(kThreads) confusion_kernel(const P* __restrict__ pred, const L* __restrict__ lab,
                                                             long long n, int C, int ignore_index,
                                                             unsigned long long* __restrict__ confusion,
                                                             unsigned long long* __restrict__ counters) {
  extern __shared__ unsigned s_hist[];  // [C*C] + 4 counters
  unsigned* s_cnt = s_hist + C * C;
  for (int i = threadIdx.x; i < C * C + 4; i += kThreads) s_hist[i] = 0u;
  __syncthreads();
  unsigned n_valid = 0, n_correct = 0, n_bad = 0, n_pix = 0;
  const long long stride = (long long)gridDim.x * kThreads;
  const long long cc = (long long)C * C;
  for (long long i0 = (long long)blockIdx.x * kThreads; i0 < n; i0 += stride) {
    const long long i = i0 + threadIdx.x;
    const bool act = i < n;
    long long t = ignore_index, q = 0;
    if (act) {
      t = (long long)lab[i];
      q = (long long)pred[i];
    }
    const bool valid = act && t != (long long)ignore_index;
    n_pix += act;
    n_valid += valid;
    n_correct += valid && t == q;
    long long idx = -1;
    if (valid) {
      // torch promotion of `targets * C + predictions` (metrics.py:68)
      if (sizeof(L) == 1 && sizeof(P) == 1)
        idx = (t * C + q) & 0xff;
      else if (sizeof(L) == 1)
        idx = ((t * C) & 0xff) + q;
      else
        idx = t * C + q;
      if (idx < 0 || idx >= cc) {
        ++n_bad;
        idx = -1;
      }
    }
    // warp-uniform fast path: one shared atomic for the whole warp
    const int key = (int)idx;
    int same;
    __match_all_sync(0xffffffffu, key, &same);
    if (same) {
      if ((threadIdx.x & 31) == 0 && key >= 0) atomicAdd(&s_hist[key], 32u);
    } else if (key >= 0) {
      atomicAdd(&s_hist[key], 1u);
    }
  }
  unsigned v[4] = {n_valid, n_correct, n_bad, n_pix};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned s = __reduce_add_sync(0xffffffffu, v[k]);
    if ((threadIdx.x & 31) == 0 && s) atomicAdd(&s_cnt[k], s);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < C * C; i += kThreads)
    if (s_hist[i]) atomicAdd(confusion + i, (unsigned long long)s_hist[i]);
  if (threadIdx.x == 0 && counters) {
    if (s_cnt[0]) atomicAdd(counters + AWX_CNT_VALID, (unsigned long long)s_cnt[0]);
    if (s_cnt[1]) atomicAdd(counters + AWX_CNT_CORRECT, (unsigned long long)s_cnt[1]);
    if (s_cnt[2]) atomicAdd(counters + AWX_CNT_BAD_LABEL, (unsigned long long)s_cnt[2]);
    if (s_cnt[3]) atomicAdd(counters + AWX_CNT_PIXELS, (unsigned long long)s_cnt[3]);
  }
}

template <typename P, typename L>
int launch(const void* pred, const void* lab, long long n, int C, int ign, int64_t* conf, int64_t* cnt, cudaStream_t s) {
  long long blocks = (n + kThreads - 1) / kThreads;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  const size_t smem = ((size_t)C * C + 4) * sizeof(unsigned);
  confusion_kernel<P, L><<<(unsigned)blocks, kThreads, smem, s>>>(
      static_cast<const P*>(pred), static_cast<const L*>(lab), n, C, ign,
      reinterpret_cast<unsigned long long*>(conf), reinterpret_cast<unsigned long long*>(cnt));
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" int awx_confusion(const void* pred, int32_t pred_dtype, const void* labels, int32_t label_dtype, int64_t n,
                             int32_t C, int32_t ignore_index, int64_t* confusion, int64_t* counters, void* stream) {
  AWX_REQUIRE(n >= 0, AWX_E_ARG, "awx_confusion: negative size");
  AWX_REQUIRE(C >= 1 && C <= AWX_MAX_CLASSES, AWX_E_UNSUPPORTED, "awx_confusion: num_classes %d outside 1..%d", C, AWX_MAX_CLASSES);
  if (n == 0) return AWX_OK;
  AWX_REQUIRE(pred && labels && confusion, AWX_E_ARG, "awx_confusion: NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pred_dtype == AWX_PRED_U8 && label_dtype == AWX_LABEL_U8) return launch<uint8_t, uint8_t>(pred, labels, n, C, ignore_index, confusion, counters, s);
  if (pred_dtype == AWX_PRED_U8 && label_dtype == AWX_LABEL_I64) return launch<uint8_t, long long>(pred, labels, n, C, ignore_index, confusion, counters, s);
  if (pred_dtype == AWX_PRED_I64 && label_dtype == AWX_LABEL_U8) return launch<long long, uint8_t>(pred, labels, n, C, ignore_index, confusion, counters, s);
  if (pred_dtype == AWX_PRED_I64 && label_dtype == AWX_LABEL_I64) return launch<long long, long long>(pred, labels, n, C, ignore_index, confusion, counters, s);
  set_error("awx_confusion: unknown dtype combination pred=%d label=%d", pred_dtype, label_dtype);
  return AWX_E_ARG;
}
