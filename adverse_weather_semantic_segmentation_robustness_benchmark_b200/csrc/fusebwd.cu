// awx_fuse_backward: gradient of EnsembleModel's logit fusion (models/model.py:442-462) for training.
//   fused = u / T,  u = w0*a + w1*b (weighted_average) | (a+b)/2 (mean) | a if conf_a > conf_b else b (max_confidence)
// With gs = g / T:  ga = gs*w0 | gs/2 | gs*[pick a],  gb likewise, and three sums the host turns into the gradients
// of the raw ensemble weights (softmax Jacobian) and of the temperature:
//   dots[0] = sum gs*a   dots[1] = sum gs*b   dots[2] = sum gs*u  (= sum g*fused)
// A thread owns a pixel and walks its C class planes (coalesced across threads); sums go fp64 per thread -> warp ->
// CTA partial -> fixed-order reduction (bit-reproducible for a given device).
#include "awx_internal.cuh"

namespace awx {
namespace {

constexpr int kFThreads = 256;
constexpr int kFMaxBlocks = 2048;

__global__ void __launch_bounds__(kFThreads) fuse_backward_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                                                   const float* __restrict__ b, float* __restrict__ ga,
                                                                   float* __restrict__ gb, long long B, int C, long long HW,
                                                                   int strategy, float w0, float w1, float inv_t,
                                                                   double* __restrict__ part /*[grid][3]*/) {
  const long long total = B * HW;
  double da = 0.0, db = 0.0, du = 0.0;
  for (long long i = (long long)blockIdx.x * kFThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kFThreads) {
    const long long img = i / HW, px = i - img * HW;
    const size_t base = (size_t)img * C * HW + px;
    float ka = 0.5f, kb = 0.5f;
    if (strategy == AWX_FUSE_WEIGHTED) {
      ka = w0;
      kb = w1;
    } else if (strategy == AWX_FUSE_MAXCONF) {
      // member confidences = max softmax = 1 / sum exp(x - max); strict > picks a (model.py:449-455)
      float ma = -INFINITY, mb = -INFINITY;
      for (int c = 0; c < C; ++c) {
        ma = fmaxf(ma, a[base + (size_t)c * HW]);
        mb = fmaxf(mb, b[base + (size_t)c * HW]);
      }
      float sa = 0.f, sb = 0.f;
      for (int c = 0; c < C; ++c) {
        sa += ex2_approx((a[base + (size_t)c * HW] - ma) * kLog2e);
        sb += ex2_approx((b[base + (size_t)c * HW] - mb) * kLog2e);
      }
      const bool pick_a = __frcp_rn(sa) > __frcp_rn(sb);
      ka = pick_a ? 1.f : 0.f;
      kb = pick_a ? 0.f : 1.f;
    }
    for (int c = 0; c < C; ++c) {
      const size_t o = base + (size_t)c * HW;
      const float gs = g[o] * inv_t;
      const float va = a[o], vb = b[o];
      const float xa = gs * ka, xb = gs * kb;
      if (ga) ga[o] = xa;
      if (gb) gb[o] = xb;
      da += (double)gs * (double)va;
      db += (double)gs * (double)vb;
      du += (double)xa * (double)va + (double)xb * (double)vb;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    da += __shfl_down_sync(0xffffffffu, da, o);
    db += __shfl_down_sync(0xffffffffu, db, o);
    du += __shfl_down_sync(0xffffffffu, du, o);
  }
  __shared__ double s[3][kFThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s[0][threadIdx.x >> 5] = da;
    s[1][threadIdx.x >> 5] = db;
    s[2][threadIdx.x >> 5] = du;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < kFThreads / 32; ++w) t += s[threadIdx.x][w];
    part[3 * blockIdx.x + threadIdx.x] = t;
  }
}

__global__ void fuse_backward_finish_kernel(const double* __restrict__ part, int n, double* __restrict__ dots) {
  if (blockIdx.x == 0 && threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += part[3 * i + threadIdx.x];
    dots[threadIdx.x] = t;
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" size_t awx_fuse_backward_workspace_bytes(void) { return (size_t)kFMaxBlocks * 3 * sizeof(double); }

extern "C" int awx_fuse_backward(const float* grad_fused, const float* logits_a, const float* logits_b, float* grad_a,
                                 float* grad_b, int64_t batch, int32_t C, int64_t pixels_per_image, int32_t strategy, float w0,
                                 float w1, float temperature, int32_t use_temperature, double* dots, void* workspace,
                                 void* stream) {
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0 && C >= 1, AWX_E_ARG, "awx_fuse_backward: bad size");
  AWX_REQUIRE(strategy == AWX_FUSE_WEIGHTED || strategy == AWX_FUSE_MAXCONF || strategy == AWX_FUSE_MEAN, AWX_E_ARG,
              "awx_fuse_backward: unknown strategy %d", strategy);
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;
  AWX_REQUIRE(grad_fused && logits_a && logits_b && dots && workspace, AWX_E_ARG, "awx_fuse_backward: NULL pointer");
  const long long total = batch * pixels_per_image;
  long long blocks = (total + kFThreads - 1) / kFThreads;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > kFMaxBlocks) blocks = kFMaxBlocks;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* part = static_cast<double*>(workspace);
  const float inv_t = use_temperature ? 1.0f / temperature : 1.0f;
  fuse_backward_kernel<<<(unsigned)blocks, kFThreads, 0, s>>>(grad_fused, logits_a, logits_b, grad_a, grad_b, batch, C,
                                                              pixels_per_image, strategy, w0, w1, inv_t, part);
  fuse_backward_finish_kernel<<<1, 32, 0, s>>>(part, (int)blocks, dots);
  AWX_CUDA(cudaGetLastError());
  note_launch(2);
  return AWX_OK;
}
