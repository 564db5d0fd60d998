// awx_fuse_forward / awx_fuse_backward: EnsembleModel's logit fusion (models/model.py:442-462) for training --
// the fused logits alone (awx_score computes them next to every statistic; a training step only wants the tensor)
// and their gradient.
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

// max_confidence: member confidences = max softmax = 1 / sum exp(x - max); strict > picks a (model.py:449-455).
// One definition for the forward and the backward kernel, so that both see the same winner.
__device__ __forceinline__ bool pick_member_a(const float* __restrict__ a, const float* __restrict__ b, size_t base, int C,
                                              long long HW) {
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
  return __frcp_rn(sa) > __frcp_rn(sb);
}

// weighted_average / mean: purely element-wise, torch eager's roundings (w0*a, w1*b and the sum round separately;
// (a+b)*0.5; then a true division by T).  16-byte vectors, grid-stride.
template <bool MEAN, bool DIV>
__device__ __forceinline__ float fuse_elem(float x, float y, float w0, float w1, float T) {
  const float u = MEAN ? __fmul_rn(__fadd_rn(x, y), 0.5f) : __fadd_rn(__fmul_rn(w0, x), __fmul_rn(w1, y));
  return DIV ? __fdiv_rn(u, T) : u;
}
template <bool MEAN, bool DIV>
__global__ void __launch_bounds__(kFThreads) fuse_forward_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                                  float* __restrict__ out, long long n, float w0, float w1,
                                                                  float T, int vec) {
  const long long stride = (long long)gridDim.x * kFThreads, t0 = (long long)blockIdx.x * kFThreads + threadIdx.x;
  if (vec) {
    const long long n4 = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    float4* o4 = reinterpret_cast<float4*>(out);
    // four independent 16-byte loads per member in flight per thread
    constexpr int U = 4;
    long long i = t0;
    for (; i + (U - 1) * stride < n4; i += U * stride) {
      float4 x[U], y[U];
#pragma unroll
      for (int k = 0; k < U; ++k) {
        x[k] = __ldcs(a4 + i + k * stride);
        y[k] = __ldcs(b4 + i + k * stride);
      }
#pragma unroll
      for (int k = 0; k < U; ++k) {
        float4 r;
        r.x = fuse_elem<MEAN, DIV>(x[k].x, y[k].x, w0, w1, T);
        r.y = fuse_elem<MEAN, DIV>(x[k].y, y[k].y, w0, w1, T);
        r.z = fuse_elem<MEAN, DIV>(x[k].z, y[k].z, w0, w1, T);
        r.w = fuse_elem<MEAN, DIV>(x[k].w, y[k].w, w0, w1, T);
        __stcs(o4 + i + k * stride, r);
      }
    }
    for (; i < n4; i += stride) {
      const float4 x = __ldcs(a4 + i), y = __ldcs(b4 + i);
      float4 r;
      r.x = fuse_elem<MEAN, DIV>(x.x, y.x, w0, w1, T);
      r.y = fuse_elem<MEAN, DIV>(x.y, y.y, w0, w1, T);
      r.z = fuse_elem<MEAN, DIV>(x.z, y.z, w0, w1, T);
      r.w = fuse_elem<MEAN, DIV>(x.w, y.w, w0, w1, T);
      __stcs(o4 + i, r);
    }
    for (long long i = (n4 << 2) + t0; i < n; i += stride) out[i] = fuse_elem<MEAN, DIV>(a[i], b[i], w0, w1, T);
  } else {
    for (long long i = t0; i < n; i += stride) out[i] = fuse_elem<MEAN, DIV>(a[i], b[i], w0, w1, T);
  }
}

// max_confidence: a thread owns a pixel; mask*l1 + (1-mask)*l2 as written (1*x and 0*y round separately: an
// infinite logit of the other member gives NaN, as in the reference), then the division.
__global__ void __launch_bounds__(kFThreads) fuse_forward_maxconf_kernel(const float* __restrict__ a,
                                                                          const float* __restrict__ b,
                                                                          float* __restrict__ out, long long B, int C,
                                                                          long long HW, float T, int use_t) {
  const long long total = B * HW;
  for (long long i = (long long)blockIdx.x * kFThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kFThreads) {
    const long long img = i / HW, px = i - img * HW;
    const size_t base = (size_t)img * C * HW + px;
    const bool pick_a = pick_member_a(a, b, base, C, HW);
    const float ka = pick_a ? 1.f : 0.f, kb = pick_a ? 0.f : 1.f;
    for (int c = 0; c < C; ++c) {
      const size_t o = base + (size_t)c * HW;
      const float u = __fadd_rn(__fmul_rn(ka, a[o]), __fmul_rn(kb, b[o]));
      out[o] = use_t ? __fdiv_rn(u, T) : u;
    }
  }
}

// fp64 partial sums: thread -> warp (shuffles) -> CTA (shared memory) -> part[3 * blockIdx.x + k]
__device__ __forceinline__ void block_reduce3(double da, double db, double du, double* __restrict__ part) {
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

// weighted_average / mean: the gradient is element-wise too (constant ka, kb): 16-byte vectors over the flat
// tensors, two vectors of each input in flight per thread; the three dot products ride along in fp64.
__global__ void __launch_bounds__(kFThreads) fuse_backward_flat_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                                                        const float* __restrict__ b, float* __restrict__ ga,
                                                                        float* __restrict__ gb, long long n, float ka, float kb,
                                                                        float inv_t, double* __restrict__ part, int vec) {
  const long long stride = (long long)gridDim.x * kFThreads, t0 = (long long)blockIdx.x * kFThreads + threadIdx.x;
  double da = 0.0, db = 0.0, du = 0.0;
  auto elem = [&](float gv, float va, float vb, float& xa, float& xb) {
    const float gs = gv * inv_t;
    xa = gs * ka;
    xb = gs * kb;
    da += (double)gs * (double)va;
    db += (double)gs * (double)vb;
    du += (double)xa * (double)va + (double)xb * (double)vb;
  };
  auto vec4 = [&](const float4& gv, const float4& va, const float4& vb, long long i) {
    float4 xa, xb;
    elem(gv.x, va.x, vb.x, xa.x, xb.x);
    elem(gv.y, va.y, vb.y, xa.y, xb.y);
    elem(gv.z, va.z, vb.z, xa.z, xb.z);
    elem(gv.w, va.w, vb.w, xa.w, xb.w);
    if (ga) __stcs(reinterpret_cast<float4*>(ga) + i, xa);
    if (gb) __stcs(reinterpret_cast<float4*>(gb) + i, xb);
  };
  long long tail0 = 0;
  if (vec) {
    const long long n4 = n >> 2;
    const float4* g4 = reinterpret_cast<const float4*>(g);
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const float4* b4 = reinterpret_cast<const float4*>(b);
    long long i = t0;
    for (; i + stride < n4; i += 2 * stride) {
      const float4 g0 = __ldcs(g4 + i), a0 = __ldcs(a4 + i), b0 = __ldcs(b4 + i);
      const float4 g1 = __ldcs(g4 + i + stride), a1 = __ldcs(a4 + i + stride), b1 = __ldcs(b4 + i + stride);
      vec4(g0, a0, b0, i);
      vec4(g1, a1, b1, i + stride);
    }
    for (; i < n4; i += stride) vec4(__ldcs(g4 + i), __ldcs(a4 + i), __ldcs(b4 + i), i);
    tail0 = n4 << 2;
  }
  for (long long i = tail0 + t0; i < n; i += stride) {
    float xa, xb;
    elem(g[i], a[i], b[i], xa, xb);
    if (ga) ga[i] = xa;
    if (gb) gb[i] = xb;
  }
  block_reduce3(da, db, du, part);
}

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
      const bool pick_a = pick_member_a(a, b, base, C, HW);
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
  block_reduce3(da, db, du, part);
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

extern "C" int awx_fuse_forward(const float* logits_a, const float* logits_b, float* fused, int64_t batch, int32_t C,
                                int64_t pixels_per_image, int32_t strategy, float w0, float w1, float temperature,
                                int32_t use_temperature, void* stream) {
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0 && C >= 1, AWX_E_ARG, "awx_fuse_forward: bad size");
  AWX_REQUIRE(strategy == AWX_FUSE_WEIGHTED || strategy == AWX_FUSE_MAXCONF || strategy == AWX_FUSE_MEAN, AWX_E_ARG,
              "awx_fuse_forward: unknown strategy %d", strategy);
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;
  AWX_REQUIRE(logits_a && logits_b && fused, AWX_E_ARG, "awx_fuse_forward: NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const long long cap = (long long)sm_count() * 8;
  if (strategy == AWX_FUSE_MAXCONF) {
    const long long total = batch * pixels_per_image;
    long long blocks = (total + kFThreads - 1) / kFThreads;
    if (blocks > cap) blocks = cap;
    fuse_forward_maxconf_kernel<<<(unsigned)blocks, kFThreads, 0, s>>>(logits_a, logits_b, fused, batch, C, pixels_per_image,
                                                                      temperature, use_temperature);
  } else {
    const long long n = batch * C * pixels_per_image;
    const int vec = (((uintptr_t)logits_a | (uintptr_t)logits_b | (uintptr_t)fused) & 15) == 0;
    long long blocks = ((vec ? (n + 3) / 4 : n) + kFThreads - 1) / kFThreads;
    if (blocks > cap) blocks = cap;
    const bool mean = strategy == AWX_FUSE_MEAN;
    auto kern = mean ? (use_temperature ? fuse_forward_kernel<true, true> : fuse_forward_kernel<true, false>)
                     : (use_temperature ? fuse_forward_kernel<false, true> : fuse_forward_kernel<false, false>);
    kern<<<(unsigned)blocks, kFThreads, 0, s>>>(logits_a, logits_b, fused, n, w0, w1, temperature, vec);
  }
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

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
  if (strategy == AWX_FUSE_MAXCONF) {
    fuse_backward_kernel<<<(unsigned)blocks, kFThreads, 0, s>>>(grad_fused, logits_a, logits_b, grad_a, grad_b, batch, C,
                                                                pixels_per_image, strategy, w0, w1, inv_t, part);
  } else {
    const long long n = total * C;
    const int vec = (((uintptr_t)grad_fused | (uintptr_t)logits_a | (uintptr_t)logits_b | (uintptr_t)grad_a |
                      (uintptr_t)grad_b) & 15) == 0;
    const bool mean = strategy == AWX_FUSE_MEAN;
    fuse_backward_flat_kernel<<<(unsigned)blocks, kFThreads, 0, s>>>(grad_fused, logits_a, logits_b, grad_a, grad_b, n,
                                                                     mean ? 0.5f : w0, mean ? 0.5f : w1, inv_t, part, vec);
  }
  fuse_backward_finish_kernel<<<1, 32, 0, s>>>(part, (int)blocks, dots);
  AWX_CUDA(cudaGetLastError());
  note_launch(2);
  return AWX_OK;
}
