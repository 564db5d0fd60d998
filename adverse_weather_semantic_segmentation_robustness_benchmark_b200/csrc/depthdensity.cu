// FogDensityAwareLoss._estimate_fog_density_from_depth (models/model.py:644-677), forward and backward:
//   unit    = (d - min d) / (max d - min d + 1e-8)                       min / max over the WHOLE [B,H,W] tensor
//   gx, gy  = |forward differences| with the LAST column / row repeating the previous difference
//             (F.pad(..., mode='replicate') of the [W-1] / [H-1] wide difference maps)
//   mag     = sqrt(gx^2 + gy^2 + 1e-8)
//   density = clamp(0.7 * unit - 0.3 * [mag > mean(mag)], 0, 1)
// Backward (what autograd derives): the indicator has no gradient; with g' = g * 0.7 * [0 <= raw <= 1] / R,
//   d/dd_i = g'_i,   d/dmin = -sum g'_i (1 - unit_i),   d/dmax = -sum g'_i unit_i,
// the two scalars spread evenly over the elements that attain the minimum / maximum (torch's full-reduction
// min / max backward).  Three global scalars are needed before any pixel can be finished, so each direction is two
// passes with a fixed-order reduction in between (deterministic for a given device).
#include "awx_internal.cuh"

namespace awx {
namespace {

constexpr int kDThreads = 256;
constexpr int kDMaxBlocks = 2048;

struct DStats {      // device-side scalars shared by the passes
  float mn, mx, mean_mag, pad;
  double s_min, s_max;           // backward: sum g'(1-unit), sum g' unit
  unsigned long long n_min, n_max;
};

__device__ __forceinline__ float grad_mag(const float* __restrict__ d, int y, int x, int H, int W) {
  // difference maps are [.., W-1] and [.., H-1] wide; replicate padding repeats their last entry
  const int xx = x < W - 1 ? x : W - 2, yy = y < H - 1 ? y : H - 2;
  const float gx = W > 1 ? fabsf(__fsub_rn(d[(size_t)y * W + xx + 1], d[(size_t)y * W + xx])) : 0.f;
  const float gy = H > 1 ? fabsf(__fsub_rn(d[(size_t)(yy + 1) * W + x], d[(size_t)yy * W + x])) : 0.f;
  return __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(gx, gx), __fmul_rn(gy, gy)), 1e-8f));
}

__global__ void __launch_bounds__(kDThreads) dd_stats_kernel(const float* __restrict__ depth, long long B, int H, int W,
                                                              float* __restrict__ pmin, float* __restrict__ pmax,
                                                              double* __restrict__ psum) {
  const long long hw = (long long)H * W, total = B * hw;
  float mn = INFINITY, mx = -INFINITY;
  double sm = 0.0;
  for (long long i = (long long)blockIdx.x * kDThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kDThreads) {
    const long long img = i / hw, r = i - img * hw;
    const int y = (int)(r / W), x = (int)(r - (long long)y * W);
    const float* d = depth + img * hw;
    const float v = d[r];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
    sm += (double)grad_mag(d, y, x, H, W);
  }
  __shared__ float s_mn[kDThreads / 32], s_mx[kDThreads / 32];
  __shared__ double s_sm[kDThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    sm += __shfl_down_sync(0xffffffffu, sm, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_mn[threadIdx.x >> 5] = mn;
    s_mx[threadIdx.x >> 5] = mx;
    s_sm[threadIdx.x >> 5] = sm;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kDThreads / 32; ++w) {
      mn = fminf(mn, s_mn[w]);
      mx = fmaxf(mx, s_mx[w]);
      t += s_sm[w];
    }
    pmin[blockIdx.x] = mn;
    pmax[blockIdx.x] = mx;
    psum[blockIdx.x] = t;
  }
}

__global__ void dd_stats_finish_kernel(const float* __restrict__ pmin, const float* __restrict__ pmax,
                                       const double* __restrict__ psum, int n, double count, DStats* __restrict__ st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    float mn = INFINITY, mx = -INFINITY;
    double t = 0.0;
    for (int i = 0; i < n; ++i) {
      mn = fminf(mn, pmin[i]);
      mx = fmaxf(mx, pmax[i]);
      t += psum[i];
    }
    st->mn = mn;
    st->mx = mx;
    st->mean_mag = (float)(t / count);
    st->s_min = st->s_max = 0.0;
    st->n_min = st->n_max = 0ull;
  }
}

// raw = 0.7 * unit - 0.3 * [mag > mean]; returns unit through *unit_out
__device__ __forceinline__ float dd_raw(const float* __restrict__ d, long long r, int H, int W, const DStats& st, float* unit_out) {
  const int y = (int)(r / W), x = (int)(r - (long long)y * W);
  const float range = __fadd_rn(__fsub_rn(st.mx, st.mn), 1e-8f);
  const float unit = __fdiv_rn(__fsub_rn(d[r], st.mn), range);
  *unit_out = unit;
  const float edge = grad_mag(d, y, x, H, W) > st.mean_mag ? 0.3f : 0.0f;
  return __fsub_rn(__fmul_rn(unit, 0.7f), edge);
}

__global__ void __launch_bounds__(kDThreads) dd_forward_kernel(const float* __restrict__ depth, float* __restrict__ density,
                                                                long long B, int H, int W, const DStats* __restrict__ stp) {
  const DStats st = *stp;
  const long long hw = (long long)H * W, total = B * hw;
  for (long long i = (long long)blockIdx.x * kDThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kDThreads) {
    const long long img = i / hw, r = i - img * hw;
    float unit;
    const float raw = dd_raw(depth + img * hw, r, H, W, st, &unit);
    density[i] = fminf(fmaxf(raw, 0.f), 1.f);
  }
}

__global__ void __launch_bounds__(kDThreads) dd_backward1_kernel(const float* __restrict__ depth, const float* __restrict__ g,
                                                                  float* __restrict__ gd, long long B, int H, int W,
                                                                  DStats* __restrict__ stp, double* __restrict__ part /*[grid][2]*/,
                                                                  unsigned long long* __restrict__ cnt /*[grid][2]*/) {
  const DStats st = *stp;
  const long long hw = (long long)H * W, total = B * hw;
  const float range = __fadd_rn(__fsub_rn(st.mx, st.mn), 1e-8f);
  double s1 = 0.0, s2 = 0.0;
  unsigned long long nmin = 0, nmax = 0;
  for (long long i = (long long)blockIdx.x * kDThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kDThreads) {
    const long long img = i / hw, r = i - img * hw;
    const float* d = depth + img * hw;
    float unit;
    const float raw = dd_raw(d, r, H, W, st, &unit);
    const bool pass = raw >= 0.f && raw <= 1.f;   // clamp passes the gradient on the closed interval
    const float gp = pass ? __fdiv_rn(__fmul_rn(g[i], 0.7f), range) : 0.f;
    gd[i] = gp;
    s1 += (double)gp * (double)(1.0f - unit);
    s2 += (double)gp * (double)unit;
    nmin += d[r] == st.mn;
    nmax += d[r] == st.mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_down_sync(0xffffffffu, s1, o);
    s2 += __shfl_down_sync(0xffffffffu, s2, o);
    nmin += __shfl_down_sync(0xffffffffu, nmin, o);
    nmax += __shfl_down_sync(0xffffffffu, nmax, o);
  }
  __shared__ double s_a[kDThreads / 32], s_b[kDThreads / 32];
  __shared__ unsigned long long s_c[kDThreads / 32], s_d[kDThreads / 32];
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = s1;
    s_b[threadIdx.x >> 5] = s2;
    s_c[threadIdx.x >> 5] = nmin;
    s_d[threadIdx.x >> 5] = nmax;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    unsigned long long c = 0, e = 0;
    for (int w = 0; w < kDThreads / 32; ++w) {
      a += s_a[w];
      b += s_b[w];
      c += s_c[w];
      e += s_d[w];
    }
    part[2 * blockIdx.x] = a;
    part[2 * blockIdx.x + 1] = b;
    cnt[2 * blockIdx.x] = c;
    cnt[2 * blockIdx.x + 1] = e;
  }
}

__global__ void dd_backward_finish_kernel(const double* __restrict__ part, const unsigned long long* __restrict__ cnt, int n,
                                          DStats* __restrict__ st) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, b = 0.0;
    unsigned long long c = 0, e = 0;
    for (int i = 0; i < n; ++i) {
      a += part[2 * i];
      b += part[2 * i + 1];
      c += cnt[2 * i];
      e += cnt[2 * i + 1];
    }
    st->s_min = a;
    st->s_max = b;
    st->n_min = c;
    st->n_max = e;
  }
}

__global__ void __launch_bounds__(kDThreads) dd_backward2_kernel(const float* __restrict__ depth, float* __restrict__ gd, long long total,
                                                                  const DStats* __restrict__ stp) {
  const DStats st = *stp;
  const float to_min = st.n_min ? (float)(-st.s_min / (double)st.n_min) : 0.f;
  const float to_max = st.n_max ? (float)(-st.s_max / (double)st.n_max) : 0.f;
  for (long long i = (long long)blockIdx.x * kDThreads + threadIdx.x; i < total; i += (long long)gridDim.x * kDThreads) {
    const float v = depth[i];
    float add = 0.f;
    if (v == st.mn) add += to_min;
    if (v == st.mx) add += to_max;
    if (add != 0.f) gd[i] += add;
  }
}

struct DWs {
  DStats* st;
  float* pmin;
  float* pmax;
  double* psum;   // also backward partials [grid][2]
  unsigned long long* cnt;
};
DWs dd_ws(void* workspace) {
  DWs w;
  unsigned char* p = static_cast<unsigned char*>(workspace);
  w.st = reinterpret_cast<DStats*>(p);
  p += 256;
  w.pmin = reinterpret_cast<float*>(p);
  p += kDMaxBlocks * sizeof(float);
  w.pmax = reinterpret_cast<float*>(p);
  p += kDMaxBlocks * sizeof(float);
  w.psum = reinterpret_cast<double*>(p);
  p += 2 * kDMaxBlocks * sizeof(double);
  w.cnt = reinterpret_cast<unsigned long long*>(p);
  return w;
}
long long dd_blocks(long long total) {
  long long blocks = (total + kDThreads - 1) / kDThreads;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks > kDMaxBlocks) blocks = kDMaxBlocks;
  return blocks < 1 ? 1 : blocks;
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" size_t awx_depth_density_workspace_bytes(void) {
  return 256 + 2 * kDMaxBlocks * sizeof(float) + 2 * kDMaxBlocks * sizeof(double) + 2 * kDMaxBlocks * sizeof(unsigned long long);
}

extern "C" int awx_depth_density_fwd(const float* depth, float* density, int64_t batch, int32_t H, int32_t W, void* workspace,
                                     void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_depth_density_fwd: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(depth && density && workspace, AWX_E_ARG, "awx_depth_density_fwd: NULL pointer");
  AWX_REQUIRE(H >= 2 && W >= 2, AWX_E_UNSUPPORTED, "awx_depth_density_fwd: needs at least 2x2 maps (the reference pads empty difference maps)");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DWs w = dd_ws(workspace);
  const long long total = batch * (long long)H * W;
  const long long blocks = dd_blocks(total);
  dd_stats_kernel<<<(unsigned)blocks, kDThreads, 0, s>>>(depth, batch, H, W, w.pmin, w.pmax, w.psum);
  dd_stats_finish_kernel<<<1, 32, 0, s>>>(w.pmin, w.pmax, w.psum, (int)blocks, (double)total, w.st);
  dd_forward_kernel<<<(unsigned)blocks, kDThreads, 0, s>>>(depth, density, batch, H, W, w.st);
  AWX_CUDA(cudaGetLastError());
  note_launch(3);
  return AWX_OK;
}

extern "C" int awx_depth_density_bwd(const float* depth, const float* grad_density, float* grad_depth, int64_t batch, int32_t H,
                                     int32_t W, void* workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_depth_density_bwd: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(depth && grad_density && grad_depth && workspace, AWX_E_ARG, "awx_depth_density_bwd: NULL pointer");
  AWX_REQUIRE(H >= 2 && W >= 2, AWX_E_UNSUPPORTED, "awx_depth_density_bwd: needs at least 2x2 maps");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  DWs w = dd_ws(workspace);
  const long long total = batch * (long long)H * W;
  const long long blocks = dd_blocks(total);
  // the statistics are recomputed: the workspace need not survive between the forward and the backward call
  dd_stats_kernel<<<(unsigned)blocks, kDThreads, 0, s>>>(depth, batch, H, W, w.pmin, w.pmax, w.psum);
  dd_stats_finish_kernel<<<1, 32, 0, s>>>(w.pmin, w.pmax, w.psum, (int)blocks, (double)total, w.st);
  dd_backward1_kernel<<<(unsigned)blocks, kDThreads, 0, s>>>(depth, grad_density, grad_depth, batch, H, W, w.st, w.psum, w.cnt);
  dd_backward_finish_kernel<<<1, 32, 0, s>>>(w.psum, w.cnt, (int)blocks, w.st);
  dd_backward2_kernel<<<(unsigned)blocks, kDThreads, 0, s>>>(depth, grad_depth, total, w.st);
  AWX_CUDA(cudaGetLastError());
  note_launch(5);
  return AWX_OK;
}
