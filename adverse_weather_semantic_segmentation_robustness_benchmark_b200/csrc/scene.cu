// Producers of the loss's fog_density / depth targets:
//   awx_local_contrast + awx_select_pair + awx_fog_density_finish
//        WeatherDegradationTransforms.get_fog_density_map          data/preprocessing.py:250-288
//   awx_estimate_depth
//        DepthEstimationPreprocessor._geometric_depth_estimation   data/preprocessing.py:332-367
//
// Third-party arithmetic restated here (and checked against the libraries by the golden vectors):
//   cv2.cvtColor(RGB2GRAY, uint8)  (R*9798 + G*19235 + B*3735 + 16384) >> 15
//   cv2.filter2D(fp32, ones(5,5)/25)  anchor centre, BORDER_REFLECT_101, one fused multiply-add per tap in
//        row-major tap order (OpenCV's vector body; its scalar tail for the last W mod 8 columns rounds the
//        product separately, so those columns can differ from a given OpenCV build by one ulp)
//   cv2.Laplacian(uint8 -> CV_64F, ksize=1)  [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101, exact integers
//   np.percentile  two adjacent order statistics selected on the device (3-level radix select on the fp32
//        bit patterns), NumPy's index / lerp scalars evaluated by the host
//   scipy.ndimage.gaussian_filter(sigma=2)  the fp64 'reflect' passes of corrupt.cu
#include "awx_internal.cuh"

namespace awx {
namespace {

constexpr int kT = 32;           // output tile edge
constexpr int kSceneThreads = 256;

__device__ __forceinline__ int reflect101s(int i, int n) {
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ unsigned gray_of(unsigned r, unsigned g, unsigned b) {
  return (r * 9798u + g * 19235u + b * 3735u + 16384u) >> 15;
}

// pixel -> uint8 RGB as the reference forms it: uint8 images as they are; float images in [0,1] through
// (image * 255).astype(np.uint8) (product in the image's own precision, truncation)
template <typename IT>
__device__ __forceinline__ unsigned gray_at(const IT* img, long long pix) {
  const IT* p = img + pix * 3;
  if constexpr (sizeof(IT) == 1) {
    return gray_of(p[0], p[1], p[2]);
  } else if constexpr (sizeof(IT) == 4) {
    return gray_of((unsigned)__float2int_rz(__fmul_rn(p[0], 255.0f)) & 0xffu, (unsigned)__float2int_rz(__fmul_rn(p[1], 255.0f)) & 0xffu,
                   (unsigned)__float2int_rz(__fmul_rn(p[2], 255.0f)) & 0xffu);
  } else {
    return gray_of((unsigned)__double2int_rz(__dmul_rn(p[0], 255.0)) & 0xffu, (unsigned)__double2int_rz(__dmul_rn(p[1], 255.0)) & 0xffu,
                   (unsigned)__double2int_rz(__dmul_rn(p[2], 255.0)) & 0xffu);
  }
}

// ------------------------------------------------------------------------- local contrast
// 32x32 output tile: gray (halo 4) -> 5x5 mean (halo 2) -> squared deviation -> 5x5 mean -> sqrt.
// The second filter reflects the SQUARED-DEVIATION image at the image border (not the gray image), so
// out-of-image halo cells of s_sq are copies of their reflected in-image cells.
template <typename IT>
__global__ void __launch_bounds__(kSceneThreads) contrast_kernel(const IT* __restrict__ img, float* __restrict__ contrast, int H, int W,
                                                                  unsigned* __restrict__ hist /*[B][4096]*/) {
  constexpr int G = kT + 8, S = kT + 4;
  __shared__ float s_g[G][G + 1];
  __shared__ float s_sq[S][S + 1];
  __shared__ unsigned s_hist[4096];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
  const IT* src = img + (size_t)b * H * W * 3;
  for (int i = threadIdx.x; i < 4096; i += kSceneThreads) s_hist[i] = 0u;
  for (int i = threadIdx.x; i < G * G; i += kSceneThreads) {
    const int ry = i / G, rx = i - ry * G;
    const int sy = reflect101s(y0 + ry - 4, H), sx = reflect101s(x0 + rx - 4, W);
    s_g[ry][rx] = __fdiv_rn((float)gray_at<IT>(src, (long long)sy * W + sx), 255.0f);
  }
  __syncthreads();
  const float k = __fdiv_rn(1.0f, 25.0f);  // np.ones((5,5), float32) / 25
  for (int i = threadIdx.x; i < S * S; i += kSceneThreads) {
    const int ry = i / S, rx = i - ry * S;
    const int gy = y0 + ry - 2, gx = x0 + rx - 2;
    float v = 0.f;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
      float acc = 0.f;
#pragma unroll
      for (int dy = 0; dy < 5; ++dy)
#pragma unroll
        for (int dx = 0; dx < 5; ++dx) acc = fmaf(s_g[ry + dy][rx + dx], k, acc);
      const float d = __fsub_rn(s_g[ry + 2][rx + 2], acc);
      v = __fmul_rn(d, d);
    }
    s_sq[ry][rx] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < S * S; i += kSceneThreads) {
    const int ry = i / S, rx = i - ry * S;
    const int gy = y0 + ry - 2, gx = x0 + rx - 2;
    if (gy >= 0 && gy < H && gx >= 0 && gx < W) continue;
    const int sy = reflect101s(gy, H) - (y0 - 2), sx = reflect101s(gx, W) - (x0 - 2);
    if (sy >= 0 && sy < S && sx >= 0 && sx < S) s_sq[ry][rx] = s_sq[sy][sx];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kT * kT; i += kSceneThreads) {
    const int ry = i / kT, rx = i - ry * kT;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy >= H || gx >= W) continue;
    float acc = 0.f;
#pragma unroll
    for (int dy = 0; dy < 5; ++dy)
#pragma unroll
      for (int dx = 0; dx < 5; ++dx) acc = fmaf(s_sq[ry + dy][rx + dx], k, acc);
    const float c = __fsqrt_rn(acc);
    contrast[((size_t)b * H + gy) * W + gx] = c;
    atomicAdd(&s_hist[__float_as_uint(c) >> 20], 1u);
  }
  __syncthreads();
  unsigned* gh = hist + (size_t)b * 4096;
  for (int i = threadIdx.x; i < 4096; i += kSceneThreads)
    if (s_hist[i]) atomicAdd(gh + i, s_hist[i]);
}

// ------------------------------------------------------------ two order statistics (radix select)
// state per image: for each of the two target ranks {remaining rank, prefix of the bits fixed so far}.
struct SelectState {
  unsigned long long rank[2];
  unsigned prefix[2];
};

// one block per image: locate the bucket of each target rank in hist[which][nbins]
__global__ void select_scan_kernel(const unsigned* __restrict__ hist, int nbins, int shared_hist, SelectState* __restrict__ state,
                                   int bits, float* __restrict__ out_values, int last) {
  const int b = blockIdx.x;
  SelectState st = state[b];
  if (threadIdx.x < 2) {
    const int j = threadIdx.x;
    const unsigned* h = hist + ((size_t)b * (shared_hist ? 1 : 2) + (shared_hist ? 0 : j)) * nbins;
    unsigned long long r = st.rank[j];
    int bin = 0;
    for (; bin < nbins - 1; ++bin) {
      const unsigned c = h[bin];
      if (r < c) break;
      r -= c;
    }
    st.rank[j] = r;
    st.prefix[j] = (st.prefix[j] << bits) | (unsigned)bin;
    state[b].rank[j] = st.rank[j];
    state[b].prefix[j] = st.prefix[j];
    if (last) out_values[2 * b + j] = __uint_as_float(st.prefix[j]);
  }
}

// histogram of the next `bits` bits of the elements whose higher bits match each target's prefix
__global__ void __launch_bounds__(kSceneThreads) select_hist_kernel(const float* __restrict__ v, long long n, const SelectState* __restrict__ state,
                                                                     int shift, int bits, unsigned* __restrict__ hist /*[B][2][1<<bits]*/) {
  extern __shared__ unsigned s_h[];  // [2][1<<bits]
  const int b = blockIdx.y;
  const int nb = 1 << bits;
  for (int i = threadIdx.x; i < 2 * nb; i += kSceneThreads) s_h[i] = 0u;
  __syncthreads();
  const unsigned p0 = state[b].prefix[0], p1 = state[b].prefix[1];
  const float* src = v + (size_t)b * n;
  for (long long i = (long long)blockIdx.x * kSceneThreads + threadIdx.x; i < n; i += (long long)gridDim.x * kSceneThreads) {
    const unsigned u = __float_as_uint(src[i]);
    const unsigned hi = u >> (shift + bits), lo = (u >> shift) & (unsigned)(nb - 1);
    if (hi == p0) atomicAdd(&s_h[lo], 1u);
    if (hi == p1) atomicAdd(&s_h[nb + lo], 1u);
  }
  __syncthreads();
  unsigned* gh = hist + (size_t)b * 2 * nb;
  for (int i = threadIdx.x; i < 2 * nb; i += kSceneThreads)
    if (s_h[i]) atomicAdd(gh + i, s_h[i]);
}

// ------------------------------------------------------------------------------ maxima
__device__ __forceinline__ unsigned long long order_key(double d) {
  const unsigned long long u = (unsigned long long)__double_as_longlong(d);
  return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double order_value(unsigned long long k) {
  const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
  return __longlong_as_double((long long)u);
}

template <typename DT>
__global__ void __launch_bounds__(256) max_kernel(const DT* __restrict__ v, long long n, unsigned long long* __restrict__ keys) {
  const int b = blockIdx.y;
  const DT* src = v + (size_t)b * n;
  unsigned long long best = 0ull;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const unsigned long long k = order_key((double)src[i]);
    best = k > best ? k : best;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
    best = other > best ? other : best;
  }
  if ((threadIdx.x & 31) == 0 && best) atomicMax(keys + b, best);
}

// fog = 1 - contrast / (max_contrast + 1e-8) in fp32; out = clip(fog * (0.3 + 0.7 * depth / max(depth)), 0, 1)
// in the depth's precision (fp64 for the synthetic depth)                     preprocessing.py:281-288
template <typename DT>
__global__ void __launch_bounds__(256) fog_finish_kernel(const float* __restrict__ contrast, const DT* __restrict__ depth,
                                                          const float* __restrict__ denom, const unsigned long long* __restrict__ dmax_keys,
                                                          DT* __restrict__ out, long long n) {
  const int b = blockIdx.y;
  const float den = denom[b];
  const double dmax = order_value(dmax_keys[b]);
  const size_t o = (size_t)b * n;
  for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
    const float fog = __fsub_rn(1.0f, __fdiv_rn(contrast[o + i], den));
    if constexpr (sizeof(DT) == 8) {
      const double nd = __ddiv_rn(depth[o + i], dmax);
      const double w = __dadd_rn(0.3, __dmul_rn(0.7, nd));
      out[o + i] = fmin(fmax(__dmul_rn((double)fog, w), 0.0), 1.0);
    } else {
      const float nd = __fdiv_rn(depth[o + i], (float)dmax);
      const float w = __fadd_rn(0.3f, __fmul_rn(0.7f, nd));
      out[o + i] = fminf(fmaxf(__fmul_rn(fog, w), 0.0f), 1.0f);
    }
  }
}

// ------------------------------------------------------------------------ depth estimation
// |Laplacian| of the gray image on a 32x32 tile (halo 1).  PASS 0: per-image maximum only.
// PASS 1: depth = clip(base(y) - 0.3 * |lap| / (max + 1e-8), 0, 1) in fp64, before the Gaussian.
template <int PASS>
__global__ void __launch_bounds__(kSceneThreads) laplace_kernel(const uint8_t* __restrict__ img, int H, int W, int* __restrict__ amax,
                                                                 double* __restrict__ out) {
  constexpr int G = kT + 2;
  __shared__ int s_g[G][G + 1];
  const int b = blockIdx.z;
  const int x0 = blockIdx.x * kT, y0 = blockIdx.y * kT;
  const uint8_t* src = img + (size_t)b * H * W * 3;
  for (int i = threadIdx.x; i < G * G; i += kSceneThreads) {
    const int ry = i / G, rx = i - ry * G;
    const int sy = reflect101s(y0 + ry - 1, H), sx = reflect101s(x0 + rx - 1, W);
    s_g[ry][rx] = (int)gray_at<uint8_t>(src, (long long)sy * W + sx);
  }
  __syncthreads();
  int best = 0;
  double denom = 1.0;
  if (PASS == 1) denom = __dadd_rn((double)amax[b], 1e-8);
  for (int i = threadIdx.x; i < kT * kT; i += kSceneThreads) {
    const int ry = i / kT, rx = i - ry * kT;
    const int gy = y0 + ry, gx = x0 + rx;
    if (gy >= H || gx >= W) continue;
    const int lap = s_g[ry][rx + 1] + s_g[ry + 2][rx + 1] + s_g[ry + 1][rx] + s_g[ry + 1][rx + 2] - 4 * s_g[ry + 1][rx + 1];
    const int a = lap < 0 ? -lap : lap;
    if (PASS == 0) {
      best = a > best ? a : best;
    } else {
      // base depth: y/h * 0.8 + 0.2; sky rows (< h//3) = 1; road rows (>= h//2) *= 0.5        :347-355
      double d = __dadd_rn(__dmul_rn(__ddiv_rn((double)gy, (double)H), 0.8), 0.2);
      if (gy < H / 3) d = 1.0;
      if (gy >= H / 2) d = __dmul_rn(d, 0.5);
      const double adj = __dmul_rn(-0.3, __ddiv_rn((double)a, denom));
      out[((size_t)b * H + gy) * W + gx] = fmin(fmax(__dadd_rn(d, adj), 0.0), 1.0);
    }
  }
  if (PASS == 0) {
    best = __reduce_max_sync(0xffffffffu, best);
    if ((threadIdx.x & 31) == 0 && best) atomicMax(amax + b, best);
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

/* workspace layout of the fog-density entry points, per image:
 *   [hist1 4096 u32][hist2 2*4096 u32][hist3 2*256 u32][SelectState][dmax key u64] */
namespace {
constexpr size_t kFogWsPerImage = (4096 + 2 * 4096 + 2 * 256) * sizeof(unsigned) + sizeof(SelectState) + sizeof(unsigned long long);
struct FogWs {
  unsigned* hist1;
  unsigned* hist2;
  unsigned* hist3;
  SelectState* state;
  unsigned long long* dmax;
};
FogWs fog_ws(void* workspace, int64_t B) {
  FogWs w;
  unsigned char* p = static_cast<unsigned char*>(workspace);
  w.hist1 = reinterpret_cast<unsigned*>(p);
  p += (size_t)B * 4096 * sizeof(unsigned);
  w.hist2 = reinterpret_cast<unsigned*>(p);
  p += (size_t)B * 2 * 4096 * sizeof(unsigned);
  w.hist3 = reinterpret_cast<unsigned*>(p);
  p += (size_t)B * 2 * 256 * sizeof(unsigned);
  w.state = reinterpret_cast<SelectState*>(p);
  p += (size_t)B * sizeof(SelectState);
  w.dmax = reinterpret_cast<unsigned long long*>(p);
  return w;
}
}  // namespace

extern "C" size_t awx_fog_density_workspace_bytes(int64_t batch) {
  if (batch <= 0) return 0;
  return ((size_t)batch * kFogWsPerImage + 255) & ~(size_t)255;
}

extern "C" int awx_local_contrast(const void* img, int32_t img_dtype, float* contrast, int64_t batch, int32_t H, int32_t W,
                                  int64_t rank_lo, int64_t rank_hi, float* order_stats, void* workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_local_contrast: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(img && contrast && order_stats && workspace, AWX_E_ARG, "awx_local_contrast: NULL pointer");
  AWX_REQUIRE(H >= 3 && W >= 3, AWX_E_UNSUPPORTED, "awx_local_contrast: images smaller than 3x3 are not supported");
  AWX_REQUIRE(batch <= 65535, AWX_E_UNSUPPORTED, "awx_local_contrast: batch %lld > 65535 per call", (long long)batch);
  const long long n = (long long)H * W;
  AWX_REQUIRE(rank_lo >= 0 && rank_lo <= rank_hi && rank_hi < n, AWX_E_ARG, "awx_local_contrast: ranks outside [0, H*W)");
  AWX_REQUIRE(img_dtype == AWX_U8 || img_dtype == AWX_F32 || img_dtype == AWX_F64, AWX_E_ARG, "awx_local_contrast: unknown image dtype");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FogWs w = fog_ws(workspace, batch);
  AWX_CUDA(cudaMemsetAsync(workspace, 0, (size_t)batch * kFogWsPerImage, s));
  // initial select state: ranks, empty prefixes
  {
    SelectState init;
    init.rank[0] = (unsigned long long)rank_lo;
    init.rank[1] = (unsigned long long)rank_hi;
    init.prefix[0] = init.prefix[1] = 0u;
    for (int64_t b = 0; b < batch; ++b) AWX_CUDA(cudaMemcpyAsync(w.state + b, &init, sizeof(init), cudaMemcpyHostToDevice, s));
  }
  dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, (unsigned)batch);
  if (img_dtype == AWX_U8)
    contrast_kernel<uint8_t><<<grid, kSceneThreads, 0, s>>>(static_cast<const uint8_t*>(img), contrast, H, W, w.hist1);
  else if (img_dtype == AWX_F32)
    contrast_kernel<float><<<grid, kSceneThreads, 0, s>>>(static_cast<const float*>(img), contrast, H, W, w.hist1);
  else
    contrast_kernel<double><<<grid, kSceneThreads, 0, s>>>(static_cast<const double*>(img), contrast, H, W, w.hist1);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  // 3-level radix select of the two ranks on the fp32 bit patterns: 12 + 12 + 8 bits
  select_scan_kernel<<<(unsigned)batch, 32, 0, s>>>(w.hist1, 4096, 1, w.state, 12, order_stats, 0);
  long long blocks = (n + kSceneThreads * 8 - 1) / (kSceneThreads * 8);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dim3 hgrid((unsigned)blocks, (unsigned)batch);
  select_hist_kernel<<<hgrid, kSceneThreads, 2 * 4096 * sizeof(unsigned), s>>>(contrast, n, w.state, 8, 12, w.hist2);
  select_scan_kernel<<<(unsigned)batch, 32, 0, s>>>(w.hist2, 4096, 0, w.state, 12, order_stats, 0);
  select_hist_kernel<<<hgrid, kSceneThreads, 2 * 256 * sizeof(unsigned), s>>>(contrast, n, w.state, 0, 8, w.hist3);
  select_scan_kernel<<<(unsigned)batch, 32, 0, s>>>(w.hist3, 256, 0, w.state, 8, order_stats, 1);
  AWX_CUDA(cudaGetLastError());
  note_launch(5);
  return AWX_OK;
}

extern "C" int awx_fog_density_finish(const float* contrast, const void* depth, int32_t depth_dtype, const float* denom, void* out,
                                      int64_t batch, int64_t pixels_per_image, void* workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0, AWX_E_ARG, "awx_fog_density_finish: negative size");
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;
  AWX_REQUIRE(contrast && depth && denom && out && workspace, AWX_E_ARG, "awx_fog_density_finish: NULL pointer");
  AWX_REQUIRE(depth_dtype == AWX_F32 || depth_dtype == AWX_F64, AWX_E_ARG, "awx_fog_density_finish: depth must be fp32 or fp64");
  AWX_REQUIRE(batch <= 65535, AWX_E_UNSUPPORTED, "awx_fog_density_finish: batch %lld > 65535 per call", (long long)batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FogWs w = fog_ws(workspace, batch);
  const long long n = pixels_per_image;
  long long blocks = (n + 256 * 4 - 1) / (256 * 4);
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  dim3 grid((unsigned)blocks, (unsigned)batch);
  AWX_CUDA(cudaMemsetAsync(w.dmax, 0, (size_t)batch * sizeof(unsigned long long), s));
  if (depth_dtype == AWX_F64) {
    max_kernel<double><<<grid, 256, 0, s>>>(static_cast<const double*>(depth), n, w.dmax);
    fog_finish_kernel<double><<<grid, 256, 0, s>>>(contrast, static_cast<const double*>(depth), denom, w.dmax, static_cast<double*>(out), n);
  } else {
    max_kernel<float><<<grid, 256, 0, s>>>(static_cast<const float*>(depth), n, w.dmax);
    fog_finish_kernel<float><<<grid, 256, 0, s>>>(contrast, static_cast<const float*>(depth), denom, w.dmax, static_cast<float*>(out), n);
  }
  AWX_CUDA(cudaGetLastError());
  note_launch(2);
  return AWX_OK;
}

extern "C" int awx_estimate_depth(const uint8_t* img, double* out, double* tmp, int64_t batch, int32_t H, int32_t W,
                                  const double* weights, int32_t radius, int32_t* amax_workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_estimate_depth: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(img && out && tmp && weights && amax_workspace, AWX_E_ARG, "awx_estimate_depth: NULL pointer");
  AWX_REQUIRE(radius >= 0 && radius <= 32, AWX_E_UNSUPPORTED, "awx_estimate_depth: radius %d outside 0..32", radius);
  AWX_REQUIRE(batch <= 65535, AWX_E_UNSUPPORTED, "awx_estimate_depth: batch %lld > 65535 per call", (long long)batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AWX_CUDA(cudaMemsetAsync(amax_workspace, 0, (size_t)batch * sizeof(int32_t), s));
  dim3 grid((W + kT - 1) / kT, (H + kT - 1) / kT, (unsigned)batch);
  laplace_kernel<0><<<grid, kSceneThreads, 0, s>>>(img, H, W, amax_workspace, nullptr);
  laplace_kernel<1><<<grid, kSceneThreads, 0, s>>>(img, H, W, amax_workspace, out);
  AWX_CUDA(cudaGetLastError());
  note_launch(2);
  // gaussian_filter(depth, sigma=2): out -> tmp (vertical) -> out (horizontal); no ramp, no floor
  return launch_gauss_f64(out, out, AWX_F64, tmp, batch, H, W, 0.0, -INFINITY, weights, radius, s);
}
