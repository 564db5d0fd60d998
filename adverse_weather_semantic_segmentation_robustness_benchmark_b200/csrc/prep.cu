// Producers / consumers either side of the hot path:
//   awx_normalize_chw     albumentations Normalize + ToTensorV2 (data/loader.py:196-199): uint8 HWC ->
//                         fp32 / bf16 CHW, the tensor the backbones consume
//   awx_style_transfer    cv2.convertScaleAbs + blue-channel gain (data/loader.py:364-385)
//   awx_temperature_nll   the 100-point NLL grid of ConfidenceCalibration.optimize_temperature
//                         (evaluation/metrics.py:283-321) in ONE pass over the logits
#include <cuda_bf16.h>

#include "awx_internal.cuh"

namespace awx {
namespace {

// --------------------------------------------------------------------- Normalize + CHW
// 1024-pixel chunks: 3072 bytes in as 192 x 16 B, staged in shared memory; thread t then owns pixels
// 4t..4t+3 (words 3t..3t+2: conflict free) and writes one 16-byte (fp32) or 8-byte (bf16) vector per
// channel plane.  Algorithmic bytes per pixel: 3 + 12 (fp32) or 3 + 6 (bf16).
constexpr int kNormChunk = 1024;
constexpr int kNormThreads = 256;

template <bool BF16>
__global__ void __launch_bounds__(kNormThreads) normalize_chw_kernel(const uint8_t* __restrict__ img, void* __restrict__ out,
                                                                      long long HW, float m0, float m1, float m2,
                                                                      float r0, float r1, float r2) {
  __shared__ __align__(16) unsigned s_in[kNormChunk * 3 / 4];
  const int b = blockIdx.y;
  const uint8_t* src = img + (size_t)b * HW * 3;
  const bool aligned = (((uintptr_t)src) & 15) == 0 && (HW & 3) == 0;
  const long long nchunks = (HW + kNormChunk - 1) / kNormChunk;
  const float mean[3] = {m0, m1, m2}, rden[3] = {r0, r1, r2};
  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long px0 = ch * kNormChunk;
    const int npx = (int)((HW - px0) < kNormChunk ? (HW - px0) : kNormChunk);
    __syncthreads();
    if (aligned && npx == kNormChunk) {
      if (threadIdx.x < kNormChunk * 3 / 16)
        reinterpret_cast<uint4*>(s_in)[threadIdx.x] = ld_stream_u4(src + px0 * 3 + threadIdx.x * 16);
    } else {
      uint8_t* sb = reinterpret_cast<uint8_t*>(s_in);
      for (int i = threadIdx.x; i < npx * 3; i += kNormThreads) sb[i] = src[px0 * 3 + i];
    }
    __syncthreads();
    const int p = threadIdx.x * 4;
    if (p >= npx) continue;
    const unsigned wv[3] = {s_in[threadIdx.x * 3], s_in[threadIdx.x * 3 + 1], s_in[threadIdx.x * 3 + 2]};
    float v[3][4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int k = j * 3 + c;
        const float x = (float)((wv[k >> 2] >> ((k & 3) * 8)) & 0xffu);
        // img -= mean*255 ; img *= 1/(std*255): two separately rounded fp32 operations
        v[c][j] = __fmul_rn(__fsub_rn(x, mean[c]), rden[c]);
      }
    const bool full4 = aligned && p + 3 < npx;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const size_t o = ((size_t)b * 3 + c) * HW + px0 + p;
      if (BF16) {
        __nv_bfloat16* d = static_cast<__nv_bfloat16*>(out) + o;
        if (full4) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v[c][0], v[c][1]), hi = __floats2bfloat162_rn(v[c][2], v[c][3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned*>(&lo);
          pk.y = *reinterpret_cast<const unsigned*>(&hi);
          *reinterpret_cast<uint2*>(d) = pk;
        } else {
          for (int j = 0; j < 4 && p + j < npx; ++j) d[j] = __float2bfloat16_rn(v[c][j]);
        }
      } else {
        float* d = static_cast<float*>(out) + o;
        if (full4) {
          *reinterpret_cast<float4*>(d) = make_float4(v[c][0], v[c][1], v[c][2], v[c][3]);
        } else {
          for (int j = 0; j < 4 && p + j < npx; ++j) d[j] = v[c][j];
        }
      }
    }
  }
}

// ----------------------------------------------------------------------- style transfer
// dst = saturate_u8(round_half_even(|src*alpha + beta|)) in fp32 (cv2.convertScaleAbs; the product and
// the sum are fused as in OpenCV's vector body); channel 2 then becomes trunc(min(dst*gain, 255)) with
// the product in fp64 (NumPy: uint8 * python float -> float64, clipped, truncated on assignment).
// Pure u8 -> u8 map per channel: 16 bytes per thread.
__global__ void __launch_bounds__(256) style_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, long long nbytes,
                                                    float alpha, float beta, double gain, int has_gain) {
  __shared__ uint8_t lut[2][256];  // [0]: channels 0/1, [1]: channel 2
  for (int i = threadIdx.x; i < 256; i += blockDim.x) {
    const float f = fabsf(fmaf((float)i, alpha, beta));
    const int r = min(__float2int_rn(f), 255);
    lut[0][i] = (uint8_t)r;
    double g = (double)r;
    if (has_gain) g = fmin(fmax(__dmul_rn(g, gain), 0.0), 255.0);
    lut[1][i] = (uint8_t)__double2int_rz(g);
  }
  __syncthreads();
  const bool aligned = ((((uintptr_t)img) | ((uintptr_t)out)) & 15) == 0;
  const long long nvec = aligned ? nbytes / 16 : 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
    const uint4 w = ld_stream_u4(img + i * 16);
    const unsigned ws[4] = {w.x, w.y, w.z, w.w};
    unsigned os[4] = {0u, 0u, 0u, 0u};
    const int c0 = (int)((i * 16) % 3);  // channel of the first byte
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const unsigned u = (ws[k >> 2] >> ((k & 3) * 8)) & 0xffu;
      const int c = (c0 + k) % 3;
      os[k >> 2] |= (unsigned)lut[c == 2][u] << ((k & 3) * 8);
    }
    st_stream_u4(out + i * 16, make_uint4(os[0], os[1], os[2], os[3]));
  }
  for (long long i = nvec * 16 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nbytes;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = lut[(i % 3) == 2][img[i]];
}

// -------------------------------------------------------------- temperature grid search
// rows = the reference's logits.view(-1, C): C consecutive floats of the flat buffer per row (for an NCHW
// tensor that is NOT a pixel's class vector -- the reference's flattening is reproduced as is; [N,C] inputs
// mean what they say).  One pass over the logits: a tile of 1024 rows is staged in shared memory with
// 16-byte loads, every thread keeps 4 rows in registers and walks the temperature grid:
//   nll_T(row) = ln sum_c exp((z_c - zmax)/T) - (z_y - zmax)/T        (max_c z_c/T = zmax/T for T > 0)
// 19 MUFU per row and temperature: the kernel is bound by the MUFU pipe, not by HBM (77 B per row).
// Sums: fp32 per row -> fp64 per thread group -> per-warp fp64 slots in shared memory (fixed order) ->
// one partial per CTA -> fixed-order final reduction: bit-reproducible for a given device.
constexpr int kTThreads = 256;
constexpr int kTRowsPerThread = 4;
constexpr int kTTileRows = kTThreads * kTRowsPerThread;
constexpr int kTMaxTemps = 128;
constexpr int kTMaxBlocks = 1024;

template <int CS>
__global__ void __launch_bounds__(kTThreads) temperature_nll_kernel(const float* __restrict__ logits, const void* __restrict__ labels,
                                                                     int label_mode, long long nrows, int C, int ignore_index,
                                                                     const float* __restrict__ inv_t, int n_t,
                                                                     double* __restrict__ partials /*[grid][n_t+2]*/) {
  constexpr int CA = CS > 0 ? CS : AWX_MAX_CLASSES;
  const int Cn = CS > 0 ? CS : C;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_rows = reinterpret_cast<float*>(smem_raw);                        // [kTTileRows * C]
  double* s_acc = reinterpret_cast<double*>(s_rows + (size_t)kTTileRows * Cn);  // [warps][n_t]
  __shared__ float s_invt[kTMaxTemps];
  __shared__ unsigned long long s_cnt[2];  // valid rows, bad labels
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = kTThreads / 32;
  for (int i = threadIdx.x; i < n_t; i += kTThreads) s_invt[i] = inv_t[i];
  for (int i = threadIdx.x; i < nwarps * n_t; i += kTThreads) s_acc[i] = 0.0;
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0ull;
  unsigned n_valid = 0, n_bad = 0;
  const long long ntiles = (nrows + kTTileRows - 1) / kTTileRows;
  const bool aligned = (((uintptr_t)logits) & 15) == 0;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long r0 = tile * kTTileRows;
    const int rows = (int)((nrows - r0) < kTTileRows ? (nrows - r0) : kTTileRows);
    const long long f0 = r0 * Cn;
    const int nfl = rows * Cn;
    __syncthreads();
    if (aligned && (f0 & 3) == 0) {
      for (int i = threadIdx.x; i < nfl / 4; i += kTThreads)
        reinterpret_cast<float4*>(s_rows)[i] = ld_stream4(logits + f0 + 4 * (long long)i);
      for (int i = (nfl / 4) * 4 + threadIdx.x; i < nfl; i += kTThreads) s_rows[i] = logits[f0 + i];
    } else {
      for (int i = threadIdx.x; i < nfl; i += kTThreads) s_rows[i] = logits[f0 + i];
    }
    __syncthreads();
    // rows of this thread: threadIdx.x + j * kTThreads (stride-C shared reads: conflict free for odd C)
    float z[kTRowsPerThread][CA];
    float zy[kTRowsPerThread];
    bool ok[kTRowsPerThread];
#pragma unroll
    for (int j = 0; j < kTRowsPerThread; ++j) {
      const int r = threadIdx.x + j * kTThreads;
      ok[j] = false;
      zy[j] = 0.f;
      if (r < rows) {
        long long y;
        if (label_mode == AWX_LABEL_U8)
          y = static_cast<const uint8_t*>(labels)[r0 + r];
        else
          y = static_cast<const long long*>(labels)[r0 + r];
        if (y != ignore_index) {
          if (y >= 0 && y < Cn) {
            ok[j] = true;
            ++n_valid;
          } else {
            ++n_bad;
          }
        }
        const float* row = s_rows + (size_t)r * Cn;
        float mx = row[0];
#pragma unroll
        for (int c = 0; c < Cn; ++c) {
          z[j][c] = row[c];
          mx = fmaxf(mx, row[c]);
        }
        if (ok[j]) zy[j] = row[(int)y] - mx;
#pragma unroll
        for (int c = 0; c < Cn; ++c) z[j][c] -= mx;
      } else {
#pragma unroll
        for (int c = 0; c < Cn; ++c) z[j][c] = 0.f;
      }
    }
    for (int t = 0; t < n_t; ++t) {
      const float it = s_invt[t];
      const float k = it * kLog2e;
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < kTRowsPerThread; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < Cn; ++c) s += ex2_approx(z[j][c] * k);
        const float nll = fmaf(kLn2, lg2_approx(s), -zy[j] * it);
        v += ok[j] ? nll : 0.f;
      }
      double dv = (double)v;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
      if (lane == 0) s_acc[warp * n_t + t] += dv;
    }
  }
  {
    const unsigned sv = __reduce_add_sync(0xffffffffu, n_valid), sb = __reduce_add_sync(0xffffffffu, n_bad);
    __syncthreads();
    if (lane == 0) {
      atomicAdd(&s_cnt[0], (unsigned long long)sv);
      atomicAdd(&s_cnt[1], (unsigned long long)sb);
    }
  }
  __syncthreads();
  double* po = partials + (size_t)blockIdx.x * (n_t + 2);
  for (int t = threadIdx.x; t < n_t; t += kTThreads) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += s_acc[w * n_t + t];
    po[t] = s;
  }
  if (threadIdx.x == 0) {
    po[n_t] = (double)s_cnt[0];
    po[n_t + 1] = (double)s_cnt[1];
  }
}

__global__ void temperature_reduce_kernel(const double* __restrict__ partials, int nblocks, int n, double* __restrict__ sums) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partials[(size_t)b * n + t];
    sums[t] += s;
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" int awx_normalize_chw(const uint8_t* img, void* out, int32_t out_dtype, int64_t batch, int32_t H, int32_t W,
                                 const float* mean255, const float* rdenom, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_normalize_chw: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(img && out && mean255 && rdenom, AWX_E_ARG, "awx_normalize_chw: NULL pointer");
  AWX_REQUIRE(out_dtype == AWX_F32 || out_dtype == AWX_BF16, AWX_E_ARG, "awx_normalize_chw: out dtype must be AWX_F32 or AWX_BF16");
  AWX_REQUIRE(batch <= 65535, AWX_E_UNSUPPORTED, "awx_normalize_chw: batch %lld > 65535 per call", (long long)batch);
  const long long HW = (long long)H * W;
  const long long chunks = (HW + kNormChunk - 1) / kNormChunk;
  const long long cap = (long long)sm_count() * 16;
  dim3 grid((unsigned)(chunks < cap ? chunks : cap), (unsigned)batch);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (out_dtype == AWX_BF16)
    normalize_chw_kernel<true><<<grid, kNormThreads, 0, s>>>(img, out, HW, mean255[0], mean255[1], mean255[2], rdenom[0],
                                                             rdenom[1], rdenom[2]);
  else
    normalize_chw_kernel<false><<<grid, kNormThreads, 0, s>>>(img, out, HW, mean255[0], mean255[1], mean255[2], rdenom[0],
                                                              rdenom[1], rdenom[2]);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

extern "C" int awx_style_transfer(const uint8_t* img, uint8_t* out, int64_t n_pixels, float alpha, float beta, double blue_gain,
                                  int32_t has_gain, void* stream) {
  AWX_REQUIRE(n_pixels >= 0, AWX_E_ARG, "awx_style_transfer: negative size");
  if (n_pixels == 0) return AWX_OK;
  AWX_REQUIRE(img && out, AWX_E_ARG, "awx_style_transfer: NULL pointer");
  const long long nbytes = n_pixels * 3;
  long long blocks = (nbytes / 16 + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  style_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(img, out, nbytes, alpha, beta, blue_gain, has_gain);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

extern "C" size_t awx_temperature_workspace_bytes(int32_t n_temps) {
  if (n_temps <= 0 || n_temps > kTMaxTemps) return 0;
  return ((size_t)kTMaxBlocks * (n_temps + 2) * sizeof(double) + n_temps * sizeof(float) + 255) & ~(size_t)255;
}

extern "C" int awx_temperature_nll(const float* logits, const void* labels, int32_t label_dtype, int64_t rows, int32_t C,
                                   int32_t ignore_index, const float* temperatures, int32_t n_temps, double* sums,
                                   void* workspace, void* stream) {
  AWX_REQUIRE(rows >= 0, AWX_E_ARG, "awx_temperature_nll: negative size");
  AWX_REQUIRE(n_temps >= 1 && n_temps <= kTMaxTemps, AWX_E_UNSUPPORTED, "awx_temperature_nll: %d temperatures outside 1..%d", n_temps, kTMaxTemps);
  AWX_REQUIRE(C >= 1 && C <= AWX_MAX_CLASSES, AWX_E_UNSUPPORTED, "awx_temperature_nll: num_classes %d outside 1..%d", C, AWX_MAX_CLASSES);
  AWX_REQUIRE(label_dtype == AWX_LABEL_U8 || label_dtype == AWX_LABEL_I64, AWX_E_ARG, "awx_temperature_nll: unknown label dtype");
  if (rows == 0) return AWX_OK;
  AWX_REQUIRE(logits && labels && temperatures && sums && workspace, AWX_E_ARG, "awx_temperature_nll: NULL pointer");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  double* partials = static_cast<double*>(workspace);
  float* d_invt = reinterpret_cast<float*>(partials + (size_t)kTMaxBlocks * (n_temps + 2));
  float h_invt[kTMaxTemps];
  for (int i = 0; i < n_temps; ++i) {
    AWX_REQUIRE(temperatures[i] > 0.f, AWX_E_ARG, "awx_temperature_nll: temperature %d is not positive", i);
    h_invt[i] = 1.0f / temperatures[i];
  }
  AWX_CUDA(cudaMemcpyAsync(d_invt, h_invt, n_temps * sizeof(float), cudaMemcpyHostToDevice, s));
  const long long ntiles = (rows + kTTileRows - 1) / kTTileRows;
  long long blocks = (long long)sm_count() * 2;
  if (blocks > ntiles) blocks = ntiles;
  if (blocks > kTMaxBlocks) blocks = kTMaxBlocks;
  const size_t smem = (size_t)kTTileRows * C * sizeof(float) + (size_t)(kTThreads / 32) * n_temps * sizeof(double);
  if (C == 19) {
    AWX_CUDA(cudaFuncSetAttribute(temperature_nll_kernel<19>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    temperature_nll_kernel<19><<<(unsigned)blocks, kTThreads, smem, s>>>(logits, labels, label_dtype, rows, C, ignore_index, d_invt,
                                                                         n_temps, partials);
  } else {
    AWX_REQUIRE(smem <= 200 * 1024, AWX_E_UNSUPPORTED, "awx_temperature_nll: num_classes %d needs too much shared memory", C);
    AWX_CUDA(cudaFuncSetAttribute(temperature_nll_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    temperature_nll_kernel<0><<<(unsigned)blocks, kTThreads, smem, s>>>(logits, labels, label_dtype, rows, C, ignore_index, d_invt,
                                                                        n_temps, partials);
  }
  AWX_CUDA(cudaGetLastError());
  note_launch();
  temperature_reduce_kernel<<<1, 128, 0, s>>>(partials, (int)blocks, n_temps + 2, sums);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}
