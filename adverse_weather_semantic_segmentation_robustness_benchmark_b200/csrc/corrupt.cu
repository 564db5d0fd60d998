// awx_corrupt: fog / rain / snow / night weather corruption of uint8 HWC frames (and clean copy).
//
// Three kernels, chosen per image by AwxCorruptParams.kind:
//   rasterize_kernel  streaks (cv2.line) and flakes (cv2.circle) -> 1 bit/pixel overlay mask
//   pointwise_kernel  clean / fog / night: 1024-pixel chunks staged through shared memory so that
//                     global traffic is 128-bit and coalesced although a pixel is 3 bytes
//   blur_kernel<R>    rain / snow: 16x128 pixel tiles (+R halo, BORDER_REFLECT_101) in shared
//                     memory: point op + overlay -> horizontal pass -> vertical pass -> uint8
//
// Arithmetic mirrors the reference's dtype promotions (data/preprocessing.py):
//   u8 -> fp32 by a true division by 255 (256-entry table built with __fdiv_rn)        :81
//   fog:   fp64, separately rounded ops; airlight is rounded to fp32 first (A*ones_like) :117-123
//   night: fp32 dim + colour shift, then fp64 noise add                                 :213-225
//   rain / snow: fp32 throughout, OpenCV-style separable Gaussian                       :135-168, :180-202
//   final: clip to [0,1], times 255, truncate toward zero
#include <cuda_bf16.h>

#include <cstdlib>

#include "awx_internal.cuh"
#include "raster.cuh"
#include "convert.cuh"
#include "blur_strip.cuh"
#include "fog_fast.cuh"

namespace awx {
namespace {

constexpr int kChunkPx = 1024;          // pointwise: pixels per chunk (3072 B = 192 x 16 B)
constexpr int kPointThreads = 256;
constexpr int kTileW = 128, kTileH = 16;  // blur tile
constexpr int kBlurThreads = 256;

__device__ __forceinline__ int reflect101(int i, int n) {
  // BORDER_REFLECT_101: gfedcb|abcdefgh|gfedcba
  if (n == 1) return 0;
  while (i < 0 || i >= n) i = i < 0 ? -i : 2 * n - 2 - i;
  return i;
}

__device__ __forceinline__ unsigned to_u8_f64(double v) {
  // (np.clip(v, 0, 1) * 255).astype(uint8): truncation toward zero and the clip commute (v * 255 in (-1, 0) and the
  // clipped 0 both give 0; v > 1 gives >= 255 either way), so the clamp runs on the integer pipe instead of two
  // fp64 min / max
  const int i = __double2int_rz(__dmul_rn(v, 255.0));
  return (unsigned)min(max(i, 0), 255);
}
__device__ __forceinline__ unsigned to_u8_f32(float v) {
  // (np.clip(v, 0, 1) * 255).astype(uint8): the clip commutes with the truncation (see to_u8_f64), and the
  // float -> u8 conversion saturates to [0, 255] by itself: one FMUL + one F2I
  unsigned r;
  asm("cvt.rzi.u8.f32 %0, %1;" : "=r"(r) : "f"(__fmul_rn(v, 255.0f)));
  return r;
}

// Optional fused epilogue: albumentations Normalize + ToTensorV2 (data/loader.py:196-199) of the corrupted
// frame, written as CHW fp32 / bf16 next to (or instead of) the uint8 HWC frame.
struct NormOut {
  void* ptr;      // [B,3,H,W]; nullptr = no normalised output
  int bf16;
  float mean[3];  // mean * 255
  float rden[3];  // 1 / (std * 255)
};
__device__ __forceinline__ float norm_value(const NormOut& n, unsigned u8, int c) {
  return __fmul_rn(__fsub_rn((float)u8, n.mean[c]), n.rden[c]);
}
__device__ __forceinline__ void norm_store1(const NormOut& n, size_t idx, float v) {
  if (n.bf16)
    static_cast<__nv_bfloat16*>(n.ptr)[idx] = __float2bfloat16_rn(v);
  else
    static_cast<float*>(n.ptr)[idx] = v;
}

// -------------------------------------------------------------------------- overlay mask
__global__ void __launch_bounds__(128) rasterize_kernel(const AwxCorruptParams* __restrict__ params,
                                                         const int32_t* __restrict__ items, unsigned* __restrict__ mask,
                                                         int H, int W, int WW) {
  const int b = blockIdx.y;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != AWX_RAIN && prm.kind != AWX_SNOW) return;
  unsigned* m = mask + (size_t)b * H * WW;
  auto emit = [&](int y, int x0, int x1) {
    unsigned* row = m + (size_t)y * WW;
    const int w0 = x0 >> 5, w1 = x1 >> 5;
    for (int wd = w0; wd <= w1; ++wd) {
      const int lo = wd == w0 ? (x0 & 31) : 0;
      const int hi = wd == w1 ? (x1 & 31) : 31;
      const unsigned bits = (hi == 31 ? 0xffffffffu : ((1u << (hi + 1)) - 1u)) & ~((1u << lo) - 1u);
      atomicOr(row + wd, bits);
    }
  };
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < prm.item_count; i += gridDim.x * blockDim.x) {
    const int32_t* it = items + 5 * (size_t)(prm.item_begin + i);
    if (prm.kind == AWX_RAIN)
      raster::line(it[0], it[1], it[2], it[3], it[4], W, H, emit);
    else
      raster::disc(it[0], it[1], it[2], W, H, emit);
  }
}

// ------------------------------------------------------------------- clean / fog / night
template <typename FT>
__global__ void __launch_bounds__(kPointThreads, 5) pointwise_kernel(const uint8_t* __restrict__ img,
                                                                   uint8_t* __restrict__ out,
                                                                   const AwxCorruptParams* __restrict__ params,
                                                                   const FT* __restrict__ field, long long HW,
                                                                   const __grid_constant__ NormOut norm, int fog_elsewhere /* bit 0: fog_kernel, bit 1: night_kernel take their images */) {
  const int b = blockIdx.y;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != AWX_CLEAN && prm.kind != AWX_FOG && prm.kind != AWX_NIGHT) return;
  if (prm.kind == AWX_FOG && (fog_elsewhere & 1) && prm.field_offset % 2 == 0 && fog::params_ok(prm.d0, prm.d1)) return;  // fog_kernel has this image
  if (prm.kind == AWX_NIGHT && (fog_elsewhere & 2) && prm.field_offset % 2 == 0) return;                                   // night_kernel has it
  __shared__ __align__(16) unsigned s_in[kChunkPx * 3 / 4];
  __shared__ __align__(16) unsigned s_out[kChunkPx * 3 / 4];
  const uint8_t* src = img + (size_t)b * HW * 3;
  uint8_t* dst = out ? out + (size_t)b * HW * 3 : nullptr;
  const bool aligned = (((uintptr_t)src | (uintptr_t)dst) & 15) == 0;
  const long long nchunks = (HW + kChunkPx - 1) / kChunkPx;
  const FT* fld = field ? field + prm.field_offset : nullptr;
  const double neg_beta = -prm.d0;  // numpy evaluates (-beta) * depth
  const double airlight = prm.d1;   // already rounded to fp32 by the host (A * ones_like(fp32))
  const float gain = prm.f0;
  const double half_i = prm.d0 * 0.5;  // night: intensity / 2 (exact)

  for (long long ch = blockIdx.x; ch < nchunks; ch += gridDim.x) {
    const long long px0 = ch * kChunkPx;
    const int npx = (int)((HW - px0) < kChunkPx ? (HW - px0) : kChunkPx);
    const bool full = aligned && npx == kChunkPx;
    __syncthreads();  // previous iteration's stores have left s_out; table is ready
    if (full) {
      if (threadIdx.x < kChunkPx * 3 / 16)
        reinterpret_cast<uint4*>(s_in)[threadIdx.x] = ld_stream_u4(src + px0 * 3 + threadIdx.x * 16);
    } else {
      uint8_t* sb = reinterpret_cast<uint8_t*>(s_in);
      for (int i = threadIdx.x; i < npx * 3; i += kPointThreads) sb[i] = src[px0 * 3 + i];
    }
    __syncthreads();
    const int p = threadIdx.x * 4;  // 4 pixels = 12 bytes = 3 words; bank = 3*t mod 32: conflict free
    if (p < npx) {
      unsigned wv[3] = {s_in[threadIdx.x * 3], s_in[threadIdx.x * 3 + 1], s_in[threadIdx.x * 3 + 2]};
      if (prm.kind != AWX_CLEAN) {
        unsigned ov[3] = {0u, 0u, 0u};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const bool live = p + j < npx;
          double tr = 0.0, veil = 0.0;
          if (prm.kind == AWX_FOG) {
            const double d = live ? (double)fld[px0 + p + j] : 1.0;
            tr = exp(__dmul_rn(neg_beta, d));
            veil = __dmul_rn(airlight, __dsub_rn(1.0, tr));
          }
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            const int k = j * 3 + c;  // byte index within the 12
            const unsigned u = (wv[k >> 2] >> ((k & 3) * 8)) & 0xffu;
            const float x = unit_of_u8(u);
            unsigned r;
            if (prm.kind == AWX_FOG) {
              r = to_u8_f64(__dadd_rn(__dmul_rn((double)x, tr), veil));
            } else {  // night
              const float shift = c == 0 ? 0.8f : (c == 1 ? 0.85f : 1.2f);
              const float v = __fmul_rn(__fmul_rn(x, gain), shift);
              const double nz = live ? (double)fld[(px0 + p + j) * 3 + c] : 0.0;
              // noise * I * 0.5: the halving is exact, so fl(fl(nz * I) * 0.5) == fl(nz * (I * 0.5)): one fp64 multiply
              r = to_u8_f64(__dadd_rn((double)v, __dmul_rn(nz, half_i)));
            }
            ov[k >> 2] |= r << ((k & 3) * 8);
          }
        }
        wv[0] = ov[0];
        wv[1] = ov[1];
        wv[2] = ov[2];
      }
      s_out[threadIdx.x * 3] = wv[0];
      s_out[threadIdx.x * 3 + 1] = wv[1];
      s_out[threadIdx.x * 3 + 2] = wv[2];
      if (norm.ptr) {  // fused Normalize + CHW: 4 pixels -> one vector per channel plane
        const bool vec4 = (HW & 3) == 0 && p + 3 < npx;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          float v[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = j * 3 + c;
            v[j] = norm_value(norm, (wv[k >> 2] >> ((k & 3) * 8)) & 0xffu, c);
          }
          const size_t o = ((size_t)b * 3 + c) * HW + px0 + p;
          if (vec4 && !norm.bf16) {
            *reinterpret_cast<float4*>(static_cast<float*>(norm.ptr) + o) = make_float4(v[0], v[1], v[2], v[3]);
          } else if (vec4) {
            const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
            uint2 pk;
            pk.x = *reinterpret_cast<const unsigned*>(&lo);
            pk.y = *reinterpret_cast<const unsigned*>(&hi);
            *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(norm.ptr) + o) = pk;
          } else {
            for (int j = 0; j < 4 && p + j < npx; ++j) norm_store1(norm, o + j, v[j]);
          }
        }
      }
    }
    __syncthreads();
    if (!dst) continue;
    if (full) {
      if (threadIdx.x < kChunkPx * 3 / 16)
        st_stream_u4(dst + px0 * 3 + threadIdx.x * 16, reinterpret_cast<const uint4*>(s_out)[threadIdx.x]);
    } else {
      const uint8_t* sb = reinterpret_cast<const uint8_t*>(s_out);
      for (int i = threadIdx.x; i < npx * 3; i += kPointThreads) dst[px0 * 3 + i] = sb[i];
    }
  }
}

// ------------------------------------------------------------------------------------------ fog, screened
// fog_fast.cuh: an fp32 screen decides the byte wherever 255 v is not within 3e-4 of an integer, the reference's fp64
// expression (exp and all) runs for the ~2e-3 of the pixels it cannot decide; outputs are those of
// pointwise_kernel<double> bit for bit (tests/test_fog_fast_cpu.py, tests/test_fog_fast_gpu.py).
// No shared memory, no barrier: a thread owns a UNIT of 16 pixels -- 48 bytes in (3 x 16 B), 16 fp64 depths (8 x 16 B),
// 48 bytes out -- all eleven loads are issued before the first use, and consecutive lanes touch consecutive units, so
// every warp-wide access is contiguous.  Needs H*W % 16 == 0 and 16-byte aligned tensors; anything else goes to the
// generic kernel.
// build knobs; measured per 64 frames: 256 threads x 3 CTAs per SM 0.386 ms, x 4 (64 registers, spills) 0.410, x 5 0.447,
// 128 threads x 6 (the same 24 warps in smaller CTAs: faster turnover, a thread handles one unit and exits) 0.352
#ifndef AWX_FOG_THREADS
#define AWX_FOG_THREADS 128
#endif
#ifndef AWX_FOG_CTAS
#define AWX_FOG_CTAS 6
#endif
constexpr int kFogThreads = AWX_FOG_THREADS;
__global__ void __launch_bounds__(kFogThreads, AWX_FOG_CTAS) fog_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out,
                                                              const AwxCorruptParams* __restrict__ params,
                                                              const double* __restrict__ field, long long HW) {
  const int b = blockIdx.y;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != AWX_FOG || prm.field_offset % 2 != 0 || !fog::params_ok(prm.d0, prm.d1)) return;  // the generic kernel has it
  const uint4* src = reinterpret_cast<const uint4*>(img + (size_t)b * HW * 3);
  uint4* dst = reinterpret_cast<uint4*>(out + (size_t)b * HW * 3);
  const double2* fld = reinterpret_cast<const double2*>(field + prm.field_offset);
  const fog::Params fp = fog::make_params(prm.d0, prm.d1);
  const long long units = HW / 16;
  for (long long u = (long long)blockIdx.x * kFogThreads + threadIdx.x; u < units; u += (long long)gridDim.x * kFogThreads) {
    double2 d[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) d[i] = __ldg(fld + u * 8 + i);
    uint4 w[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) w[i] = ld_stream_u4(src + u * 3 + i);
    const unsigned wi[12] = {w[0].x, w[0].y, w[0].z, w[0].w, w[1].x, w[1].y, w[1].z, w[1].w, w[2].x, w[2].y, w[2].z, w[2].w};
    unsigned wo[12];
#pragma unroll
    for (int g = 0; g < 4; ++g) {  // four pixels = three words at a time
      const unsigned in3[3] = {wi[3 * g], wi[3 * g + 1], wi[3 * g + 2]};
      const double dep[4] = {d[2 * g].x, d[2 * g].y, d[2 * g + 1].x, d[2 * g + 1].y};
      unsigned o3[3];
      fog::fog4(in3, dep, fp, o3);
      wo[3 * g] = o3[0];
      wo[3 * g + 1] = o3[1];
      wo[3 * g + 2] = o3[2];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) st_stream_u4(dst + u * 3 + i, make_uint4(wo[4 * i], wo[4 * i + 1], wo[4 * i + 2], wo[4 * i + 3]));
  }
}

// the screened fog kernel takes whole 16-pixel units of 16-byte aligned tensors
bool fog_kernel_ok(const uint8_t* img, const uint8_t* out, const void* field, long long HW) {
  return HW % 16 == 0 && (((uintptr_t)img | (uintptr_t)out | (uintptr_t)field) & 15) == 0;
}

// ------------------------------------------------------------------------------------------ night, flat
// Night is a pure element-wise map over the 3 H W values of a frame with a period-3 constant (the colour shift):
//   out[i] = trunc(clip((double)fl(fl(fl(u[i] / 255) * gain) * shift[i % 3]) + noise[i] * (I / 2), 0, 1) * 255)
// so the kernel ignores pixels altogether.  A warp owns 512 consecutive values per iteration; in load j (of 8) lane
// l takes the PAIR 64 j + 2 l: one 16-byte load of two fp64 noise values (a warp reads 512 contiguous bytes per
// instruction -- the 30 bytes per pixel are all noise, and they stream fully coalesced with all eight loads in
// flight before the first use), one 2-byte load of the two input bytes, one 2-byte store.  No shared memory, no
// barrier (the staged generic kernel: three barriers per 1024 pixels).  Arithmetic as in pointwise_kernel: fp32 dim and
// shift, separately rounded; fp64 noise add; exact.  Needs an even number of values per frame, a 16-byte aligned
// field and an even field offset; anything else goes to the generic kernel.
#ifndef AWX_NIGHT_THREADS
#define AWX_NIGHT_THREADS 256
#endif
constexpr int kNightThreads = AWX_NIGHT_THREADS;
constexpr int kNightPairs = 8;                                  // pairs per lane and iteration
constexpr int kNightBlockValues = kNightThreads * 2 * kNightPairs;  // 4096
__global__ void __launch_bounds__(kNightThreads, 1024 / kNightThreads) night_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out,
                                                                  const AwxCorruptParams* __restrict__ params,
                                                                  const double* __restrict__ field, long long HW) {
  const int b = blockIdx.y;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != AWX_NIGHT || (prm.field_offset & 1) != 0) return;  // the generic kernel has it
  const long long N = HW * 3;
  const uint8_t* src = img + (size_t)b * N;
  uint8_t* dst = out + (size_t)b * N;
  const double* fld = field + prm.field_offset;
  const float gain = prm.f0;
  const double half_i = prm.d0 * 0.5;  // noise * I * 0.5: the halving is exact, one fp64 multiply
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (long long base = (long long)blockIdx.x * kNightBlockValues; base < N; base += (long long)gridDim.x * kNightBlockValues) {
    const long long i0 = base + warp * (64 * kNightPairs) + 2 * lane;  // this lane's first pair
    double2 nz[kNightPairs];
    unsigned short by[kNightPairs];
#pragma unroll
    for (int j = 0; j < kNightPairs; ++j) {
      const long long i = i0 + 64 * j;
      if (i < N) {
        nz[j] = __ldg(reinterpret_cast<const double2*>(fld + i));
        by[j] = __ldg(reinterpret_cast<const unsigned short*>(src + i));
      }
    }
    // colour shifts rotated to this lane's phase: value i has channel i % 3, pair j starts at channel (c0 + j) % 3
    const int c0 = (int)(i0 % 3);
    const float sh0 = c0 == 0 ? 0.8f : (c0 == 1 ? 0.85f : 1.2f);
    const float sh1 = c0 == 0 ? 0.85f : (c0 == 1 ? 1.2f : 0.8f);
    const float sh2 = c0 == 0 ? 1.2f : (c0 == 1 ? 0.8f : 0.85f);
    const float sh[3] = {sh0, sh1, sh2};
#pragma unroll
    for (int j = 0; j < kNightPairs; ++j) {
      const long long i = i0 + 64 * j;
      if (i < N) {
        const float va = __fmul_rn(__fmul_rn(unit_of_u8(by[j] & 0xffu), gain), sh[j % 3]);
        const float vb = __fmul_rn(__fmul_rn(unit_of_u8(by[j] >> 8), gain), sh[(j + 1) % 3]);
        const unsigned ra = to_u8_f64(__dadd_rn((double)va, __dmul_rn(nz[j].x, half_i)));
        const unsigned rb = to_u8_f64(__dadd_rn((double)vb, __dmul_rn(nz[j].y, half_i)));
        *reinterpret_cast<unsigned short*>(dst + i) = (unsigned short)(ra | (rb << 8));
      }
    }
  }
}

bool night_kernel_ok(const uint8_t* img, const uint8_t* out, const void* field, long long HW) {
  const char* e = getenv("AWX_NIGHT_KERNEL");  // =staged: the generic kernel (A/B measurements, parity tests of both)
  return !(e && e[0] == 's') && (HW * 3) % 2 == 0 && (((uintptr_t)img | (uintptr_t)out) & 1) == 0 && (((uintptr_t)field) & 15) == 0;
}

// --------------------------------------------------------------------------- rain / snow
// One CTA filters a 16 x 128 pixel tile.  Everything is indexed in ELEMENTS (bytes of the HWC row: 3 per
// pixel), because both filter passes work per channel and an element's neighbours are 3 elements away.
//   stage 1  load.  A thread takes a UNIT of 16 pixels = 48 bytes = three 16-byte loads (units start
//            at multiples of 16 pixels, so they are pixel aligned AND 16-byte aligned), converts through
//            the u8 -> fp32 table, applies the point operation and the overlay bit (one mask word
//            covers the unit) and writes 12 float4 to s_pre.  The tile needs R <= 3 halo pixels on each
//            side; it stages one whole unit per side (rows: BORDER_REFLECT_101 on the row index;
//            columns: units that are not fully inside the image take the per-pixel path).
//   stage 2  horizontal pass, register blocked: a thread produces 12 consecutive elements of one row
//            from 5 (R=1) or 9 (R=3) float4 loads of s_pre -> s_h (3 float4 stores).
//   stage 3  vertical pass, register blocked: a thread owns 4 adjacent element columns and 8 output
//            rows: 8 + 2R float4 loads of s_h, then clip * 255, truncate, one 32-bit store per row
//            (a warp writes 128 contiguous bytes).
// Shared-memory traffic per output element drops from 14 scalar loads to ~1.3 vector accesses.
constexpr int kUnitPx = 16;                          // pixels per load unit
constexpr int kPreW = (kTileW + 2 * kUnitPx) * 3 + 4;  // s_pre row: tile + one unit of halo per side, +4 so that
                                                      // 4 consecutive rows start 4 banks apart (elements)
constexpr int kRowE = kTileW * 3;                    // output elements per tile row
constexpr int kHBlock = 12;                          // elements per horizontal-pass task

template <int R, bool RAIN>
__global__ void __launch_bounds__(kBlurThreads) blur_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out,
                                                             const AwxCorruptParams* __restrict__ params,
                                                             const unsigned* __restrict__ mask, int H, int W, int WW,
                                                             const __grid_constant__ NormOut norm) {
  constexpr int PH = kTileH + 2 * R;
  const int b = blockIdx.z;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != (RAIN ? AWX_RAIN : AWX_SNOW) || prm.blur_k != 2 * R + 1) return;

  extern __shared__ __align__(16) unsigned char smem[];
  float* s_pre = reinterpret_cast<float*>(smem);  // [PH][kPreW] point-op'ed, overlaid, fp32
  float* s_h = s_pre + PH * kPreW;                // [PH][kRowE] after the horizontal pass

  const int x0 = blockIdx.x * kTileW, y0 = blockIdx.y * kTileH;
  const uint8_t* src = img + (size_t)b * H * W * 3;
  uint8_t* dst = out ? out + (size_t)b * H * W * 3 : nullptr;
  const unsigned* m = mask + (size_t)b * H * WW;
  const float k1 = prm.f0, k2 = prm.f1;  // rain: x*k1 + k2 ; snow: clip(x + k1)
  const bool vec_ok = ((W * 3) & 15) == 0 && (((uintptr_t)src) & 15) == 0;

  auto point = [&](unsigned u8, int c, bool over) -> float {
    const float x = unit_of_u8(u8);
    if (RAIN) {
      const float v = __fadd_rn(__fmul_rn(x, k1), k2);
      return over ? (c == 0 ? 0.8f : (c == 1 ? 0.9f : 1.0f)) : v;
    }
    const float v = fminf(fmaxf(__fadd_rn(x, k1), 0.0f), 1.0f);
    return over ? 1.0f : v;
  };

  // ---- stage 1: load units, point op, overlay -> s_pre
  // Interior units first: a warp takes 4 rows x 8 units (rows fastest), so that a quarter warp (4 rows x 2 units)
  // writes 8 distinct 16-byte bank groups and the warp still reads 8 consecutive units of each row.  The two halo
  // units of a row only contribute their R <= 3 pixels next to the tile: one 16-byte chunk each, in a second loop
  // (in the same loop their lanes would idle through the other two chunks of the interior lanes).
  constexpr int UPR = kTileW / kUnitPx + 2;  // units per staged row (8 interior + 2 halo)
  auto stage_unit = [&](int ry, int un, int q_lo, int q_hi) {
    const int sy = reflect101(y0 + ry - R, H);
    const int ux = x0 + (un - 1) * kUnitPx;  // first pixel of the unit (may be outside the image)
    float* o = s_pre + ry * kPreW + un * (kUnitPx * 3);
    if (vec_ok && ux >= 0 && ux + kUnitPx <= W) {
      const uint8_t* g = src + ((size_t)sy * W + ux) * 3;
      const unsigned mw = m[(size_t)sy * WW + (ux >> 5)] >> (ux & 31);  // bit j = pixel ux + j
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        if (q < q_lo || q > q_hi) continue;
        const uint4 w = ld_stream_u4(g + 16 * q);
        const unsigned ws[4] = {w.x, w.y, w.z, w.w};
        float f[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          const int e = 16 * q + k;  // element within the unit: pixel e / 3, channel e % 3
          f[k] = point((ws[k >> 2] >> ((k & 3) * 8)) & 0xffu, e % 3, (mw >> (e / 3)) & 1u);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          reinterpret_cast<float4*>(o + 16 * q)[k] = make_float4(f[4 * k], f[4 * k + 1], f[4 * k + 2], f[4 * k + 3]);
      }
    } else {
      // image border / unaligned rows: per pixel, only the pixels the filter can reach
      for (int j = 0; j < kUnitPx; ++j) {
        const int xx = ux + j;
        if (xx < x0 - R || xx >= x0 + kTileW + R) continue;
        const int sx = reflect101(xx, W);
        const bool over = (m[(size_t)sy * WW + (sx >> 5)] >> (sx & 31)) & 1u;
        const uint8_t* px = src + ((size_t)sy * W + sx) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) o[j * 3 + c] = point(px[c], c, over);
      }
    }
  };
  constexpr int UI = UPR - 2;
  constexpr int NI = ((PH + 3) / 4) * 4 * UI;  // interior tasks (rows padded to a multiple of 4)
  static_assert(NI + 2 * PH <= kBlurThreads, "one task per thread: all global loads of the tile are in flight at once");
  {
    const int i = threadIdx.x;
    if (i < NI) {
      const int grp = i / (4 * UI), j = i - grp * (4 * UI);
      const int ry = grp * 4 + (j & 3);
      if (ry < PH) stage_unit(ry, 1 + (j >> 2), 0, 2);
    } else if (i < NI + 2 * PH) {
      // halo: the last chunk (pixels 10.67 .. 15) of the left unit, the first chunk (pixels 0 .. 5.33) of the right one
      const int h = i - NI, ry = h >> 1;
      if (h & 1)
        stage_unit(ry, UPR - 1, 0, 0);
      else
        stage_unit(ry, 0, 2, 2);
    }
  }
  __syncthreads();

  // ---- stage 2: horizontal pass (12 elements per task)
  const float t0 = prm.taps[0], t1 = prm.taps[1], t2 = prm.taps[2], t3 = prm.taps[3];
  constexpr int HB = kRowE / kHBlock;  // 32 tasks per row
  // output element col <-> s_pre element kUnitPx*3 + col; taps reach 3R elements either way.
  // first float4 that covers element 48 + col0 - 3R (col0 a multiple of 12): 36 + col0 (R=3), 44 + col0 (R=1)
  constexpr int LEAD = R == 3 ? 3 : 1;     // elements loaded before the first one needed
  constexpr int NV = R == 3 ? 9 : 5;       // float4 loads per task (12 + 6R elements after LEAD)
  constexpr int BASE = kUnitPx * 3 - 3 * R - LEAD;
  for (int i = threadIdx.x; i < PH * HB; i += kBlurThreads) {
    const int ry = i / HB, blk = i - ry * HB;
    const float4* pv = reinterpret_cast<const float4*>(s_pre + ry * kPreW + BASE + blk * kHBlock);
    float w[4 * NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const float4 q = pv[k];
      w[4 * k] = q.x;
      w[4 * k + 1] = q.y;
      w[4 * k + 2] = q.z;
      w[4 * k + 3] = q.w;
    }
    float r[kHBlock];
#pragma unroll
    for (int e = 0; e < kHBlock; ++e) {
      const float* p = w + LEAD + 3 * R + e;  // centre tap
      float v;
      if (R == 1) {
        v = fmaf(p[0], t0, __fmul_rn(p[-3] + p[3], t1));  // SymmRowSmallVec_32f: the side product is rounded, the centre fused
      } else {
        // RowVec_32f: leftmost tap first, its product rounded, every further tap fused (written out: left to the
        // compiler, `a * b + c * d` fuses the FIRST product and rounds the second)
        v = __fmul_rn(p[-9], t3);
        v = fmaf(p[-6], t2, v);
        v = fmaf(p[-3], t1, v);
        v = fmaf(p[0], t0, v);
        v = fmaf(p[3], t1, v);
        v = fmaf(p[6], t2, v);
        v = fmaf(p[9], t3, v);
      }
      r[e] = v;
    }
    float4* ov = reinterpret_cast<float4*>(s_h + ry * kRowE + blk * kHBlock);
#pragma unroll
    for (int k = 0; k < 3; ++k) ov[k] = make_float4(r[4 * k], r[4 * k + 1], r[4 * k + 2], r[4 * k + 3]);
  }
  __syncthreads();

  // ---- stage 3: vertical pass (4 columns x 8 rows per task), clip, scale, truncate, store
  const int tw3 = min(kTileW, W - x0) * 3, th = min(kTileH, H - y0);
  constexpr int CG = kRowE / 4;  // 96 column groups
  constexpr int VR = 8;          // output rows per task
  const bool st32 = ((W * 3) & 3) == 0 && (((uintptr_t)dst) & 3) == 0;
  const size_t row_bytes = (size_t)W * 3;
  unsigned* s_o = reinterpret_cast<unsigned*>(s_pre);  // [kTileH][CG] packed output bytes (norm epilogue only)
  for (int i = threadIdx.x; i < CG * (kTileH / VR); i += kBlurThreads) {
    const int half = i / CG, cg = i - half * CG;
    const float4* pv = reinterpret_cast<const float4*>(s_h + (half * VR) * kRowE + cg * 4);
    float4 win[VR + 2 * R];
#pragma unroll
    for (int k = 0; k < VR + 2 * R; ++k) win[k] = pv[k * (kRowE / 4)];
    uint8_t* drow = dst ? dst + ((size_t)(y0 + half * VR) * W + x0) * 3 + cg * 4 : nullptr;  // first output row of the task
#pragma unroll
    for (int rr = 0; rr < VR; ++rr) {
      const int ry = half * VR + rr;
      if (ry >= th) break;
      auto vfilt = [&](float c0, float m1, float p1, float m2, float p2, float m3, float p3) -> unsigned {
        float v = fmaf(m1 + p1, t1, __fmul_rn(c0, t0));  // SymmColumnVec_32f: the centre product is rounded, the rest fused
        if (R == 3) {
          v = fmaf(m2 + p2, t2, v);
          v = fmaf(m3 + p3, t3, v);
        }
        return to_u8_f32(v);
      };
      const float4 c = win[rr + R], a1 = win[rr + R - 1], b1 = win[rr + R + 1];
      const float4 a2 = win[R == 3 ? rr + R - 2 : 0], b2 = win[R == 3 ? rr + R + 2 : 0];
      const float4 a3 = win[R == 3 ? rr + R - 3 : 0], b3 = win[R == 3 ? rr + R + 3 : 0];
      const unsigned o0 = vfilt(c.x, a1.x, b1.x, a2.x, b2.x, a3.x, b3.x);
      const unsigned o1 = vfilt(c.y, a1.y, b1.y, a2.y, b2.y, a3.y, b3.y);
      const unsigned o2 = vfilt(c.z, a1.z, b1.z, a2.z, b2.z, a3.z, b3.z);
      const unsigned o3 = vfilt(c.w, a1.w, b1.w, a2.w, b2.w, a3.w, b3.w);
      const unsigned ob[4] = {o0, o1, o2, o3};
      // fused Normalize + CHW: park the packed bytes in shared memory (s_pre is dead by now); the plane
      // writes below need 4 pixels of ONE channel per thread to be vector stores
      if (norm.ptr) s_o[ry * CG + cg] = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
      if (dst) {
        uint8_t* d = drow + (size_t)rr * row_bytes;
        if (st32 && cg * 4 + 3 < tw3) {
          *reinterpret_cast<unsigned*>(d) = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            if (cg * 4 + k < tw3) d[k] = (uint8_t)ob[k];
        }
      }
    }
  }
  if (norm.ptr) {
    __syncthreads();
    // thread task = 4 consecutive pixels of one row: 12 bytes = words 3q..3q+2 (conflict free), one 16-byte
    // (fp32) or 8-byte (bf16) store per channel plane
    const int tw = min(kTileW, W - x0);
    const bool vec4 = (W & 3) == 0;
    for (int i = threadIdx.x; i < th * (kTileW / 4); i += kBlurThreads) {
      const int ry = i / (kTileW / 4), q = i - ry * (kTileW / 4);
      if (q * 4 >= tw) continue;
      const unsigned wv[3] = {s_o[ry * CG + 3 * q], s_o[ry * CG + 3 * q + 1], s_o[ry * CG + 3 * q + 2]};
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int k = j * 3 + c;
          v[j] = norm_value(norm, (wv[k >> 2] >> ((k & 3) * 8)) & 0xffu, c);
        }
        const size_t o = (((size_t)b * 3 + c) * H + (y0 + ry)) * W + x0 + q * 4;
        if (vec4 && q * 4 + 3 < tw && !norm.bf16) {
          *reinterpret_cast<float4*>(static_cast<float*>(norm.ptr) + o) = make_float4(v[0], v[1], v[2], v[3]);
        } else if (vec4 && q * 4 + 3 < tw) {
          const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
          uint2 pk;
          pk.x = *reinterpret_cast<const unsigned*>(&lo);
          pk.y = *reinterpret_cast<const unsigned*>(&hi);
          *reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(norm.ptr) + o) = pk;
        } else {
          for (int j = 0; j < 4 && q * 4 + j < tw; ++j) norm_store1(norm, o + j, v[j]);
        }
      }
    }
  }
}

template <int R>
constexpr size_t blur_smem() {
  return (size_t)(kTileH + 2 * R) * (kPreW + kRowE) * sizeof(float);
}

template <int R, bool RAIN>
int launch_blur(const uint8_t* img, uint8_t* out, const AwxCorruptParams* dparams, const unsigned* mask, int64_t B, int H,
                int W, int WW, const NormOut& norm, cudaStream_t s) {
  auto kern = blur_kernel<R, RAIN>;
  AWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)blur_smem<R>()));
  dim3 grid((W + kTileW - 1) / kTileW, (H + kTileH - 1) / kTileH, (unsigned)B);
  kern<<<grid, kBlurThreads, blur_smem<R>(), s>>>(img, out, dparams, mask, H, W, WW, norm);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

// ------------------------------------------------------------- rain / snow, row-walking strip kernel
// blur_strip.cuh has the design and the per-thread code (shared with the host emulation of the tests).  This is the
// CTA: strip = blockIdx.x (32 or 27 units), row segment = blockIdx.y, image = blockIdx.z.  One barrier per
// iteration (the filtered rows are double buffered).  The bytes of the NEXT iteration's row travel global -> shared
// with cp.async (LDGSTS: no register staging) into a lane-private 80-byte slot, issued right after the H phase has
// consumed the current ones: they are in flight across the barrier and the whole V phase without holding 18
// registers per thread, which the V phase's register window needs (R = 3: 56 of the 128 a thread may have).
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_smem, const void* src, bool valid) {  // zero fill when !valid
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst_smem), "l"(src), "r"(valid ? 4 : 0) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
constexpr int kRawSlotBytes = 80;  // 3 own chunks + the 16 bytes before + the 16 bytes after; 5 x 16 B at a lane
                                   // stride of 5 chunks: conflict free
constexpr int kRawMaskBytes = 8;   // + the two mask words around the unit, in an array of their own
template <int R, bool RAIN>
__global__ void __launch_bounds__(strip::Geo<R>::kThreads, strip::Geo<R>::kCtasPerSm)
    blur_strip_kernel(const uint8_t* __restrict__ img, uint8_t* __restrict__ out, const AwxCorruptParams* __restrict__ params,
                      const unsigned* __restrict__ mask, int H, int W, int WW, int seg, float negzero) {
  using G = strip::Geo<R>;
  constexpr int NR = G::kNR;
  const int b = blockIdx.z;
  const AwxCorruptParams prm = params[b];
  if (prm.kind != (RAIN ? AWX_RAIN : AWX_SNOW) || prm.blur_k != 2 * R + 1) return;

  extern __shared__ __align__(16) unsigned char smem[];
  float4* s_h = reinterpret_cast<float4*>(smem);  // [G::kBuffers][NR][G::kRowFloat4]
  uint4* s_raw = reinterpret_cast<uint4*>(smem + G::kSmemBytes) + threadIdx.x * (kRawSlotBytes / 16);  // this lane's slot
  const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(s_raw);
  uint2* s_mask = reinterpret_cast<uint2*>(smem + G::kSmemBytes + G::kThreads * kRawSlotBytes) + threadIdx.x;
  const uint32_t mask_addr = (uint32_t)__cvta_generic_to_shared(s_mask);

  const int NU = W / strip::kUnitPx;
  // H-phase task of this thread: row `hrl` of every iteration, unit `hul` of the strip
  const int hrl = threadIdx.x / G::kStripUnits, hul = threadIdx.x - hrl * G::kStripUnits;
  const int unit = blockIdx.x * G::kStripUnits + hul;
  const bool hact = hrl < NR && unit < NU;
  const int ys = blockIdx.y * seg, ye = min(ys + seg, H);
  const int nH = ye - ys + 2 * R;  // filtered rows this CTA produces: ys - R .. ye + R - 1 (reflected)
  const size_t row_bytes = (size_t)W * 3;
  const uint8_t* src = img + (size_t)b * H * row_bytes;
  uint8_t* dst = out + (size_t)b * H * row_bytes;
  const unsigned* m = mask + (size_t)b * H * WW;

  strip::PointParams pp;
  pp.k1 = prm.f0;
  pp.k2 = prm.f1;
  pp.t[0] = prm.taps[0];
  pp.t[1] = prm.taps[1];
  pp.t[2] = prm.taps[2];
  pp.t[3] = prm.taps[3];
  pp.negzero = negzero;

  // issue the copies of filtered row `hr` (nothing here waits on memory)
  auto prefetch = [&](int hr) {
    const int sy = strip::reflect101(ys - R + hr, H);
    const uint8_t* g = src + (size_t)sy * row_bytes + (size_t)unit * strip::kUnitE;
    cp_async16(raw_addr, g);
    cp_async16(raw_addr + 16, g + 16);
    cp_async16(raw_addr + 32, g + 32);
    if (unit > 0) cp_async16(raw_addr + 48, g - 16);        // else: reflect_left fills the halo
    if (unit + 1 < NU) cp_async16(raw_addr + 64, g + 48);   // else: reflect_right does
    // the two mask words around the unit (see strip::mask_words), zero filled outside the row
    const unsigned* mrow = m + (size_t)sy * WW;
    const int wi = unit >> 1, w0 = (unit & 1) ? wi : wi - 1;
    cp_async4(mask_addr, mrow + (w0 >= 0 ? w0 : 0), w0 >= 0);
    cp_async4(mask_addr + 4, mrow + (w0 + 1 < WW ? w0 + 1 : 0), w0 + 1 < WW);
    cp_async_commit();
  };
  // ... and collect them at the start of the H phase
  auto collect = [&](strip::Raw& r) {
    cp_async_wait_all();
    const uint4 c0 = s_raw[0], c1 = s_raw[1], c2 = s_raw[2], l = s_raw[3], rr = s_raw[4];
    r.own[0] = c0.x, r.own[1] = c0.y, r.own[2] = c0.z, r.own[3] = c0.w;
    r.own[4] = c1.x, r.own[5] = c1.y, r.own[6] = c1.z, r.own[7] = c1.w;
    r.own[8] = c2.x, r.own[9] = c2.y, r.own[10] = c2.z, r.own[11] = c2.w;
    r.hl[0] = l.y, r.hl[1] = l.z, r.hl[2] = l.w;
    r.hr[0] = rr.x, r.hr[1] = rr.y, r.hr[2] = rr.z;
    const uint2 mw = *s_mask;
    r.m0 = mw.x, r.m1 = mw.y;
    strip::finish_raw(r, unit, NU);
  };

  // V-phase ownership
  const int g = threadIdx.x;
  const int vu = g / 6, vkg = g - vu * 6;
  const int vunit = blockIdx.x * G::kStripUnits + vu;
  const bool vact = g < G::kGroups && vunit < NU;
  const int vslot0 = G::group_slot(g < G::kGroups ? g : 0, 0), vslot1 = G::group_slot(g < G::kGroups ? g : 0, 1);
  // the thread's output pointer walks down the rows: row ys - 2R (virtual) at the first filtered row
  uint8_t* vdst = dst + (size_t)vunit * strip::kUnitE + vkg * 4 + ((ptrdiff_t)ys - 2 * R) * (ptrdiff_t)row_bytes;

  strip::Raw raw;
  if (hact && hrl < nH) prefetch(hrl);
  strip::Window<R> win;
  constexpr unsigned kOvBits = ((1u << (strip::kUnitPx + 2 * R)) - 1u) << (8 - R);  // the pixels the filter can reach

  const int iters = (nH + NR - 1) / NR;
  int buf = 0;
  for (int it = 0; it < iters; ++it) {
    const int hr = it * NR + hrl;
    const bool hrow = hact && hr < nH;
    float4* rowbuf = s_h + (size_t)(buf * NR + (hrl < NR ? hrl : 0)) * G::kRowFloat4;
    auto store = [&](int q, const strip::F2& a, const strip::F2& c) {
      rowbuf[G::quad_slot(hul, q)] = make_float4(a.x, a.y, c.x, c.y);
    };
    // overlays are rare per pixel but not per warp: one vote picks the instruction stream with or without selects
    if (hrow) collect(raw);
    const bool any_ov = __any_sync(0xffffffffu, hrow && (raw.mb & kOvBits) != 0u);
    if (hrow) {
      if (any_ov)
        strip::h_row<R, RAIN, true>(raw, pp, store);
      else
        strip::h_row<R, RAIN, false>(raw, pp, store);
    }
    if (hact && hr + NR < nH) prefetch(hr + NR);
    __syncthreads();
    if (vact) {
      const float4* vb = s_h + (size_t)(buf * NR) * G::kRowFloat4;
#define AWX_VROW(J)                                                                                          \
  if constexpr ((J) < NR) {                                                                                  \
    const int hj = it * NR + (J);                                                                            \
    if (hj < nH) {                                                                                           \
      const float4 a = vb[(J) * G::kRowFloat4 + vslot0], c = vb[(J) * G::kRowFloat4 + vslot1];               \
      unsigned wlo, whi;                                                                                     \
      const bool emit = hj >= 2 * R;                                                                         \
      strip::v_row<R, (J)>(win, strip::Quad{a.x, a.y, a.z, a.w}, strip::Quad{c.x, c.y, c.z, c.w}, pp.t, emit, wlo, whi); \
      if (emit) {                                                                                            \
        *reinterpret_cast<unsigned*>(vdst) = wlo;                                                            \
        *reinterpret_cast<unsigned*>(vdst + 24) = whi;                                                       \
      }                                                                                                      \
      vdst += row_bytes;                                                                                     \
    }                                                                                                        \
  }
      AWX_VROW(0) AWX_VROW(1) AWX_VROW(2) AWX_VROW(3) AWX_VROW(4) AWX_VROW(5) AWX_VROW(6)
#undef AWX_VROW
    }
    if (G::kBuffers == 2)
      buf ^= 1;
    else
      __syncthreads();  // single buffer: the next H phase overwrites the rows this V phase read
  }
}

template <int R, bool RAIN>
int launch_blur_strip(const uint8_t* img, uint8_t* out, const AwxCorruptParams* dparams, const unsigned* mask, int64_t B,
                      int64_t n_match, int H, int W, int WW, cudaStream_t s) {
  using G = strip::Geo<R>;
  auto kern = blur_strip_kernel<R, RAIN>;
  constexpr int kSmem = G::kSmemBytes + G::kThreads * (kRawSlotBytes + kRawMaskBytes);
  AWX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
  const int strips = (W / strip::kUnitPx + G::kStripUnits - 1) / G::kStripUnits;
  // Row segments: a CTA filters seg + 2R rows to emit seg, and the launch runs in waves of kCtasPerSm CTAs per SM; pick the
  // segment count that minimises waves x rows per CTA (long segments amortise the vertical halo, short ones fill the
  // last wave), between 32 and 128 rows per segment.
  const long long slots = (long long)G::kCtasPerSm * sm_count();
  const long long lo = (H + 127) / 128, hi = H / 32 > 0 ? H / 32 : 1;
  long long best = lo, best_cost = -1;
  for (long long n = lo; n <= hi; ++n) {
    const long long sg = (H + n - 1) / n, ctas = (long long)strips * ((H + sg - 1) / sg) * n_match;
    const long long cost = ((ctas + slots - 1) / slots) * (sg + 2 * R + 4);
    if (best_cost < 0 || cost < best_cost) best = n, best_cost = cost;
  }
  const int seg = (int)((H + best - 1) / best);
  dim3 grid((unsigned)strips, (unsigned)((H + seg - 1) / seg), (unsigned)B);
  kern<<<grid, G::kThreads, kSmem, s>>>(img, out, dparams, mask, H, W, WW, seg, -0.0f);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

// the strip kernel needs whole 16-pixel units, 16-byte aligned rows and a uint8 output; AWX_BLUR_KERNEL=tile forces the
// tile kernel (A/B measurements, parity tests of both)
bool strip_kernel_ok(const uint8_t* img, const uint8_t* out, const NormOut& norm, int H, int W) {
  const char* e = getenv("AWX_BLUR_KERNEL");  // read per call: the parity tests switch kernels within one process
  const bool forced_tile = e && e[0] == 't';
  return !forced_tile && out != nullptr && norm.ptr == nullptr && W % strip::kUnitPx == 0 && H >= 1 &&
         (((uintptr_t)img | (uintptr_t)out) & 15) == 0;
}

size_t params_bytes(int64_t B) { return ((size_t)B * sizeof(AwxCorruptParams) + 255) & ~(size_t)255; }

// ------------------------------------------------------------------------ synthetic depth
// scipy.ndimage.gaussian_filter: correlate1d along axis 0 then axis 1, mode='reflect' (edge
// sample repeated), symmetric accumulation x0*w0 + sum_{j=-r..-1} (x[j] + x[-j]) * w[j], fp64.
__device__ __forceinline__ int reflect_dup(int i, int n) {
  while (i < 0 || i >= n) i = i < 0 ? -i - 1 : 2 * n - 1 - i;
  return i;
}

struct GaussTaps {
  double w[33];  // w[0] = centre, w[j] = tap at distance j
  int radius;
};

template <bool VERTICAL, typename OT>
__global__ void __launch_bounds__(256) depth_pass_kernel(const double* __restrict__ in, OT* __restrict__ out, int H, int W,
                                                          long long total, const GaussTaps taps, double scale, double floor_value) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % W);
    const long long t = i / W;
    const int y = (int)(t % H);
    const double* plane = in + (t / H) * (long long)H * W;
    auto sample = [&](int yy, int xx) -> double {
      const double v = plane[(long long)yy * W + xx];
      // first pass reads noise and adds the vertical ramp (y/H)*scale on the fly
      return VERTICAL ? __dadd_rn(__dmul_rn(__ddiv_rn((double)yy, (double)H), scale), v) : v;
    };
    double acc = __dmul_rn(VERTICAL ? sample(y, x) : sample(y, x), taps.w[0]);
    for (int j = taps.radius; j >= 1; --j) {
      double lo, hi;
      if (VERTICAL) {
        lo = sample(reflect_dup(y - j, H), x);
        hi = sample(reflect_dup(y + j, H), x);
      } else {
        lo = sample(y, reflect_dup(x - j, W));
        hi = sample(y, reflect_dup(x + j, W));
      }
      acc = __dadd_rn(acc, __dmul_rn(__dadd_rn(lo, hi), taps.w[j]));
    }
    if (!VERTICAL) acc = fmax(acc, floor_value);  // np.maximum(depth, 1.0) for the synthetic depth; -inf: none
    out[i] = (OT)acc;
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" size_t awx_corrupt_workspace_bytes(int64_t batch, int32_t height, int32_t width) {
  if (batch <= 0 || height <= 0 || width <= 0) return 0;
  const size_t ww = (size_t)(width + 31) / 32;
  return params_bytes(batch) + (size_t)batch * height * ww * sizeof(unsigned);
}

namespace {
int corrupt_impl(const uint8_t* img, uint8_t* out, const NormOut& norm, int64_t batch, int32_t H, int32_t W,
                 const AwxCorruptParams* params, const void* field, int32_t field_dtype, const int32_t* items,
                 int64_t n_items, void* workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_corrupt: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(img && (out || norm.ptr) && params && workspace, AWX_E_ARG, "awx_corrupt: NULL pointer (img/out/params/workspace)");
  AWX_REQUIRE(batch <= 65535, AWX_E_UNSUPPORTED, "awx_corrupt: batch %lld > 65535 per call", (long long)batch);
  AWX_REQUIRE(field_dtype == AWX_F32 || field_dtype == AWX_F64, AWX_E_ARG, "awx_corrupt: unknown field dtype %d", field_dtype);
  // fog with fp64 depths and in-range coefficients goes to the screened kernel (AWX_FOG_KERNEL=exact: the generic
  // pointwise kernel evaluates the fp64 expression for every pixel -- A/B measurements, parity tests of both)
  const char* fog_env = getenv("AWX_FOG_KERNEL");
  const bool fog_fast = !(fog_env && fog_env[0] == 'e') && field_dtype == AWX_F64 && out != nullptr && norm.ptr == nullptr &&
                        fog_kernel_ok(img, out, field, (long long)H * W);
  const bool night_fast = field_dtype == AWX_F64 && out != nullptr && norm.ptr == nullptr &&
                          night_kernel_ok(img, out, field, (long long)H * W);
  bool any_fog_fast = false, any_night_fast = false;
  bool any_point = false, any_overlay = false;
  int64_t blur[4] = {0, 0, 0, 0};  // images per (kind, blur size): rain 3 / 7, snow 3 / 7
  for (int64_t b = 0; b < batch; ++b) {
    const AwxCorruptParams& q = params[b];
    switch (q.kind) {
      case AWX_CLEAN: any_point = true; break;
      case AWX_FOG:
      case AWX_NIGHT:
        if (q.kind == AWX_FOG && fog_fast && q.field_offset % 2 == 0 && fog::params_ok(q.d0, q.d1))
          any_fog_fast = true;
        else if (q.kind == AWX_NIGHT && night_fast && q.field_offset % 2 == 0)
          any_night_fast = true;
        else
          any_point = true;
        AWX_REQUIRE(field != nullptr, AWX_E_ARG, "awx_corrupt: image %lld (fog/night) needs a depth/noise field", (long long)b);
        AWX_REQUIRE(q.field_offset >= 0, AWX_E_ARG, "awx_corrupt: negative field offset");
        break;
      case AWX_RAIN:
      case AWX_SNOW:
        any_overlay = true;
        AWX_REQUIRE(q.blur_k == 3 || q.blur_k == 7, AWX_E_UNSUPPORTED, "awx_corrupt: blur kernel %d (3 or 7 supported)", q.blur_k);
        AWX_REQUIRE(q.item_count >= 0 && q.item_begin >= 0 && (int64_t)q.item_begin + q.item_count <= n_items, AWX_E_ARG,
                    "awx_corrupt: image %lld item range [%d,+%d) outside %lld items", (long long)b, q.item_begin, q.item_count, (long long)n_items);
        AWX_REQUIRE(q.item_count == 0 || items != nullptr, AWX_E_ARG, "awx_corrupt: items is NULL");
        ++blur[(q.kind == AWX_RAIN ? 0 : 2) + (q.blur_k == 3 ? 0 : 1)];
        break;
      default:
        set_error("awx_corrupt: unknown kind %d for image %lld", q.kind, (long long)b);
        return AWX_E_ARG;
    }
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  AwxCorruptParams* dparams = static_cast<AwxCorruptParams*>(workspace);
  AWX_CUDA(cudaMemcpyAsync(dparams, params, (size_t)batch * sizeof(AwxCorruptParams), cudaMemcpyHostToDevice, s));
  const long long HW = (long long)H * W;
  if (any_point) {
    long long chunks = (HW + kChunkPx - 1) / kChunkPx;
    const long long cap = (long long)sm_count() * 16;
    dim3 grid((unsigned)(chunks < cap ? chunks : cap), (unsigned)batch);
    if (field_dtype == AWX_F64)
      pointwise_kernel<double><<<grid, kPointThreads, 0, s>>>(img, out, dparams, static_cast<const double*>(field), HW, norm,
                                                              (fog_fast ? 1 : 0) | (night_fast ? 2 : 0));
    else
      pointwise_kernel<float><<<grid, kPointThreads, 0, s>>>(img, out, dparams, static_cast<const float*>(field), HW, norm, 0);
    AWX_CUDA(cudaGetLastError());
    note_launch();
  }
  if (any_fog_fast) {
    long long blocks = (HW / 16 + kFogThreads - 1) / kFogThreads;
    const long long cap = (long long)sm_count() * 12;
    dim3 grid((unsigned)(blocks < cap ? blocks : cap), (unsigned)batch);
    fog_kernel<<<grid, kFogThreads, 0, s>>>(img, out, dparams, static_cast<const double*>(field), HW);
    AWX_CUDA(cudaGetLastError());
    note_launch();
  }
  if (any_night_fast) {
    long long blocks = (HW * 3 + kNightBlockValues - 1) / kNightBlockValues;
    const long long cap = (long long)sm_count() * 16;
    dim3 grid((unsigned)(blocks < cap ? blocks : cap), (unsigned)batch);
    night_kernel<<<grid, kNightThreads, 0, s>>>(img, out, dparams, static_cast<const double*>(field), HW);
    AWX_CUDA(cudaGetLastError());
    note_launch();
  }
  if (any_overlay) {
    const int WW = (W + 31) / 32;
    unsigned* mask = reinterpret_cast<unsigned*>(static_cast<unsigned char*>(workspace) + params_bytes(batch));
    AWX_CUDA(cudaMemsetAsync(mask, 0, (size_t)batch * H * WW * sizeof(unsigned), s));
    rasterize_kernel<<<dim3(4, (unsigned)batch), 128, 0, s>>>(dparams, items, mask, H, W, WW);
    AWX_CUDA(cudaGetLastError());
  note_launch();
    int rc = AWX_OK;
    // one launch per (kind, blur size) present in the batch; every CTA of a launch skips the images of the others
    if (strip_kernel_ok(img, out, norm, H, W)) {
      if (blur[0]) rc = launch_blur_strip<1, true>(img, out, dparams, mask, batch, blur[0], H, W, WW, s);
      if (rc == AWX_OK && blur[1]) rc = launch_blur_strip<3, true>(img, out, dparams, mask, batch, blur[1], H, W, WW, s);
      if (rc == AWX_OK && blur[2]) rc = launch_blur_strip<1, false>(img, out, dparams, mask, batch, blur[2], H, W, WW, s);
      if (rc == AWX_OK && blur[3]) rc = launch_blur_strip<3, false>(img, out, dparams, mask, batch, blur[3], H, W, WW, s);
    } else {
      if (blur[0]) rc = launch_blur<1, true>(img, out, dparams, mask, batch, H, W, WW, norm, s);
      if (rc == AWX_OK && blur[1]) rc = launch_blur<3, true>(img, out, dparams, mask, batch, H, W, WW, norm, s);
      if (rc == AWX_OK && blur[2]) rc = launch_blur<1, false>(img, out, dparams, mask, batch, H, W, WW, norm, s);
      if (rc == AWX_OK && blur[3]) rc = launch_blur<3, false>(img, out, dparams, mask, batch, H, W, WW, norm, s);
    }
    if (rc != AWX_OK) return rc;
  }
  return AWX_OK;
}
}  // namespace

extern "C" int awx_corrupt(const uint8_t* img, uint8_t* out, int64_t batch, int32_t H, int32_t W,
                           const AwxCorruptParams* params, const void* field, int32_t field_dtype, const int32_t* items,
                           int64_t n_items, void* workspace, void* stream) {
  AWX_REQUIRE(batch == 0 || H == 0 || W == 0 || out != nullptr, AWX_E_ARG, "awx_corrupt: out is NULL");
  NormOut norm{};
  return corrupt_impl(img, out, norm, batch, H, W, params, field, field_dtype, items, n_items, workspace, stream);
}

extern "C" int awx_corrupt_normalized(const uint8_t* img, uint8_t* out, void* norm_out, int32_t norm_dtype,
                                      const float* mean255, const float* rdenom, int64_t batch, int32_t H, int32_t W,
                                      const AwxCorruptParams* params, const void* field, int32_t field_dtype,
                                      const int32_t* items, int64_t n_items, void* workspace, void* stream) {
  AWX_REQUIRE(batch == 0 || H == 0 || W == 0 || (norm_out && mean255 && rdenom), AWX_E_ARG,
              "awx_corrupt_normalized: NULL pointer (norm_out/mean255/rdenom)");
  AWX_REQUIRE(norm_dtype == AWX_F32 || norm_dtype == AWX_BF16, AWX_E_ARG, "awx_corrupt_normalized: norm dtype must be AWX_F32 or AWX_BF16");
  NormOut norm{};
  norm.ptr = norm_out;
  norm.bf16 = norm_dtype == AWX_BF16;
  for (int c = 0; c < 3 && mean255 && rdenom; ++c) {
    norm.mean[c] = mean255[c];
    norm.rden[c] = rdenom[c];
  }
  return corrupt_impl(img, out, norm, batch, H, W, params, field, field_dtype, items, n_items, workspace, stream);
}

namespace awx {
// scipy.ndimage.gaussian_filter (fp64, mode='reflect') of [B,H,W] planes: vertical pass (adding the ramp
// (y/H)*ramp_scale on the fly) into tmp, horizontal pass into out, then max(., floor_value).
int launch_gauss_f64(const double* in, void* out, int32_t out_dtype, double* tmp, int64_t batch, int32_t H, int32_t W,
                     double ramp_scale, double floor_value, const double* weights, int32_t radius, cudaStream_t s) {
  GaussTaps taps{};
  taps.radius = radius;
  for (int j = 0; j <= radius; ++j) taps.w[j] = weights[radius + j];
  const long long total = (long long)batch * H * W;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  depth_pass_kernel<true, double><<<(unsigned)blocks, 256, 0, s>>>(in, tmp, H, W, total, taps, ramp_scale, floor_value);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  if (out_dtype == AWX_F64)
    depth_pass_kernel<false, double><<<(unsigned)blocks, 256, 0, s>>>(tmp, static_cast<double*>(out), H, W, total, taps, ramp_scale, floor_value);
  else
    depth_pass_kernel<false, float><<<(unsigned)blocks, 256, 0, s>>>(tmp, static_cast<float*>(out), H, W, total, taps, ramp_scale, floor_value);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}
}  // namespace awx

extern "C" int awx_synth_depth(const double* noise, void* out, int32_t out_dtype, double* tmp, int64_t batch, int32_t H,
                               int32_t W, double depth_scale, const double* weights, int32_t radius, void* stream) {
  AWX_REQUIRE(batch >= 0 && H >= 0 && W >= 0, AWX_E_ARG, "awx_synth_depth: negative size");
  if (batch == 0 || H == 0 || W == 0) return AWX_OK;
  AWX_REQUIRE(noise && out && tmp && weights, AWX_E_ARG, "awx_synth_depth: NULL pointer");
  AWX_REQUIRE(radius >= 0 && radius <= 32, AWX_E_UNSUPPORTED, "awx_synth_depth: radius %d outside 0..32", radius);
  AWX_REQUIRE(out_dtype == AWX_F32 || out_dtype == AWX_F64, AWX_E_ARG, "awx_synth_depth: unknown out dtype");
  return launch_gauss_f64(noise, out, out_dtype, tmp, batch, H, W, depth_scale, 1.0, weights, radius,
                          static_cast<cudaStream_t>(stream));
}
