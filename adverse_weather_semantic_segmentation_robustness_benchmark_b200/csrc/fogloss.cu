// awx_fogloss: fog-density-aware pixel loss, forward and gradients in one pass over HBM.
//
//   loss_i  = ce_i * (1 + s * fog_i)            ce_i = logsumexp(x_i) - x_i[y_i]   (or focal(ce_i))
//   sums[0] = sum_i loss_i        sums[1] = sum_i (depth_pred_i - depth_tgt_i)^2
//   dlogits[i,c] = (1 + s*fog_i) * g(ce_i) * (softmax_c - [c == y_i]) / N      g = 1 (CE) or focal'
//   ddepth[i]    = 2 * (depth_pred_i - depth_tgt_i) / N
//   dfog[i]      = s * base_loss_i / N     (needed when fog density itself depends on the depth head)
// (FogDensityAwareLoss.forward, models/model.py:577-611; autograd of the same expressions).
// Two kernels.  fogloss_kernel (any C, any alignment): load pattern as in score.cu, a thread owns PX consecutive
// pixels, 64-bit loads per class plane, everything in registers; gradients are stored with the same pattern.
// fogloss_ring_kernel (C == 19, 16-byte aligned planes): the score kernel's TMA ring -- one persistent CTA per SM,
// a producer warp streams 608-pixel tiles of the 19 planes into shared memory with bulk async copies while 19
// consumer warps (one pixel per thread) work on the tiles that have landed, so HBM reads never wait for the
// arithmetic or for the gradient stores (the register kernel is latency bound: 16 warps per SM, every load phase
// exposed).  Same per-pixel arithmetic in both.
// Sums: fp64 per thread -> warp shuffle -> CTA -> one partial per CTA, reduced in CTA order by a
// second single-block kernel, so the result is bit-reproducible for a given device.
#include "awx_internal.cuh"
#include "tma_ring.cuh"

namespace awx {
namespace {

constexpr int kThreads = 256;
constexpr int kMaxBlocks = 4096;

struct LossParams {
  const float* logits;
  const void* labels;
  int label_mode;
  const float* fog;
  const float* dpred;
  const float* dtgt;
  float sens;
  int focal;
  long long B, HW;
  int C;
  float inv_n;
  float* dlogits;
  float* ddepth;
  float* dfog;                  // d(sum loss)/d fog_i / N = s * base_loss_i / N (path B of the reference)
  double* partials;             // [gridDim.x][2]
  unsigned long long* bad;      // labels outside [0,C): torch raises; counted, loss/grad contribution 0
};

template <int CS, int PX>
__global__ void __launch_bounds__(kThreads, CS > 0 ? 2 : 1) fogloss_kernel(const __grid_constant__ LossParams p) {
  constexpr int CA = CS > 0 ? CS : AWX_MAX_CLASSES;
  const int C = CS > 0 ? CS : p.C;
  const long long HW = p.HW;
  const long long gpi = HW / PX;
  const long long total = p.B * gpi;
  double acc_seg = 0.0, acc_depth = 0.0;
  unsigned n_bad = 0;
  for (long long g = (long long)blockIdx.x * kThreads + threadIdx.x; g < total; g += (long long)gridDim.x * kThreads) {
    const long long img = g / gpi;
    const long long px = (g - img * gpi) * PX;
    const float* gl = p.logits + img * C * HW + px;
    float x[PX][CA];
#pragma unroll
    for (int c = 0; c < C; ++c) {
      if (PX == 2) {
        const float2 t = ld_stream2(gl + c * HW);
        x[0][c] = t.x;
        x[PX - 1][c] = t.y;
      } else {
        x[0][c] = ld_stream(gl + c * HW);
      }
    }
    const long long li = img * HW + px;
#pragma unroll
    for (int j = 0; j < PX; ++j) {
      long long y;
      if (p.label_mode == AWX_LABEL_U8)
        y = static_cast<const uint8_t*>(p.labels)[li + j];
      else
        y = static_cast<const long long*>(p.labels)[li + j];
      const bool ok = y >= 0 && y < C;
      n_bad += !ok;
      float mx = x[j][0];
#pragma unroll
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, x[j][c]);
      float s = 0.f, xy = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float d = x[j][c] - mx;
        if (c == (int)y) xy = d;
        const float e = ex2_approx(d * kLog2e);
        s += e;
        x[j][c] = e;
      }
      const float ce = kLn2 * lg2_approx(s) - xy;
      float w = 1.0f;
      if (p.fog) w = fmaf(p.sens, p.fog[li + j], 1.0f);
      float lossv = ce, gscale = 1.0f;
      if (p.focal) {
        const float pt = ex2_approx(-ce * kLog2e);
        const float om = 1.0f - pt;
        lossv = om * om * ce;
        gscale = om * om + 2.0f * om * pt * ce;
      }
      if (ok) acc_seg += (double)(lossv * w);
      if (p.dfog) p.dfog[li + j] = ok ? p.sens * lossv * p.inv_n : 0.f;
      if (p.dlogits) {
        const float k = ok ? w * gscale * p.inv_n : 0.f;
        const float r = __frcp_rn(s);
        float* go = p.dlogits + img * C * HW + px + j;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const float prob = x[j][c] * r;
          x[j][c] = k * (prob - (c == (int)y ? 1.0f : 0.0f));
        }
        if (PX == 1) {
#pragma unroll
          for (int c = 0; c < C; ++c) go[c * HW] = x[0][c];
        }
      }
      if (p.dpred && p.dtgt) {
        const float diff = p.dpred[li + j] - p.dtgt[li + j];
        acc_depth += (double)(diff * diff);
        if (p.ddepth) p.ddepth[li + j] = 2.0f * diff * p.inv_n;
      }
    }
    if (PX == 2 && p.dlogits) {
      float* go = p.dlogits + img * C * HW + px;
#pragma unroll
      for (int c = 0; c < C; ++c) *reinterpret_cast<float2*>(go + c * HW) = make_float2(x[0][c], x[PX - 1][c]);
    }
  }
  // ---- CTA reduction (fixed order) -> one partial per CTA
  __shared__ double s_red[2][kThreads / 32];
  __shared__ unsigned s_bad;
  if (threadIdx.x == 0) s_bad = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_seg += __shfl_down_sync(0xffffffffu, acc_seg, o);
    acc_depth += __shfl_down_sync(0xffffffffu, acc_depth, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = acc_seg;
    s_red[1][threadIdx.x >> 5] = acc_depth;
  }
  const unsigned wb = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0 && wb) atomicAdd(&s_bad, wb);
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) {
      a += s_red[0][w];
      d += s_red[1][w];
    }
    p.partials[2 * blockIdx.x] = a;
    p.partials[2 * blockIdx.x + 1] = d;
    if (s_bad && p.bad) atomicAdd(p.bad, (unsigned long long)s_bad);
  }
}

__global__ void fogloss_finish_kernel(const double* __restrict__ partials, int n, double* __restrict__ sums) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int i = 0; i < n; ++i) {
      a += partials[2 * i];
      d += partials[2 * i + 1];
    }
    sums[0] += a;
    sums[1] += d;
  }
}

template <int CS, int PX>
int launch_loss(LossParams& p, double* sums, cudaStream_t s) {
  auto kern = fogloss_kernel<CS, PX>;
  int occ = 0;
  AWX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kThreads, 0));
  if (occ < 1) occ = 1;
  const long long groups = p.B * (p.HW / PX);
  long long blocks = (groups + kThreads - 1) / kThreads;
  long long cap = (long long)sm_count() * occ;
  if (cap > kMaxBlocks) cap = kMaxBlocks;
  if (blocks > cap) blocks = cap;
  kern<<<(unsigned)blocks, kThreads, 0, s>>>(p);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  fogloss_finish_kernel<<<1, 32, 0, s>>>(p.partials, (int)blocks, sums);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

// ------------------------------------------------------------------------------------------- TMA-staged kernel
constexpr int kRC = 19;                           // classes
#ifndef AWX_LOSS_RING_WARPS
#define AWX_LOSS_RING_WARPS 19
#endif
#ifndef AWX_LOSS_RING_UNITS
#define AWX_LOSS_RING_UNITS 4
#endif
constexpr int kRingWarps = AWX_LOSS_RING_WARPS;   // consumer warps
constexpr int kRingTile = 32 * kRingWarps;        // pixels per tile
constexpr int kRingThreads = kRingTile + 32;      // + producer warp
constexpr int kRingUnits = AWX_LOSS_RING_UNITS;   // ring depth (tiles)
constexpr int kRingUnitBytes = kRC * kRingTile * 4;
constexpr size_t kRingSmem = 128 + (size_t)kRingUnits * kRingUnitBytes;  // [full[4] | empty[4] | pad | units]

__global__ void __launch_bounds__(kRingThreads, 1) fogloss_ring_kernel(const __grid_constant__ LossParams p) {
  extern __shared__ __align__(128) unsigned char smem[];
  u64* full = reinterpret_cast<u64*>(smem);
  u64* empty = full + kRingUnits;
  float* units = reinterpret_cast<float*>(smem + 128);
  if (threadIdx.x == 0) {
    for (int u = 0; u < kRingUnits; ++u) {
      mbar_init(full + u, 1);
      mbar_init(empty + u, kRingWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  const long long HW = p.HW;
  const long long tpi = (HW + kRingTile - 1) / kRingTile;  // tiles per image
  const long long ntiles = p.B * tpi;
  double acc_seg = 0.0, acc_depth = 0.0;
  unsigned n_bad = 0;

  if (threadIdx.x >= kRingTile) {
    // ------------------------------------------------------------------ producer warp
    unsigned u = 0, ph = 0;
    long long img = blockIdx.x / tpi, tin = blockIdx.x - img * tpi;  // image and tile-in-image, advanced incrementally
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long p0 = tin * kRingTile;
      const unsigned npx = (unsigned)((HW - p0) < kRingTile ? (HW - p0) : kRingTile);
      mbar_wait(empty + u, ph ^ 1u);
      if (elect_one()) {  // one thread, warp-uniform operands: the copies stay on the uniform datapath
        mbar_expect_tx(full + u, kRC * npx * 4u);
        const float* src = p.logits + img * kRC * HW + p0;
        float* dst = units + (size_t)u * (kRC * kRingTile);
#pragma unroll
        for (int c = 0; c < kRC; ++c) bulk_load(dst + c * kRingTile, src + c * HW, npx * 4u, full + u);
      }
      __syncwarp();
      if (++u == (unsigned)kRingUnits) {
        u = 0;
        ph ^= 1u;
      }
      tin += gridDim.x;
      while (tin >= tpi) {
        tin -= tpi;
        ++img;
      }
    }
  } else {
    // ------------------------------------------------------------------ consumer warps: thread t owns pixel t
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t my0 = sbase + 128u + 4u * (uint32_t)t;
    unsigned u = 0, ph = 0;
    long long img_next = blockIdx.x / tpi, tin_next = blockIdx.x - img_next * tpi;
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const long long img = img_next, p0 = tin_next * kRingTile;
      tin_next += gridDim.x;
      while (tin_next >= tpi) {
        tin_next -= tpi;
        ++img_next;
      }
      const bool act = p0 + t < HW;  // false only in the tail tile of an image (stale ring contents, nothing stored)
      const long long li = img * HW + p0 + t;
      // side inputs first: their latency overlaps the wait for the tile
      long long y = 0;
      float fogv = 0.f, dp = 0.f, dt = 0.f;
      if (act) {
        y = p.label_mode == AWX_LABEL_U8 ? (long long)static_cast<const uint8_t*>(p.labels)[li]
                                         : static_cast<const long long*>(p.labels)[li];
        if (p.fog) fogv = p.fog[li];
        if (p.dpred && p.dtgt) {
          dp = p.dpred[li];
          dt = p.dtgt[li];
        }
      }
      float x[kRC];
      mbar_wait_a(sbase + 8u * u, ph);
      const uint32_t s0 = my0 + u * (uint32_t)kRingUnitBytes;
#pragma unroll
      for (int c = 0; c < kRC; ++c) x[c] = lds_f32(s0 + c * kRingTile * 4);
      __syncwarp();
      if (lane == 0) mbar_arrive_a(sbase + 8u * (kRingUnits + u));
      if (++u == (unsigned)kRingUnits) {
        u = 0;
        ph ^= 1u;
      }
      if (!act) continue;
      // ---- the per-pixel arithmetic of fogloss_kernel, unchanged
      const bool ok = y >= 0 && y < kRC;
      n_bad += !ok;
      float mx = x[0];
#pragma unroll
      for (int c = 1; c < kRC; ++c) mx = fmaxf(mx, x[c]);
      float sum = 0.f, xy = 0.f;
#pragma unroll
      for (int c = 0; c < kRC; ++c) {
        const float d = x[c] - mx;
        if (c == (int)y) xy = d;
        const float e = ex2_approx(d * kLog2e);
        sum += e;
        x[c] = e;
      }
      const float ce = kLn2 * lg2_approx(sum) - xy;
      float w = 1.0f;
      if (p.fog) w = fmaf(p.sens, fogv, 1.0f);
      float lossv = ce, gscale = 1.0f;
      if (p.focal) {
        const float pt = ex2_approx(-ce * kLog2e);
        const float om = 1.0f - pt;
        lossv = om * om * ce;
        gscale = om * om + 2.0f * om * pt * ce;
      }
      if (ok) acc_seg += (double)(lossv * w);
      if (p.dfog) p.dfog[li] = ok ? p.sens * lossv * p.inv_n : 0.f;
      if (p.dlogits) {
        const float k = ok ? w * gscale * p.inv_n : 0.f;
        const float r = __frcp_rn(sum);
        float* go = p.dlogits + img * kRC * HW + p0 + t;
#pragma unroll
        for (int c = 0; c < kRC; ++c) {
          const float prob = x[c] * r;
          __stcs(go + c * HW, k * (prob - (c == (int)y ? 1.0f : 0.0f)));
        }
      }
      if (p.dpred && p.dtgt) {
        const float diff = dp - dt;
        acc_depth += (double)(diff * diff);
        if (p.ddepth) p.ddepth[li] = 2.0f * diff * p.inv_n;
      }
    }
  }
  // ---- CTA reduction (fixed order) -> one partial per CTA; the producer warp contributes zeros
  __shared__ double s_red[2][kRingThreads / 32];
  __shared__ unsigned s_bad;
  if (threadIdx.x == 0) s_bad = 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    acc_seg += __shfl_down_sync(0xffffffffu, acc_seg, o);
    acc_depth += __shfl_down_sync(0xffffffffu, acc_depth, o);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) {
    s_red[0][threadIdx.x >> 5] = acc_seg;
    s_red[1][threadIdx.x >> 5] = acc_depth;
  }
  const unsigned wb = __reduce_add_sync(0xffffffffu, n_bad);
  if ((threadIdx.x & 31) == 0 && wb) atomicAdd(&s_bad, wb);
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, d = 0.0;
    for (int w = 0; w < kRingThreads / 32; ++w) {
      a += s_red[0][w];
      d += s_red[1][w];
    }
    p.partials[2 * blockIdx.x] = a;
    p.partials[2 * blockIdx.x + 1] = d;
    if (s_bad && p.bad) atomicAdd(p.bad, (unsigned long long)s_bad);
  }
}

int launch_loss_ring(LossParams& p, double* sums, cudaStream_t s) {
  AWX_CUDA(cudaFuncSetAttribute(fogloss_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kRingSmem));
  const long long ntiles = p.B * ((p.HW + kRingTile - 1) / kRingTile);
  long long blocks = sm_count();
  if (blocks > ntiles) blocks = ntiles;
  fogloss_ring_kernel<<<(unsigned)blocks, kRingThreads, kRingSmem, s>>>(p);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  fogloss_finish_kernel<<<1, 32, 0, s>>>(p.partials, (int)blocks, sums);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}

__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, long long n, const float* __restrict__ scale) {
  const float k = *scale;
  const long long n4 = n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  const bool aligned = ((uintptr_t)x & 15) == 0;
  if (aligned) {
    float4* x4 = reinterpret_cast<float4*>(x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
      float4 v = x4[i];
      v.x *= k; v.y *= k; v.z *= k; v.w *= k;
      x4[i] = v;
    }
    for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= k;
  } else {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= k;
  }
}

}  // namespace
}  // namespace awx

using namespace awx;

extern "C" size_t awx_fogloss_workspace_bytes(void) { return (size_t)kMaxBlocks * 2 * sizeof(double); }

extern "C" int awx_fogloss(const float* logits, const void* labels, int32_t label_dtype, const float* fog_density,
                           const float* depth_pred, const float* depth_tgt, float fog_sensitivity, int32_t focal,
                           int64_t batch, int32_t C, int64_t pixels_per_image, double* sums, float* dlogits,
                           float* ddepth, float* dfog, int64_t* bad_labels, void* workspace, void* stream) {
  AWX_REQUIRE(batch >= 0 && pixels_per_image >= 0, AWX_E_ARG, "awx_fogloss: negative size");
  AWX_REQUIRE(C >= 1 && C <= AWX_MAX_CLASSES, AWX_E_UNSUPPORTED, "awx_fogloss: num_classes %d outside 1..%d", C, AWX_MAX_CLASSES);
  if (batch == 0 || pixels_per_image == 0) return AWX_OK;
  AWX_REQUIRE(logits && labels && sums && workspace, AWX_E_ARG, "awx_fogloss: NULL pointer (logits/labels/sums/workspace)");
  AWX_REQUIRE(label_dtype == AWX_LABEL_U8 || label_dtype == AWX_LABEL_I64, AWX_E_ARG, "awx_fogloss: unknown label dtype");
  LossParams p{};
  p.logits = logits;
  p.labels = labels;
  p.label_mode = label_dtype;
  p.fog = fog_density;
  p.dpred = depth_pred;
  p.dtgt = depth_tgt;
  p.sens = fog_sensitivity;
  p.focal = focal;
  p.B = batch;
  p.HW = pixels_per_image;
  p.C = C;
  p.inv_n = (float)(1.0 / ((double)batch * (double)pixels_per_image));
  p.dlogits = dlogits;
  p.ddepth = ddepth;
  p.dfog = dfog;
  p.partials = static_cast<double*>(workspace);
  p.bad = reinterpret_cast<unsigned long long*>(bad_labels);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec2 = (pixels_per_image % 2 == 0) && ((uintptr_t)logits & 7) == 0 && ((uintptr_t)dlogits & 7) == 0;
  // TMA ring whenever the bulk copies' alignment rules hold (AWX_LOSS_KERNEL=v1 forces the register kernel: A/B
  // measurements and parity tests of both)
  const char* force = getenv("AWX_LOSS_KERNEL");
  const bool force_v1 = force && force[0] == 'v' && force[1] == '1';
  if (C == 19 && !force_v1 && pixels_per_image % 4 == 0 && ((uintptr_t)logits & 15) == 0) return launch_loss_ring(p, sums, s);
  if (C == 19) return vec2 ? launch_loss<19, 2>(p, sums, s) : launch_loss<19, 1>(p, sums, s);
  return launch_loss<0, 1>(p, sums, s);
}

extern "C" int awx_scale_inplace(float* x, int64_t n, const float* scale, void* stream) {
  AWX_REQUIRE(n >= 0, AWX_E_ARG, "awx_scale_inplace: negative size");
  if (n == 0) return AWX_OK;
  AWX_REQUIRE(x && scale, AWX_E_ARG, "awx_scale_inplace: NULL pointer");
  long long blocks = (n / 4 + 255) / 256 + 1;
  const long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  scale_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, n, scale);
  AWX_CUDA(cudaGetLastError());
  note_launch();
  return AWX_OK;
}
