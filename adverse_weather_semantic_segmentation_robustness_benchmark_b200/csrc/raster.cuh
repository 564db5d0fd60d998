// Scan conversion of the rain streaks and snow flakes, as integer span emitters.
//
// The reference draws with OpenCV (cv2.line thickness 1 or 3, LINE_8; cv2.circle filled;
// data/preprocessing.py:160,194).  OpenCV is a third-party dependency that is not vendored in
// the reference (requirements floor: opencv-python>=4.8; 4.13.0 in this image), so this file
// restates OpenCV's published drawing algorithms (modules/imgproc/src/drawing.cpp: Line /
// LineIterator, Line2, clipLine, FillConvexPoly, Circle, ThickLine) as span emitters; the CPU
// test-suite compares them with cv2 itself over exhaustive placements (tests/test_raster_cpu.py).
//
// All functions are __host__ __device__ and template on an `Emit` callable:
//     emit(int y, int x0, int x1)   -- fill pixels x0..x1 (inclusive, already clipped) of row y
// so the same code fills a global bit mask, a shared-memory tile mask, or a host array in tests.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define AWX_HD __host__ __device__ __forceinline__
#else
#define AWX_HD inline
#endif

namespace awx {
namespace raster {

constexpr int kShift = 16;                 // XY_SHIFT
constexpr long long kOne = 1LL << kShift;  // XY_ONE

struct P64 {
  long long x, y;
};

AWX_HD long long round_half_even(double v) {
  // cvRound: round to nearest, ties to even (lrint in the default rounding mode)
#if defined(__CUDA_ARCH__)
  return __double2ll_rn(v);
#else
  double r = __builtin_rint(v);
  return (long long)r;
#endif
}

// ---------------------------------------------------------------- filled circle (Circle, fill=1)
template <class Emit>
AWX_HD void disc(int cx, int cy, int radius, int W, int H, Emit&& emit) {
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  while (dx >= dy) {
    const int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
    int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
    if (x11 < W && x12 >= 0 && y21 < H && y22 >= 0) {
      x11 = x11 > 0 ? x11 : 0;
      x12 = x12 < W - 1 ? x12 : W - 1;
      if ((unsigned)y11 < (unsigned)H) emit(y11, x11, x12);
      if ((unsigned)y12 < (unsigned)H) emit(y12, x11, x12);
      if (x21 < W && x22 >= 0) {
        x21 = x21 > 0 ? x21 : 0;
        x22 = x22 < W - 1 ? x22 : W - 1;
        if ((unsigned)y21 < (unsigned)H) emit(y21, x21, x22);
        if ((unsigned)y22 < (unsigned)H) emit(y22, x21, x22);
      }
    }
    dy++;
    err += plus;
    plus += 2;
    const int mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}

// ------------------------------------------------------------------ clipLine on 64-bit points
AWX_HD bool clip_line(long long width, long long height, P64& p1, P64& p2) {
  if (width <= 0 || height <= 0) return false;
  const long long right = width - 1, bottom = height - 1;
  long long &x1 = p1.x, &y1 = p1.y, &x2 = p2.x, &y2 = p2.y;
  int c1 = (x1 < 0) + (x1 > right) * 2 + (y1 < 0) * 4 + (y1 > bottom) * 8;
  int c2 = (x2 < 0) + (x2 > right) * 2 + (y2 < 0) * 4 + (y2 > bottom) * 8;
  if ((c1 & c2) == 0 && (c1 | c2) != 0) {
    long long a;
    if (c1 & 12) {
      a = c1 < 8 ? 0 : bottom;
      x1 += (long long)((double)(a - y1) * (x2 - x1) / (y2 - y1));
      y1 = a;
      c1 = (x1 < 0) + (x1 > right) * 2;
    }
    if (c2 & 12) {
      a = c2 < 8 ? 0 : bottom;
      x2 += (long long)((double)(a - y2) * (x2 - x1) / (y2 - y1));
      y2 = a;
      c2 = (x2 < 0) + (x2 > right) * 2;
    }
    if ((c1 & c2) == 0 && (c1 | c2) != 0) {
      if (c1) {
        a = c1 == 1 ? 0 : right;
        y1 += (long long)((double)(a - x1) * (y2 - y1) / (x2 - x1));
        x1 = a;
        c1 = 0;
      }
      if (c2) {
        a = c2 == 1 ? 0 : right;
        y2 += (long long)((double)(a - x2) * (y2 - y1) / (x2 - x1));
        x2 = a;
        c2 = 0;
      }
    }
  }
  return (c1 | c2) == 0;
}

// ------------------------------------------- 8-connected integer line (Line via LineIterator)
template <class Emit>
AWX_HD void line8(int x0, int y0, int x1, int y1, int W, int H, Emit&& emit) {
  P64 a{x0, y0}, b{x1, y1};
  if ((unsigned)x0 >= (unsigned)W || (unsigned)x1 >= (unsigned)W || (unsigned)y0 >= (unsigned)H ||
      (unsigned)y1 >= (unsigned)H) {
    if (!clip_line(W, H, a, b)) return;
  }
  int px = (int)a.x, py = (int)a.y;
  int dx = (int)(b.x - a.x), dy = (int)(b.y - a.y);
  int step_x = 1, step_y = 1;
  if (dx < 0) {  // leftToRight: start from the other end
    dx = -dx;
    dy = -dy;
    px = (int)b.x;
    py = (int)b.y;
  }
  if (dy < 0) {
    dy = -dy;
    step_y = -1;
  }
  const bool vert = dy > dx;
  int major = vert ? dy : dx, minor = vert ? dx : dy;
  int err = major - (minor + minor);
  const int plus_delta = major + major, minus_delta = -(minor + minor);
  const int count = major + 1;
  for (int i = 0; i < count; ++i) {
    emit(py, px, px);
    const int mask = err < 0 ? -1 : 0;
    err += minus_delta + (plus_delta & mask);
    if (vert) {
      py += step_y;
      px += step_x & mask;
    } else {
      px += step_x;
      py += step_y & mask;
    }
  }
}

// ------------------------------------------------ fixed-point outline segment (Line2, 16.16)
template <class Emit>
AWX_HD void line2_fixed(P64 pt1, P64 pt2, int W, int H, Emit&& emit) {
  if (!clip_line((long long)W << kShift, (long long)H << kShift, pt1, pt2)) return;
  long long dx = pt2.x - pt1.x, dy = pt2.y - pt1.y;
  const long long j = dx < 0 ? -1 : 0;
  const long long ax = (dx ^ j) - j;
  const long long i = dy < 0 ? -1 : 0;
  const long long ay = (dy ^ i) - i;
  long long x_step, y_step;
  int ecount;
  if (ax > ay) {
    dy = (dy ^ j) - j;
    pt1.x ^= pt2.x & j;
    pt2.x ^= pt1.x & j;
    pt1.x ^= pt2.x & j;
    pt1.y ^= pt2.y & j;
    pt2.y ^= pt1.y & j;
    pt1.y ^= pt2.y & j;
    x_step = kOne;
    y_step = (dy * kOne) / (ax | 1);
    ecount = (int)((pt2.x - pt1.x) >> kShift);
  } else {
    dx = (dx ^ i) - i;
    pt1.x ^= pt2.x & i;
    pt2.x ^= pt1.x & i;
    pt1.x ^= pt2.x & i;
    pt1.y ^= pt2.y & i;
    pt2.y ^= pt1.y & i;
    pt1.y ^= pt2.y & i;
    x_step = (dx * kOne) / (ay | 1);
    y_step = kOne;
    ecount = (int)((pt2.y - pt1.y) >> kShift);
  }
  pt1.x += (kOne >> 1);
  pt1.y += (kOne >> 1);
  auto put = [&](long long x, long long y) {
    if (0 <= x && x < W && 0 <= y && y < H) emit((int)y, (int)x, (int)x);
  };
  put((pt2.x + (kOne >> 1)) >> kShift, (pt2.y + (kOne >> 1)) >> kShift);
  if (ax > ay) {
    pt1.x >>= kShift;
    while (ecount >= 0) {
      put(pt1.x, pt1.y >> kShift);
      pt1.x++;
      pt1.y += y_step;
      ecount--;
    }
  } else {
    pt1.y >>= kShift;
    while (ecount >= 0) {
      put(pt1.x >> kShift, pt1.y);
      pt1.x += x_step;
      pt1.y++;
      ecount--;
    }
  }
}

// ------------------------------------ convex polygon, vertices in 16.16 (FillConvexPoly, LINE_8)
template <class Emit>
AWX_HD void fill_convex_fixed(const P64* v, int npts, int W, int H, Emit&& emit) {
  struct Edge {
    int idx, di;
    long long x, dx;
    int ye;
  } edge[2];
  const int shift = kShift;
  const long long delta = (1LL << shift) >> 1;
  int imin = 0;
  int edges = npts;
  long long xmin, xmax, ymin, ymax;
  const long long delta1 = kOne >> 1, delta2 = kOne >> 1;
  P64 p0 = v[npts - 1];
  xmin = xmax = v[0].x;
  ymin = ymax = v[0].y;
  for (int i = 0; i < npts; i++) {
    const P64 p = v[i];
    if (p.y < ymin) {
      ymin = p.y;
      imin = i;
    }
    ymax = ymax > p.y ? ymax : p.y;
    xmax = xmax > p.x ? xmax : p.x;
    xmin = xmin < p.x ? xmin : p.x;
    line2_fixed(p0, p, W, H, emit);
    p0 = p;
  }
  xmin = (xmin + delta) >> shift;
  xmax = (xmax + delta) >> shift;
  ymin = (ymin + delta) >> shift;
  ymax = (ymax + delta) >> shift;
  if (npts < 3 || (int)xmax < 0 || (int)ymax < 0 || (int)xmin >= W || (int)ymin >= H) return;
  ymax = ymax < H - 1 ? ymax : H - 1;
  edge[0].idx = edge[1].idx = imin;
  int y = (int)ymin;
  edge[0].ye = edge[1].ye = y;
  edge[0].di = 1;
  edge[1].di = npts - 1;
  edge[0].x = edge[1].x = -kOne;
  edge[0].dx = edge[1].dx = 0;
  do {
    for (int i = 0; i < 2; i++) {
      if (y >= edge[i].ye) {
        int idx0 = edge[i].idx, di = edge[i].di;
        int idx = idx0 + di;
        if (idx >= npts) idx -= npts;
        int ty = 0;
        for (; edges-- > 0;) {
          ty = (int)((v[idx].y + delta) >> shift);
          if (ty > y) {
            const long long xs = v[idx0].x, xe = v[idx].x;
            edge[i].ye = ty;
            edge[i].dx = ((xe - xs) * 2 + (ty - y)) / (2 * (ty - y));
            edge[i].x = xs;
            edge[i].idx = idx;
            break;
          }
          idx0 = idx;
          idx += di;
          if (idx >= npts) idx -= npts;
        }
      }
    }
    if (edges < 0) break;
    if (y >= 0) {
      int left = 0, right = 1;
      if (edge[0].x > edge[1].x) {
        left = 1;
        right = 0;
      }
      int xx1 = (int)((edge[left].x + delta1) >> kShift);
      int xx2 = (int)((edge[right].x + delta2) >> kShift);
      if (xx2 >= 0 && xx1 < W) {
        if (xx1 < 0) xx1 = 0;
        if (xx2 >= W) xx2 = W - 1;
        emit(y, xx1, xx2);
      }
    }
    edge[0].x += edge[0].dx;
    edge[1].x += edge[1].dx;
  } while (++y <= (int)ymax);
}

// ---------------------------------------------- cv2.line(img, p0, p1, color, thickness, LINE_8)
template <class Emit>
AWX_HD void line(int x0, int y0, int x1, int y1, int thickness, int W, int H, Emit&& emit) {
  if (thickness <= 1) {
    line8(x0, y0, x1, y1, W, H, emit);
    return;
  }
  P64 p0{(long long)x0 << kShift, (long long)y0 << kShift};
  P64 p1{(long long)x1 << kShift, (long long)y1 << kShift};
  const double inv = 1.0 / (double)kOne;
  const double dx = (double)(p0.x - p1.x) * inv, dy = (double)(p1.y - p0.y) * inv;
  double r = dx * dx + dy * dy;
  const int odd = thickness & 1;
  const long long th = (long long)thickness << (kShift - 1);
  if (r > 2.220446049250313e-16) {
#if defined(__CUDA_ARCH__)
    r = ((double)th + odd * (double)kOne * 0.5) / sqrt(r);
#else
    r = ((double)th + odd * (double)kOne * 0.5) / __builtin_sqrt(r);
#endif
    P64 dp{round_half_even(dy * r), round_half_even(dx * r)};
    P64 pt[4];
    pt[0] = P64{p0.x + dp.x, p0.y + dp.y};
    pt[1] = P64{p0.x - dp.x, p0.y - dp.y};
    pt[2] = P64{p1.x - dp.x, p1.y - dp.y};
    pt[3] = P64{p1.x + dp.x, p1.y + dp.y};
    fill_convex_fixed(pt, 4, W, H, emit);
  }
  const int rad = (int)((th + (kOne >> 1)) >> kShift);
  disc(x0, y0, rad, W, H, emit);
  disc(x1, y1, rad, W, H, emit);
}

}  // namespace raster
}  // namespace awx
