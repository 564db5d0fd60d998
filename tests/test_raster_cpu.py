"""The span emitters of csrc/raster.cuh (compiled for the host with g++) against cv2 itself.

OpenCV is the reference's un-vendored dependency for streaks and flakes; the CUDA path restates
its drawing algorithms.  Exhaustive over every (dx, dy, thickness) a rain drop can take and every
placement within 6 px of each border and corner, plus random interior placements; discs of both
radii at every border offset.  Bit-exact.
"""

import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "raster_host.cpp")
OUT_DIR = os.path.join(HERE, "native", "_build")
OUT = os.path.join(OUT_DIR, "libraster_host.so")
CSRC = os.path.join(ROOT, "adverse_weather_semantic_segmentation_robustness_benchmark_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, SRC, "-o", OUT], check=True)
    l = ctypes.CDLL(OUT)
    for fn in (l.raster_lines, l.raster_discs):
        fn.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
        fn.restype = None
    return l


def _ours(fn, items, h, w):
    items = np.ascontiguousarray(items, dtype=np.int32)
    mask = np.zeros((h, w), np.uint8)
    fn(items.ctypes.data, len(items), h, w, mask.ctypes.data)
    return mask


def _cv_lines(items, h, w):
    img = np.zeros((h, w, 3), np.float32)
    for x0, y0, x1, y1, th in items.tolist():
        cv2.line(img, (x0, y0), (x1, y1), [0.8, 0.9, 1.0], th)
    return (img[:, :, 2] > 0).astype(np.uint8)


def _cv_discs(items, h, w):
    img = np.zeros((h, w, 3), np.float32)
    for x, y, r, _, _ in items.tolist():
        cv2.circle(img, (x, y), r, (1.0, 1.0, 1.0), -1)
    return (img[:, :, 0] > 0).astype(np.uint8)


def _drop_shapes():
    # end = start + (int(L sin a), int(L cos a)), L in 5..19, |a| <= 15 deg  ->  dx in -4..4, dy in 4..19;
    # clipping of the end point into the image adds every shorter (dx, dy) including 0
    return [(dx, dy) for dx in range(-5, 6) for dy in range(0, 20)]


def test_lines_exhaustive_single(lib):
    h, w = 40, 36
    bad = 0
    total = 0
    starts = [(x, y) for x in list(range(0, 7)) + [17] + list(range(w - 7, w))
              for y in list(range(0, 7)) + [20] + list(range(h - 7, h))]
    for th in (1, 3):
        for dx, dy in _drop_shapes():
            for x, y in starts:
                ex = int(np.clip(x + dx, 0, w - 1))
                ey = int(np.clip(y + dy, 0, h - 1))
                it = np.array([[x, y, ex, ey, th]], np.int32)
                total += 1
                if not np.array_equal(_ours(lib.raster_lines, it, h, w), _cv_lines(it, h, w)):
                    bad += 1
    assert bad == 0, f"{bad}/{total} single-line footprints differ from cv2.line"


def test_lines_random_scenes(lib):
    rng = np.random.RandomState(5)
    for h, w in ((24, 32), (64, 48), (33, 17)):
        for _ in range(20):
            n = 200
            x = rng.randint(0, w, n); y = rng.randint(0, h, n)
            length = rng.randint(5, 20, n); ang = np.radians(rng.uniform(-15, 15, n))
            ex = np.clip((x + length * np.sin(ang)).astype(int), 0, w - 1)
            ey = np.clip((y + length * np.cos(ang)).astype(int), 0, h - 1)
            th = rng.choice((1, 3), n)
            it = np.stack([x, y, ex, ey, th], 1).astype(np.int32)
            assert np.array_equal(_ours(lib.raster_lines, it, h, w), _cv_lines(it, h, w))


def test_lines_general_directions(lib):
    """Beyond what a rain drop can produce: every octant, both thicknesses."""
    rng = np.random.RandomState(9)
    h, w = 48, 52
    bad = 0
    for _ in range(3000):
        x0, x1 = rng.randint(0, w, 2); y0, y1 = rng.randint(0, h, 2)
        it = np.array([[x0, y0, x1, y1, int(rng.choice((1, 3)))]], np.int32)
        bad += not np.array_equal(_ours(lib.raster_lines, it, h, w), _cv_lines(it, h, w))
    assert bad == 0


def test_discs_exhaustive(lib):
    h, w = 30, 34
    for r in (2, 8, 1, 5):
        for x in list(range(0, 10)) + list(range(w - 10, w)):
            for y in list(range(0, 10)) + [15] + list(range(h - 10, h)):
                it = np.array([[x, y, r, 0, 0]], np.int32)
                assert np.array_equal(_ours(lib.raster_discs, it, h, w), _cv_discs(it, h, w)), (x, y, r)
    # row widths of the two templates
    m = _ours(lib.raster_discs, np.array([[15, 15, 2, 0, 0]], np.int32), h, w)
    assert m.sum(1)[13:18].tolist() == [1, 3, 5, 3, 1]
    m = _ours(lib.raster_discs, np.array([[15, 15, 8, 0, 0]], np.int32), h, w)
    assert m.sum(1)[7:24].tolist() == [1, 7, 11, 13, 13, 15, 15, 15, 17, 15, 15, 15, 13, 13, 11, 7, 1]


def test_unit_of_u8(lib):
    """csrc/convert.cuh: the table-free u8 -> fp32 conversion equals astype(float32) / 255 for every byte."""
    import ctypes as C
    out = np.zeros(256, np.float32)
    lib.unit_table(out.ctypes.data_as(C.c_void_p))
    assert np.array_equal(out, np.arange(256, dtype=np.uint8).astype(np.float32) / np.float32(255.0))
