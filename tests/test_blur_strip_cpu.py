"""csrc/blur_strip.cuh (the per-thread code of the row-walking rain / snow kernel) compiled for the host with g++
and run under a sequential CTA emulation, against the oracle (cv2.line / cv2.circle / cv2.GaussianBlur themselves).

The device kernel (blur_strip_kernel, csrc/corrupt.cu) and the emulation (tests/native/blur_strip_host.cpp) share
every function of the header; only the loop over threads, the barrier and the 16-byte loads differ.  What this pins
without a GPU: the unit / pair / halo index algebra, BORDER_REFLECT_101 on both axes and at strip and segment seams,
the overlay bits of halo pixels, the static register-window rotation of the vertical pass, and OpenCV's operation
order (which products are fused) -- bit for bit, on every width class the kernel accepts.
"""

import ctypes
import os
import subprocess

import cv2
import numpy as np
import pytest

from oracle import weather as ow
from parity import cv2_blur_follows_the_restated_order

# cv2 chooses its filter kernels by CPU dispatch; the emulation is compared bit for bit where this machine's cv2
# follows the order the header restates (a self-check on a random image; true on every machine seen so far)
pytestmark = pytest.mark.skipif(not cv2_blur_follows_the_restated_order(),
                                reason="this machine's cv2.GaussianBlur uses another operation order")

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "blur_strip_host.cpp")
OUT_DIR = os.path.join(HERE, "native", "_build")
OUT = os.path.join(OUT_DIR, "libblur_strip_host.so")
CSRC = os.path.join(ROOT, "adverse_weather_semantic_segmentation_robustness_benchmark_b200", "csrc")


@pytest.fixture(scope="module")
def lib():
    os.makedirs(OUT_DIR, exist_ok=True)
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-I", CSRC, SRC, "-o", OUT], check=True)
    l = ctypes.CDLL(OUT)
    l.blur_strip_emulate.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                     ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.c_void_p, ctypes.c_int]
    l.blur_strip_emulate.restype = ctypes.c_int
    return l


def _taps(ksize, sigma):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import gaussian_taps
    t = gaussian_taps(ksize, sigma)
    half = np.zeros(4, np.float32)
    half[: ksize // 2 + 1] = t[ksize // 2:]
    return half


def _mask_words(mask_hw):
    """1 bit per pixel, bit j of word w = pixel 32 w + j (the layout rasterize_kernel writes)."""
    h, w = mask_hw.shape
    ww = (w + 31) // 32
    padded = np.zeros((h, ww * 32), np.uint8)
    padded[:, :w] = mask_hw
    return np.packbits(padded.reshape(h, ww, 32), axis=2, bitorder="little").view(np.uint32).reshape(h, ww).copy()


def _emulate(lib, img, mask_hw, radius, rain, k1, k2, taps, seg):
    h, w = img.shape[:2]
    img = np.ascontiguousarray(img)
    out = np.zeros_like(img)
    words = _mask_words(mask_hw)
    taps = np.ascontiguousarray(taps, np.float32)
    rc = lib.blur_strip_emulate(img.ctypes.data, out.ctypes.data, words.ctypes.data, h, w, radius, int(rain),
                                ctypes.c_float(k1), ctypes.c_float(k2), taps.ctypes.data, seg)
    assert rc == 0
    return out


def _rain_case(lib, h, w, seed, seg, dense=False):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    np.random.seed(seed)
    d = ow.draw_rain(h, w, 0.9 if dense else None)
    drops = d["drops"]
    if dense:  # streaks on and across every border
        extra = [(x, y, min(w - 1, x + 2), min(h - 1, y + 9), 3) for x in (0, 1, w - 2, w - 1, w // 2) for y in (0, h - 3, h // 2)]
        drops = np.concatenate([drops, np.array(extra, np.int32)])
    ref = ow.rain_apply(img, d["intensity"], drops)
    canvas = np.zeros((h, w, 3), np.float32)
    for x0, y0, x1, y1, th in drops.tolist():
        cv2.line(canvas, (x0, y0), (x1, y1), [1.0, 1.0, 1.0], th)
    haze = d["intensity"] * 0.3
    got = _emulate(lib, img, (canvas[:, :, 0] > 0).astype(np.uint8), 1, True, np.float32(1 - haze), np.float32(haze * 0.7),
                   _taps(3, 0.5), seg)
    return got, ref


def _snow_case(lib, h, w, seed, seg, blur_k):
    rng = np.random.RandomState(seed)
    img = rng.randint(0, 256, (h, w, 3)).astype(np.uint8)
    np.random.seed(seed)
    d = ow.draw_snow(h, w)
    flakes = np.concatenate([d["flakes"], np.array([(0, 0, 8), (w - 1, h - 1, 8), (w - 1, 0, 2), (0, h - 1, 2), (w // 2, 0, 8)], np.int32)])
    ref = ow.snow_apply(img, d["intensity"], flakes, blur_k)
    canvas = np.zeros((h, w, 3), np.float32)
    for x, y, r in flakes.tolist():
        cv2.circle(canvas, (x, y), r, (1.0, 1.0, 1.0), -1)
    got = _emulate(lib, img, (canvas[:, :, 0] > 0).astype(np.uint8), blur_k // 2, False, np.float32(d["intensity"] * 0.2), 0.0,
                   _taps(blur_k, 1.0), seg)
    return got, ref


SHAPES = [(1, 16), (2, 16), (3, 32), (7, 48), (16, 16), (23, 512), (40, 528), (37, 1040), (64, 2048), (9, 496)]


@pytest.mark.parametrize("h,w", SHAPES)
def test_rain_matches_cv2_bit_for_bit(lib, h, w):
    for seed, seg in ((1, 128), (2, 5), (3, 1)):
        got, ref = _rain_case(lib, h, w, seed, seg, dense=seed == 2)
        assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {got.size} values differ (seg {seg})"


@pytest.mark.parametrize("h,w", SHAPES)
@pytest.mark.parametrize("blur_k", [3, 7])
def test_snow_matches_cv2_bit_for_bit(lib, h, w, blur_k):
    for seed, seg in ((4, 128), (5, 11), (6, 2)):
        got, ref = _snow_case(lib, h, w, seed, seg, blur_k)
        assert np.array_equal(got, ref), f"{int((got != ref).sum())} of {got.size} values differ (seg {seg})"


def test_full_width_frame_rows(lib):
    """A Cityscapes-wide band (4 strips of 32 units), segments of 128 rows as the launcher picks them."""
    got, ref = _rain_case(lib, 150, 2048, 11, 128)
    assert np.array_equal(got, ref)
    got, ref = _snow_case(lib, 150, 2048, 12, 128, 7)
    assert np.array_equal(got, ref)
