"""GPU parity of the row-walking rain / snow kernel (blur_strip_kernel, csrc/corrupt.cu + csrc/blur_strip.cuh).

It restates cv2.GaussianBlur's operation order (which products are fused), so on frames it accepts -- whole 16-pixel
units per row -- the corrupted image is expected to be IDENTICAL to the oracle's (cv2 itself), not merely within
1 LSB; the tile kernel (any width; AWX_BLUR_KERNEL=tile forces it) must produce the same bytes.  The per-thread code
is also checked on the host, without a GPU, by tests/test_blur_strip_cpu.py.
"""

import numpy as np
import pytest
import torch

from oracle import weather as ow

pytestmark = pytest.mark.gpu

from parity import cv2_blur_follows_the_restated_order

# cv2 picks its filter kernels by CPU dispatch and fuses multiply-adds only in FMA builds: the 0-LSB bar holds where a
# self-check of this machine's cv2 against the restated order passes (every box seen so far), the 1-LSB bar elsewhere
EXACT = cv2_blur_follows_the_restated_order()


@pytest.fixture(scope="module")
def wdt():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    _lib.load()
    return WeatherDegradationTransforms


def _want(img, d):
    if d.kind == "rain":
        return ow.rain_apply(img, d.intensity, d.items)
    return ow.snow_apply(img, d.intensity, d.items[:, :3], d.blur_k)


def _check(got, want, what):
    diff = np.abs(got.astype(np.int16) - want.astype(np.int16))
    if EXACT:
        assert diff.max() == 0, f"{what}: {int((diff > 0).sum())} of {diff.size} values differ (max {int(diff.max())})"
    else:
        assert diff.max() <= 1 and (diff > 0).mean() < 2e-3, what


@pytest.mark.parametrize("h,w", [(16, 16), (1, 32), (33, 48), (96, 160), (64, 528), (150, 2048), (131, 1040)])
def test_strip_kernel_identical_to_cv2(wdt, h, w):
    t = wdt(seed=h * 7 + w)
    rng = np.random.RandomState(h + w)
    imgs = rng.randint(0, 256, (4, h, w, 3)).astype(np.uint8)
    draws = [t.draw("rain", h, w, 0.8), t.draw("snow", h, w, 0.7), t.draw("snow", h, w, 0.3), t.draw("rain", h, w)]
    draws[1].blur_k, draws[2].blur_k = 3, 7
    out = t.corrupt_batch(imgs, draws).cpu().numpy()
    for i, d in enumerate(draws):
        _check(out[i], _want(imgs[i], d), f"{d.kind} k={getattr(d, 'blur_k', 3)} {h}x{w}")


def test_full_size_frames_identical_to_cv2(wdt):
    t = wdt(seed=77)
    h, w = 1024, 2048
    rng = np.random.RandomState(5)
    imgs = rng.randint(0, 256, (3, h, w, 3)).astype(np.uint8)
    draws = [t.draw("snow", h, w, 0.7), t.draw("rain", h, w, 0.8), t.draw("snow", h, w, 0.5)]
    draws[0].blur_k, draws[2].blur_k = 7, 3
    out = t.corrupt_batch(imgs, draws).cpu().numpy()
    for i, d in enumerate(draws):
        _check(out[i], _want(imgs[i], d), f"{d.kind} full size")


@pytest.mark.parametrize("h,w,b", [(1024, 2048, 4), (200, 512, 6), (37, 64, 5)])
def test_strip_and_tile_kernels_agree(wdt, monkeypatch, h, w, b):
    """The two kernels implement the same arithmetic: identical bytes, mixed batch (kinds and blur sizes)."""
    t = wdt(seed=3)
    rng = np.random.RandomState(b)
    imgs = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    kinds = ["rain", "snow", "snow", "clean", "rain", "snow"][:b]
    draws = [t.draw(k, h, w) for k in kinds]
    draws[1].blur_k, draws[2].blur_k = 7, 3
    monkeypatch.delenv("AWX_BLUR_KERNEL", raising=False)
    a = t.corrupt_batch(imgs, draws)
    monkeypatch.setenv("AWX_BLUR_KERNEL", "tile")
    c = t.corrupt_batch(imgs, draws)
    assert torch.equal(a, c), f"{int((a != c).sum())} values differ between the strip and the tile kernel"
