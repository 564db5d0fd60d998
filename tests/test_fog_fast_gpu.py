"""GPU parity of the screened fog kernel (fog_kernel, csrc/corrupt.cu + csrc/fog_fast.cuh): an fp32 screen in front
of the reference's fp64 expression.  Its bytes must be those of the generic kernel, which evaluates the fp64
expression for every pixel (AWX_FOG_KERNEL=exact forces it), and those of the oracle -- 0 LSB, as before."""

import numpy as np
import pytest
import torch

from oracle import weather as ow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def wdt():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    _lib.load()
    return WeatherDegradationTransforms


@pytest.mark.parametrize("h,w,b", [(1024, 2048, 3), (96, 160, 4), (37, 53, 5), (1, 7, 2)])
def test_screened_and_exact_kernels_agree_and_match_the_oracle(wdt, monkeypatch, h, w, b):
    t = wdt(seed=h + b)
    rng = np.random.RandomState(w)
    imgs = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    kinds = ["fog", "fog", "night", "fog", "clean"][:b]
    intens = [0.3, 0.9, 0.6, 1.0, None][:b]
    draws = [t.draw(k, h, w, i) for k, i in zip(kinds, intens)]
    for d in draws:
        if d.kind == "fog":
            d.depth = ow.depth_from_noise(d.depth_noise)
    monkeypatch.delenv("AWX_FOG_KERNEL", raising=False)
    a = t.corrupt_batch(imgs, draws)
    monkeypatch.setenv("AWX_FOG_KERNEL", "exact")
    c = t.corrupt_batch(imgs, draws)
    assert torch.equal(a, c), f"{int((a != c).sum())} values differ between the screened and the exact fog kernel"
    a = a.cpu().numpy()
    for i, d in enumerate(draws):
        if d.kind == "fog":
            want = ow.fog_apply(imgs[i], d.depth, d.intensity)
            assert np.array_equal(a[i], want), f"image {i}: {int((a[i] != want).sum())} values differ from the oracle"


def test_out_of_range_coefficients_fall_back_to_the_exact_kernel(wdt):
    """intensity 1.5 -> airlight 1.15 > 1: the screen's range argument does not hold, the generic kernel runs."""
    t = wdt(seed=1)
    h, w = 64, 96
    rng = np.random.RandomState(0)
    imgs = rng.randint(0, 256, (1, h, w, 3)).astype(np.uint8)
    d = t.draw("fog", h, w, 1.5)
    d.depth = ow.depth_from_noise(d.depth_noise)
    got = t.corrupt_batch(imgs, [d]).cpu().numpy()[0]
    assert np.array_equal(got, ow.fog_apply(imgs[0], d.depth, d.intensity))


@pytest.mark.parametrize("h,w,b", [(1024, 2048, 2), (96, 160, 3), (37, 54, 4), (5, 6, 2)])
def test_flat_and_staged_night_kernels_agree_and_match_the_oracle(wdt, monkeypatch, h, w, b):
    """night_kernel (flat element-wise, no shared memory) against the staged generic kernel (AWX_NIGHT_KERNEL=staged)
    and the oracle: 0 LSB.  Batches mix kinds, so both point kernels and the fog kernel run side by side."""
    t = wdt(seed=w + b)
    rng = np.random.RandomState(h)
    imgs = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    kinds = ["night", "fog", "night", "clean"][:b]
    draws = [t.draw(k, h, w) for k in kinds]
    for d in draws:
        if d.kind == "fog":
            d.depth = ow.depth_from_noise(d.depth_noise)
    monkeypatch.delenv("AWX_NIGHT_KERNEL", raising=False)
    a = t.corrupt_batch(imgs, draws)
    monkeypatch.setenv("AWX_NIGHT_KERNEL", "staged")
    c = t.corrupt_batch(imgs, draws)
    assert torch.equal(a, c), f"{int((a != c).sum())} values differ between the flat and the staged night kernel"
    a = a.cpu().numpy()
    for i, d in enumerate(draws):
        if d.kind == "night":
            want = ow.night_apply(imgs[i], d.intensity, d.reduction, d.noise)
            assert np.array_equal(a[i], want), f"image {i}: {int((a[i] != want).sum())} values differ from the oracle"
