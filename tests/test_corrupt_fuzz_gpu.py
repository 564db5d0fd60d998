"""Seeded fuzz of awx_corrupt / awx_corrupt_normalized against the oracle: random frame sizes (down to a few
pixels, odd widths, tiles that straddle the border), kinds, intensities and blur sizes.  Bars as in
test_corrupt_gpu.py: fog / night / clean exact with fp64 fields, rain / snow within 1 LSB; the fused
Normalize + CHW epilogue bit-identical to normalising the uint8 result."""

import numpy as np
import pytest
import torch

from oracle import weather as ow, prep as op

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", range(24))
def test_corrupt_fuzz(seed):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    rng = np.random.RandomState(1000 + seed)
    h = int(rng.choice([4, 9, 16, 17, 31, 48, 65, 130]))
    w = int(rng.choice([5, 12, 16, 33, 64, 100, 129, 260]))
    kinds = list(rng.choice(["clean", "fog", "rain", "snow", "night"], size=int(rng.randint(1, 5))))
    t = WeatherDegradationTransforms(seed=seed)
    imgs = rng.randint(0, 256, (len(kinds), h, w, 3)).astype(np.uint8)
    draws = []
    for k in kinds:
        inten = None if rng.rand() < 0.3 else float(rng.uniform(0.05, 1.0))
        d = t.draw(k, h, w, inten)
        if k == "snow":
            d.blur_k = int(rng.choice([3, 7]))
        if k == "fog":
            d.depth = ow.depth_from_noise(d.depth_noise)
        draws.append(d)
    out = t.corrupt_batch(imgs, draws)
    got = out.cpu().numpy()
    for i, d in enumerate(draws):
        if d.kind == "clean":
            want = imgs[i]
        elif d.kind == "fog":
            want = ow.fog_apply(imgs[i], d.depth, d.intensity)
        elif d.kind == "rain":
            want = ow.rain_apply(imgs[i], d.intensity, d.items)
        elif d.kind == "snow":
            want = ow.snow_apply(imgs[i], d.intensity, d.items[:, :3], d.blur_k)
        else:
            want = ow.night_apply(imgs[i], d.intensity, d.reduction, d.noise)
        diff = np.abs(got[i].astype(np.int16) - want.astype(np.int16))
        assert diff.max() <= (0 if d.kind in ("clean", "fog", "night") else 1), (d.kind, h, w, int(diff.max()))
        if d.kind in ("rain", "snow"):
            assert (diff > 0).mean() < 1e-2, (d.kind, h, w, float((diff > 0).mean()))
    norm, u8 = t.corrupt_batch_normalized(imgs, draws, keep_u8=True)
    assert torch.equal(u8, out)
    assert torch.equal(norm, ops_prep.normalize_chw(out))
    assert np.array_equal(norm[0].cpu().numpy(), op.normalize_chw(got[0]))
