"""GPU parity of the rows either side of the hot path against the golden
vectors produced by the reference and against the oracle on larger seeded inputs.

Bars: style transfer, depth estimation: exact (uint8 / fp64 integer-derived arithmetic).  Local contrast:
bit-exact where OpenCV runs its vector body (W a multiple of 16), 1 ulp in the last W mod 8 columns.
Fog-density map: exact given the contrast (fp64 blend), i.e. <= 1e-6 absolute overall.  Normalize + CHW:
bit-exact with the restated albumentations algorithm (fp32), bf16 = round-to-nearest-even of it.
Temperature grid: NLL within 2e-6 relative, the selected temperature identical.
"""

import numpy as np
import pytest
import torch

from oracle import prep as op, weather as ow

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pkg():
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    _lib.load()
    return p


@pytest.mark.parametrize("kind", ["fog", "rain", "snow", "night", "clean"])
def test_style_transfer_golden(pkg, golden, kind):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.loader import WeatherAugmentationPipeline
    g = golden("prep")
    pipe = WeatherAugmentationPipeline()
    for src in ("ramp", "rnd"):
        got = pipe._apply_style_transfer(g[f"style_{src}"].copy(), kind)
        assert got.dtype == np.uint8 and np.array_equal(got, g[f"style_{src}_{kind}"]), (kind, src)


def test_style_transfer_unaligned_and_large(pkg):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    rng = np.random.RandomState(1)
    for shape in ((37, 53, 3), (1024, 2048, 3)):
        img = rng.randint(0, 256, shape).astype(np.uint8)
        for kind in ("rain", "night"):
            got = ops_prep.style_transfer(torch.from_numpy(img), kind).cpu().numpy()
            assert np.array_equal(got, op.style_transfer(img.copy(), kind))
    # a view that starts at an odd byte offset (vector path must not be taken blindly)
    flat = torch.from_numpy(rng.randint(0, 256, (3 * 41 + 3,)).astype(np.uint8)).cuda()
    view = flat[3:].view(41, 1, 3)
    got = ops_prep.style_transfer(view, "fog").cpu().numpy()
    assert np.array_equal(got, op.style_transfer(view.cpu().numpy().copy(), "fog"))


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_domain_adaptation_augmentation_golden(pkg, golden, seed):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.loader import WeatherAugmentationPipeline
    g = golden("prep")
    np.random.seed(seed)
    pipe = WeatherAugmentationPipeline(style_transfer_prob=0.6)
    np.random.seed(seed)
    got = pipe.apply_domain_adaptation_augmentation(g["aug_frame"].copy())
    d = np.abs(got.astype(int) - g[f"aug_seed{seed}"].astype(int))
    assert d.max() <= 1 and (d > 0).mean() < 5e-3


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_estimate_depth_golden(pkg, golden, tag):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        DepthEstimationPreprocessor)
    g = golden("prep")
    got = DepthEstimationPreprocessor().estimate_depth(g[f"{tag}_image"])
    want = g[f"{tag}_est_depth"]
    assert got.dtype == np.float64 and got.shape == want.shape
    assert np.array_equal(got, want), f"max diff {np.abs(got - want).max():.3e}"


def test_estimate_depth_batch_vs_oracle(pkg):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        DepthEstimationPreprocessor)
    rng = np.random.RandomState(4)
    imgs = rng.randint(0, 255, (3, 130, 200, 3)).astype(np.uint8)
    imgs[1] = 17  # constant frame: Laplacian identically 0 -> 0 / (0 + 1e-8)
    got = DepthEstimationPreprocessor().estimate_depth_batch(imgs).cpu().numpy()
    for i in range(3):
        assert np.array_equal(got[i], op.estimate_depth(imgs[i])), i


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fog_density_map_golden(pkg, golden, tag):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    g = golden("prep")
    img = g[f"{tag}_image"].astype(np.float32) / 255.0
    t = WeatherDegradationTransforms(seed=13)
    got = t.get_fog_density_map(img, g[f"{tag}_depth"])
    want = g[f"{tag}_fogmap"]
    assert got.dtype == np.float64 and got.shape == want.shape
    w = img.shape[1]
    if w % 16 == 0:
        assert np.array_equal(got, want), f"max diff {np.abs(got - want).max():.3e}"
    else:
        assert np.abs(got - want).max() <= 1e-6
    # depth=None: synthetic depth from the global RNG, filtered on the device (bit-exact with scipy)
    t2 = WeatherDegradationTransforms(seed=14)
    got2 = t2.get_fog_density_map(img)
    assert np.abs(got2 - g[f"{tag}_fogmap_seed14"]).max() <= (0 if w % 16 == 0 else 1e-6)


def test_local_contrast_and_percentile_vs_oracle(pkg):
    """Larger frame, all three input dtypes; the order statistics behind np.percentile are exact."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    rng = np.random.RandomState(8)
    h, w = 192, 256
    img = rng.randint(0, 255, (2, h, w, 3)).astype(np.uint8)
    img[1, :, : w // 2] = 90  # flat half: many identical (zero) contrast values -> ties in the select
    # (x/255)*255 truncates a few bytes down by one, differently in fp32 and fp64: every input dtype is
    # compared with the oracle fed the very same array
    for arr in (img, img.astype(np.float32) / 255.0, img.astype(np.float32).astype(np.float64) / 255.0):
        got = ops_prep.local_contrast(torch.from_numpy(arr)).cpu().numpy()
        want = np.stack([op.local_contrast(arr[i]) for i in range(2)])
        assert np.array_equal(got, want), arr.dtype
    depth = np.stack([ow.depth_from_noise(rng.normal(0, 10, (h, w))) for _ in range(2)])
    got = ops_prep.fog_density_map(torch.from_numpy(img.astype(np.float32) / 255.0), torch.from_numpy(depth)).cpu().numpy()
    for i in range(2):
        ref = op.fog_density_map(img[i].astype(np.float32) / 255.0, depth[i])
        assert np.array_equal(got[i], ref), f"frame {i}: max diff {np.abs(got[i] - ref).max():.3e}"
    # fp32 depth: the blend runs in fp32 as NumPy would
    got32 = ops_prep.fog_density_map(torch.from_numpy(img), torch.from_numpy(depth.astype(np.float32))).cpu().numpy()
    assert got32.dtype == np.float32 and np.abs(got32.astype(np.float64) - got).max() <= 1e-6


@pytest.mark.parametrize("shape", [(2, 64, 96), (1, 37, 53), (2, 1024, 2048)])
def test_normalize_chw(pkg, shape):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.loader import normalize_to_tensor
    b, h, w = shape
    rng = np.random.RandomState(b * h)
    img = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    got = normalize_to_tensor(img)
    assert got.dtype == torch.float32 and tuple(got.shape) == (b, 3, h, w)
    want = np.stack([op.normalize_chw(img[i]) for i in range(b)])
    assert np.array_equal(got.cpu().numpy(), want)
    bf = normalize_to_tensor(img, out_dtype=torch.bfloat16)
    assert bf.dtype == torch.bfloat16
    assert torch.equal(bf.cpu(), torch.from_numpy(want).to(torch.bfloat16))
    single = normalize_to_tensor(img[0])
    assert tuple(single.shape) == (3, h, w) and np.array_equal(single.cpu().numpy(), want[0])


@pytest.mark.parametrize("tag", ["t19", "t5"])
def test_temperature_grid_golden(pkg, golden, tag):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    g = golden("prep")
    logits, targets = torch.from_numpy(g[f"{tag}_logits"]), torch.from_numpy(g[f"{tag}_targets"])
    sums, n_valid, n_bad = ops_prep.temperature_nll(logits, targets, op.temperature_grid())
    assert n_bad == 0 and n_valid == int((g[f"{tag}_targets"] != 255).sum())
    np.testing.assert_allclose(sums / n_valid, g[f"{tag}_nll"], rtol=2e-6, atol=1e-7)
    best = pkg.ConfidenceCalibration().optimize_temperature(logits, targets)
    assert best == float(g[f"{tag}_best_t"])


def test_temperature_grid_large_and_bad_labels(pkg):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    gen = torch.Generator().manual_seed(12)
    logits = torch.randn(2, 19, 96, 160, generator=gen) * 2.5
    targets = torch.randint(0, 19, (2, 96, 160), generator=gen)
    targets[0, :4] = 255
    want = op.temperature_nll(logits, targets).double().numpy()
    for lab in (targets, targets.to(torch.uint8)):
        sums, n_valid, _ = ops_prep.temperature_nll(logits, lab, op.temperature_grid())
        np.testing.assert_allclose(sums / n_valid, want, rtol=2e-6, atol=1e-7)
    assert pkg.ConfidenceCalibration().optimize_temperature(logits, targets) == op.optimize_temperature(logits, targets)
    bad = targets.clone()
    bad[1, 5, 5] = 40
    with pytest.raises(IndexError):
        pkg.ConfidenceCalibration().optimize_temperature(logits, bad)


@pytest.mark.parametrize("h,w", [(64, 128), (50, 70), (1024, 2048)])
def test_corrupt_with_fused_normalize_epilogue(pkg, h, w):
    """awx_corrupt_normalized == awx_normalize_chw(awx_corrupt(...)) bit for bit, for every kind (both blur
    sizes, the clean copy, ragged sizes), fp32 and bf16, with and without the uint8 frame."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    t = WeatherDegradationTransforms(seed=31)
    rng = np.random.RandomState(h + w)
    kinds = ["clean", "fog", "rain", "snow", "night", "snow"] if h < 1000 else ["rain", "snow", "fog"]
    imgs = torch.from_numpy(rng.randint(0, 255, (len(kinds), h, w, 3)).astype(np.uint8)).cuda()
    draws = [t.draw(k, h, w) for k in kinds]
    snow = [d for d in draws if d.kind == "snow"]
    snow[0].blur_k = 3
    snow[-1].blur_k = 7
    for d in draws:
        if d.kind == "fog":
            d.depth = ow.depth_from_noise(d.depth_noise)
    want_u8 = t.corrupt_batch(imgs, draws)
    want = ops_prep.normalize_chw(want_u8)
    norm, u8 = t.corrupt_batch_normalized(imgs, draws, keep_u8=True)
    assert torch.equal(u8, want_u8) and torch.equal(norm, want)
    only = t.corrupt_batch_normalized(imgs, draws)
    assert torch.equal(only, want)
    bf = t.corrupt_batch_normalized(imgs, draws, out_dtype=torch.bfloat16)
    assert bf.dtype == torch.bfloat16 and torch.equal(bf, ops_prep.normalize_chw(want_u8, out_dtype=torch.bfloat16))
