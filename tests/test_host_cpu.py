"""CPU-side checks: the C-ABI library builds, loads and exports every symbol include/awx.h declares
(no compute calls without a GPU), the host RNG replay draws the reference's parameters, the Gaussian
taps equal OpenCV's / SciPy's, and the finalisers reproduce the oracle from integer bins."""

import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(lib):
    handle = lib.load()
    header = open(os.path.join(ROOT, "include", "awx.h")).read()
    declared = set(re.findall(r"\b(awx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no entry points parsed from awx.h"
    assert declared == set(lib.exported_symbols()), declared ^ set(lib.exported_symbols())
    for name in declared:
        assert getattr(handle, name) is not None
    assert handle.awx_version() == 101
    assert handle.awx_launch_count() == 0


def test_bins_layout_and_argument_errors(lib):
    lay = lib.bins_layout(19, 15, 4096)
    assert lay.confusion == 0 and lay.ece_count == 361
    assert lay.total_words == 361 + 4 * 15 + 2 * 4096 + lib.NUM_COUNTERS
    with pytest.raises(RuntimeError, match="num_classes"):
        lib.bins_layout(65, 15, 0)
    with pytest.raises(RuntimeError, match="ece_bins"):
        lib.bins_layout(19, 0, 0)
    # argument validation happens before any CUDA call: usable without a device
    h = lib.load()
    assert h.awx_score(None, None, None, 1, 16, None, None, None, None) == -1
    assert b"cfg is NULL" in h.awx_last_error()
    cfg = lib.ScoreConfig()
    cfg.num_classes, cfg.strategy, cfg.ece_bins = 19, 7, 15
    assert h.awx_score(None, None, None, 1, 16, ctypes.byref(cfg), None, None, None) == -1
    assert b"logits_a is NULL" in h.awx_last_error()
    assert h.awx_score(None, None, None, 0, 16, ctypes.byref(cfg), None, None, None) == 0   # empty batch: no-op
    assert h.awx_confusion(None, 0, None, 0, -1, 19, 255, None, None, None) == -1
    assert h.awx_corrupt_workspace_bytes(2, 8, 40) == 256 + 2 * 8 * 2 * 4
    assert ctypes.sizeof(lib.ScoreConfig) == 4 * 11 + 4 * 65
    assert lib.CORRUPT_PARAMS_DTYPE.itemsize == 64


def test_no_cpu_fallback_without_a_device():
    if torch.cuda.is_available():
        pytest.skip("a device is present")
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.IoUMetrics(19).compute_iou(torch.zeros(1, 19, 4, 4), torch.zeros(1, 4, 4, dtype=torch.long))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.WeatherDegradationTransforms(seed=0).apply_weather_effect(np.zeros((4, 4, 3), np.uint8), "fog")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        p.FogDensityAwareLoss()({"segmentation": torch.zeros(1, 19, 4, 4)}, {"label": torch.zeros(1, 4, 4, dtype=torch.long)})


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "adverse_weather_semantic_segmentation_robustness_benchmark_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in src


@pytest.mark.parametrize("kind", ["fog", "rain", "snow", "night"])
@pytest.mark.parametrize("intensity", [None, 0.5])
def test_host_draws_replay_the_reference_rng(kind, intensity):
    """Same seed -> the product's host draws equal the oracle's (which are pinned to the reference)."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    from oracle import weather as ow
    h, w = 37, 53
    t = WeatherDegradationTransforms(seed=17)
    d = t.draw(kind, h, w, intensity)
    state_after = np.random.get_state()[1].copy()
    np.random.seed(17)
    if kind == "fog":
        o = ow.draw_fog(h, w, intensity)
        assert np.array_equal(d.depth_noise, o["noise"]) and d.intensity == o["intensity"]
    elif kind == "rain":
        o = ow.draw_rain(h, w, intensity)
        assert np.array_equal(d.items, o["drops"]) and d.intensity == o["intensity"] and d.blur_k == 3
    elif kind == "snow":
        o = ow.draw_snow(h, w, intensity)
        assert np.array_equal(d.items[:, :3], o["flakes"]) and d.blur_k == o["blur_k"] and d.intensity == o["intensity"]
    else:
        o = ow.draw_night((h, w, 3), intensity)
        assert np.array_equal(d.noise, o["noise"]) and d.reduction == o["reduction"] and d.intensity == o["intensity"]
    assert np.array_equal(np.random.get_state()[1], state_after), "RNG consumption differs"


def test_gaussian_taps_equal_opencv_and_scipy():
    import cv2
    from scipy.ndimage import gaussian_filter1d
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        gaussian_taps, scipy_gaussian_weights)
    for k, s in ((3, 0.5), (3, 1.0), (7, 1.0)):
        assert np.array_equal(gaussian_taps(k, s), cv2.getGaussianKernel(k, s, cv2.CV_32F).ravel())
    w = scipy_gaussian_weights(2.0)
    assert len(w) == 17
    impulse = np.zeros(33)
    impulse[16] = 1.0
    assert np.array_equal(gaussian_filter1d(impulse, 2.0)[8:25], w)


def test_pack_fills_the_abi_records():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    t = WeatherDegradationTransforms(seed=3)
    h, w = 8, 12
    draws = [t.draw(k, h, w, 0.5) for k in ("fog", "rain", "night", "snow", "clean")]
    draws[0].depth = np.ones((h, w))
    prm, fld, items = t.pack(draws, h, w)
    assert prm["kind"].tolist() == [_lib.FOG, _lib.RAIN, _lib.NIGHT, _lib.SNOW, _lib.CLEAN]
    assert prm["field_offset"][2] == h * w and fld.size == h * w + h * w * 3
    assert prm["d0"][0] == 0.005 + 0.5 * (0.05 - 0.005) and prm["d1"][0] == np.float64(np.float32(0.7 + 0.5 * (1.0 - 0.7)))
    assert prm["f0"][1] == np.float32(1 - 0.15) and prm["f1"][1] == np.float32(0.15 * 0.7)
    assert prm["item_begin"][3] == prm["item_count"][1] and len(items) == prm["item_count"][1] + prm["item_count"][3]
    assert prm["blur_k"][1] == 3 and prm["blur_k"][3] in (3, 7)
    prm2, fld2, _ = t.pack(draws, h, w, gather_fields=False)
    assert fld2 is None and np.array_equal(prm2["field_offset"], prm["field_offset"])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_finalisers_reproduce_the_oracle_from_bins(seed):
    """Build the integer bins on the CPU from the oracle's own intermediate maps, run the product's
    finalisers, compare with the oracle's end results."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import finalize
    from oracle import metrics as om
    gen = torch.Generator().manual_seed(seed)
    c, h, w = 19, 48, 64
    la = torch.randn(2, c, h, w, generator=gen)
    lb = torch.randn(2, c, h, w, generator=gen)
    tgt = torch.randint(0, c, (2, h, w), generator=gen)
    tgt[torch.rand(2, h, w, generator=gen) < 0.05] = 255
    cm = om.confusion_matrix(la, tgt, c).numpy()
    want = om.iou_from_confusion(torch.from_numpy(cm))
    got = finalize.iou_from_confusion(cm)
    assert got["mean_iou"] == want["mean_iou"] and np.array_equal(got["per_class_iou"], want["per_class_iou"])
    ref = om.ece(la, tgt)
    conf, _ = om.confidence_and_prediction(la)
    valid = (tgt != 255)
    idx = om.ece_bin_index(conf[valid].numpy(), om.ece_edges(15).numpy())
    sums = np.array([conf[valid].double().numpy()[idx == b].sum() for b in range(15)])
    out = finalize.ece_from_bins(ref["count"], ref["correct"], sums, int(valid.sum()), om.ece_edges(15).numpy())
    np.testing.assert_allclose(out["ece"], ref["ece"], rtol=1e-5, atol=1e-8)
    for d, r in zip(out["bin_details"], ref["bin_details"]):
        np.testing.assert_allclose(d["accuracy"], r["accuracy"], rtol=1e-6)
        np.testing.assert_allclose(d["confidence"], r["confidence"], rtol=1e-6)
    mi = om.mi_map([la, lb]).numpy()
    wrong = (om.mean_prob_prediction([la, lb]) != tgt).numpy()
    v = valid.numpy()
    nb = 4096
    b = om.mi_bin_index(mi, nb, float(np.float32(np.log(2.0))))
    pos = np.bincount(b[wrong & v], minlength=nb)
    neg = np.bincount(b[~wrong & v], minlength=nb)
    val, bound = finalize.auroc_from_histogram(pos, neg)
    assert abs(val - om.disagreement_auroc([la, lb], tgt)) <= bound + 1e-9
    a2, b2 = om.auroc_from_histogram(pos, neg)
    assert abs(a2 - val) < 1e-12 and abs(b2 - bound) < 1e-12
    assert finalize.auroc_from_histogram(np.zeros(4), np.array([1, 2, 3, 4])) == (0.5, 0.0)
    assert finalize.degradation_ratio(0.5, 0.4) == om.degradation_ratio(0.5, 0.4)
    assert finalize.degradation_ratio(0.0, 0.4) == 1.0


def test_robustness_summary_matches_reference(golden):
    import json
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    g = golden("metrics")
    rob = p.RobustnessMetrics()
    summ = rob.create_robustness_summary({
        "clean": {"mean_iou": 0.5, "expected_calibration_error": 0.02, "ensemble_disagreement_auroc": 0.7},
        "fog": {"mean_iou": 0.3, "expected_calibration_error": 0.05},
        "night": {"mean_iou": 0.45, "expected_calibration_error": 0.03, "ensemble_disagreement_auroc": 0.6},
    })
    keys = json.loads(str(g["summary_keys"]))
    assert sorted(summ) == keys
    assert np.array_equal(np.array([summ[k] for k in keys], dtype=np.float64), g["summary_vals"])
    got = [rob.compute_robustness_degradation_ratio(a, b) for a, b in ((0.5, 0.4), (0.0, 0.3), (0.4, 0.5), (0.78, 0.65))]
    assert np.array_equal(np.array(got), g["degr"])
    assert rob.weather_conditions == ["clean", "fog", "rain", "snow", "night"]


def test_percentile_scalars_match_numpy():
    """The host part of get_fog_density_map's 95th percentile: NumPy's index / gamma / lerp scalars."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    rng = np.random.RandomState(0)
    for n in (3, 5, 17, 1000, 6144, 50 * 70, 1024 * 2048):
        a = rng.rand(n).astype(np.float32)
        lo, hi, gamma = ops_prep.percentile_indices(n, 95, np.float32)
        srt = np.sort(a)
        got = ops_prep.lerp(srt[lo], srt[hi], gamma)
        assert got.dtype == np.float32 and got == np.percentile(a, 95), n


def test_normalize_params_are_fp32_albumentations_scalars():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
    m, r = ops_prep.normalize_params(ops_prep.IMAGENET_MEAN, ops_prep.IMAGENET_STD)
    assert m.dtype == np.float32 and r.dtype == np.float32
    assert np.array_equal(m, np.array(ops_prep.IMAGENET_MEAN, np.float32) * np.float32(255.0))
    assert np.array_equal(r, np.float32(1) / (np.array(ops_prep.IMAGENET_STD, np.float32) * np.float32(255.0)))


def test_argument_errors_of_the_prep_entry_points(lib):
    """Every 8f-row entry point validates its arguments before touching CUDA: empty inputs are no-ops,
    NULL pointers / bad dtypes / bad sizes return AWX_E_ARG or AWX_E_UNSUPPORTED with a message."""
    h = lib.load()
    E_ARG, E_UNSUP = -1, -2
    f3 = (ctypes.c_float * 3)(1, 2, 3)
    assert h.awx_normalize_chw(None, None, lib.F32, 0, 8, 8, f3, f3, None) == 0
    assert h.awx_normalize_chw(None, None, lib.F32, 1, 8, 8, f3, f3, None) == E_ARG
    assert b"NULL pointer" in h.awx_last_error()
    assert h.awx_normalize_chw(None, None, lib.F64, 1, 8, 8, f3, f3, None) == E_ARG
    assert h.awx_style_transfer(None, None, 0, 1.0, 0.0, 0.0, 0, None) == 0
    assert h.awx_style_transfer(None, None, 4, 1.0, 0.0, 0.0, 0, None) == E_ARG
    assert h.awx_temperature_workspace_bytes(100) > 0 and h.awx_temperature_workspace_bytes(0) == 0
    assert h.awx_temperature_workspace_bytes(129) == 0
    t = (ctypes.c_float * 2)(1.0, -1.0)
    assert h.awx_temperature_nll(None, None, lib.LABEL_I64, 0, 19, 255, t, 2, None, None, None) == 0
    assert h.awx_temperature_nll(None, None, lib.LABEL_I64, 4, 19, 255, t, 2, None, None, None) == E_ARG
    assert h.awx_temperature_nll(None, None, lib.LABEL_I64, 4, 65, 255, t, 2, None, None, None) == E_UNSUP
    assert h.awx_temperature_nll(None, None, lib.LABEL_I64, 4, 19, 255, t, 500, None, None, None) == E_UNSUP
    assert h.awx_fog_density_workspace_bytes(0) == 0 and h.awx_fog_density_workspace_bytes(2) >= 2 * (4096 * 3 * 4)
    assert h.awx_local_contrast(None, lib.U8, None, 0, 8, 8, 0, 0, None, None, None) == 0
    assert h.awx_local_contrast(None, lib.U8, None, 1, 8, 8, 0, 0, None, None, None) == E_ARG
    assert h.awx_fog_density_finish(None, None, lib.F64, None, None, 0, 64, None, None) == 0
    assert h.awx_fog_density_finish(None, None, lib.F64, None, None, 1, 64, None, None) == E_ARG
    assert h.awx_estimate_depth(None, None, None, 0, 8, 8, None, 8, None, None) == 0
    assert h.awx_estimate_depth(None, None, None, 1, 8, 8, None, 8, None, None) == E_ARG
    assert h.awx_corrupt_normalized(None, None, None, lib.F32, None, None, 1, 8, 8, None, None, lib.F64, None, 0, None, None) == E_ARG
    assert h.awx_corrupt_normalized(None, None, None, lib.F32, None, None, 0, 8, 8, None, None, lib.F64, None, 0, None, None) == 0


def test_evaluate_model_signature_matches_the_reference():
    import inspect
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import evaluate_model
    params = list(inspect.signature(evaluate_model).parameters)
    assert params[:5] == ["model", "test_loader", "metrics", "device", "config"]   # scripts/evaluate.py:134-140
    # the evaluator (device bins of the common size) exists before the loop, so that a rank with an empty shard still
    # takes part in the all_reduce: without a CUDA device the call fails loudly even for an empty loader
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            evaluate_model(object(), [], None, None, None)
