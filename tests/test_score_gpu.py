"""GPU parity: awx_score / awx_confusion / metric classes against the oracle and the golden vectors.

Bars: confusion matrices, argmax maps, fused logits: ``==``; ECE bin counts
``==`` except for the pixels the kernel itself reports as ambiguous (confidence within 3 ulp of a
bin edge, i.e. inside the reference's own rounding noise), and ``==`` unconditionally with respect
to the kernel's own emitted confidence map; float maps: |d| <= 1e-5*|ref| + 2e-6; AUROC: within the
stated histogram bound (+1e-6) of sklearn's exact value.
"""

import numpy as np
import pytest
import torch

from oracle import metrics as om, fusion as of_
import parity

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-5, 2e-6


@pytest.fixture(scope="module")
def pkg():
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib
    _lib.load()
    return p, ops, _lib


def _close(got, want, rtol=RTOL, atol=ATOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    err = np.abs(got - want) - (atol + rtol * np.abs(want))
    assert err.max() <= 0, f"max excess {err.max():.3e}, max abs diff {np.abs(got - want).max():.3e}"


CASES = [("c5", 5), ("c19_i64_ign", 19), ("c19_u8", 19)]


@pytest.mark.parametrize("tag,c", CASES)
def test_golden_metric_classes(pkg, golden, tag, c):
    p, ops, _lib = pkg
    g = golden("metrics")
    la, lb = torch.from_numpy(g[f"{tag}_la"]), torch.from_numpy(g[f"{tag}_lb"])
    tgt = torch.from_numpy(g[f"{tag}_target"])
    iou = p.IoUMetrics(c)
    r = iou.compute_iou(la, tgt)
    # integer confusion -> same fp32 finalisation -> identical numbers
    assert r["mean_iou"] == float(g[f"{tag}_mean_iou"])
    assert np.array_equal(r["per_class_iou"], g[f"{tag}_per_class_iou"])
    assert np.array_equal(r["valid_classes"], g[f"{tag}_valid_classes"])
    assert iou.compute_pixel_accuracy(la, tgt) == float(g[f"{tag}_pixel_accuracy"])
    # prediction-map entry (evaluate.py / trainer call compute_miou with argmax'd maps)
    r2 = iou.compute_iou(la.argmax(1), tgt)
    assert r2["mean_iou"] == r["mean_iou"]

    cal = p.ConfidenceCalibration()
    d = cal.compute_ece(la, tgt, return_details=True)
    assert d["ambiguous_pixels"] == 0
    _close(d["ece"], g[f"{tag}_ece"], rtol=1e-5, atol=1e-7)
    assert cal.compute_ece(la, tgt) == d["ece"]
    for key in ("accuracy", "confidence", "proportion", "error"):
        _close([x[key] for x in d["bin_details"]], g[f"{tag}_ece_{key}"], rtol=1e-5, atol=1e-7)
    for key in ("bin_lower", "bin_upper"):
        assert np.array_equal(np.array([x[key] for x in d["bin_details"]]), g[f"{tag}_ece_{key}"])
    _close(d["overall_accuracy"], g[f"{tag}_ece_overall_accuracy"], rtol=1e-6, atol=0)
    _close(d["overall_confidence"], g[f"{tag}_ece_overall_confidence"], rtol=1e-6, atol=0)
    _close(cal.compute_reliability_diagram_data(la, tgt)["bin_centers"], g[f"{tag}_rel_centers"], rtol=1e-6, atol=0)
    assert torch.equal(cal.temperature_scale(la, 1.7).cpu(), torch.from_numpy(g[f"{tag}_temp_scaled"]))

    ens = p.EnsembleDisagreementMetrics()
    _close(ens.compute_disagreement_map([la, lb]).cpu().numpy(), g[f"{tag}_mi"])
    _close(ens.compute_jensen_shannon_divergence(la, lb).cpu().numpy(), g[f"{tag}_js"])
    _close(ens.compute_variance_map([la, lb]).cpu().numpy(), g[f"{tag}_var"], rtol=1e-5, atol=1e-7)
    val, bound = ens.compute_disagreement_auroc([la, lb], tgt, return_bound=True)
    assert abs(val - float(g[f"{tag}_auroc"])) <= bound + 1e-6
    assert bound < 2e-3

    rob = p.RobustnessMetrics(num_classes=c)
    comp = rob.compute_comprehensive_metrics(la, tgt, [la, lb], "fog")
    import json
    keys = json.loads(str(g[f"{tag}_comp_keys"]))
    assert sorted(comp) == keys
    want = dict(zip(keys, g[f"{tag}_comp_vals"]))
    assert comp["mean_iou"] == want["mean_iou"] and comp["miou_fog"] == want["miou_fog"]
    assert comp["pixel_accuracy"] == want["pixel_accuracy"]
    _close(comp["expected_calibration_error"], want["expected_calibration_error"], rtol=1e-5, atol=1e-7)
    assert abs(comp["ensemble_disagreement_auroc"] - want["ensemble_disagreement_auroc"]) <= bound + 1e-6


def _rand_case(seed, b, c, h, w, label_dtype=torch.int64, ignore=0.0, scale=1.0):
    gen = torch.Generator().manual_seed(seed)
    la = torch.randn(b, c, h, w, generator=gen) * scale
    lb = torch.randn(b, c, h, w, generator=gen) * scale + 0.4 * la
    tgt = torch.randint(0, c, (b, h, w), generator=gen).to(label_dtype)
    if ignore > 0:
        tgt[torch.rand(b, h, w, generator=gen) < ignore] = 255
    return la, lb, tgt


@pytest.mark.parametrize("c,h,w,ldt", [(19, 96, 128, torch.int64), (19, 64, 66, torch.uint8),
                                        (19, 33, 35, torch.int64), (7, 40, 50, torch.uint8), (1, 8, 8, torch.int64)])
def test_single_model_bins_vs_oracle(pkg, c, h, w, ldt):
    """Confusion + accuracy exact; ECE bins exact given the emitted confidence and equal to the
    oracle's up to the ambiguous count (odd sizes exercise the scalar path, C!=19 the generic one)."""
    p, ops, _lib = pkg
    la, _, tgt = _rand_case(100 + c + h, 2, c, h, w, ldt, ignore=0.03 if c > 1 else 0.0)
    out = ops.score(la, None, tgt, want_pred=torch.int64, want_conf=True)
    bins = ops.read_bins(out["bins"], c, 15, 0)
    assert torch.equal(out["pred"].cpu(), la.argmax(1))
    assert np.array_equal(bins.confusion, om.confusion_matrix(la, tgt, c).numpy())
    valid = tgt != 255
    assert bins.counter(_lib.CNT_VALID) == int(valid.sum())
    assert bins.counter(_lib.CNT_PIXELS) == tgt.numel()
    assert bins.counter(_lib.CNT_CORRECT) == int(((la.argmax(1) == tgt) & valid).sum())
    assert bins.counter(_lib.CNT_BAD_LABEL) == 0
    ref = om.ece(la, tgt)
    conf = out["conf"].cpu()
    conf_ref, _ = om.confidence_and_prediction(la)
    # confidence within 2 ulp of torch's
    assert (conf - conf_ref).abs().max() <= 2.4e-7
    # (ii) counts are exact w.r.t. the kernel's own confidence map
    idx = om.ece_bin_index(conf[valid].numpy(), om.ece_edges(15).numpy())
    own = np.bincount(idx[idx >= 0], minlength=15)
    assert np.array_equal(bins.ece_count, own)
    # (iii) and equal to the oracle's up to the reported ambiguous pixels, of which there may be at most one here
    parity.assert_ece_parity(bins, ref, int(valid.sum()), _lib, floor=1)
    ref_sum = np.array([float(conf_ref[valid][torch.from_numpy(idx == b)].double().sum()) for b in range(15)])
    _close(bins.ece_conf_sum, ref_sum, rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean"])
@pytest.mark.parametrize("temp", [None, 1.0, 1.7, 0.5])
def test_ensemble_fusion_exact(pkg, strategy, temp):
    p, ops, _lib = pkg
    la, lb, tgt = _rand_case(7, 2, 19, 48, 64)
    raw_w = torch.tensor([0.3, 0.9])
    w = of_.member_weights(raw_w)
    code = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF, "mean": _lib.FUSE_MEAN}[strategy]
    tt = None if temp is None else torch.tensor([temp])
    want = of_.fuse_logits(la, lb, strategy, raw_w, tt)
    out = ops.score(la, lb, tgt, strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=temp,
                    want_pred=torch.int64, want_fused=True, auroc_bins=1024)
    assert torch.equal(out["fused"].cpu(), want), "fused logits must be bit-exact"
    assert torch.equal(out["pred"].cpu(), want.argmax(1))
    bins = ops.read_bins(out["bins"], 19, 15, 1024)
    assert np.array_equal(bins.confusion, om.confusion_matrix(want, tgt, 19).numpy())
    # same statistics without materialising the fused logits (the fast path divides only near ties)
    out2 = ops.score(la, lb, tgt, strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=temp, auroc_bins=1024)
    b2 = ops.read_bins(out2["bins"], 19, 15, 1024)
    assert np.array_equal(b2.confusion, bins.confusion)
    ref = om.ece(want, tgt)
    n_valid = int((tgt != 255).sum())
    if bins.counter(_lib.CNT_PICK_AMBIG) == 0:
        parity.assert_ece_parity(bins, ref, n_valid, _lib, floor=1)
        parity.assert_ece_parity(b2, ref, n_valid, _lib, floor=1)
    wrong = int(((om.mean_prob_prediction([la, lb]) != tgt) & (tgt != 255)).sum())
    parity.assert_ens_wrong_parity(bins, wrong, n_valid, _lib, floor=1)
    parity.assert_ens_wrong_parity(b2, wrong, n_valid, _lib, floor=1)


def test_ensemble_maps_and_auroc(pkg):
    p, ops, _lib = pkg
    la, lb, tgt = _rand_case(11, 2, 19, 64, 96, ignore=0.02, scale=2.0)
    nbins = 4096
    out = ops.score(la, lb, tgt, strategy=_lib.FUSE_MEAN, auroc_bins=nbins, want_mi=True, want_js=True)
    mi = out["mi"].cpu()
    _close(mi.numpy(), om.mi_map([la, lb]).numpy())
    _close(out["js"].cpu().numpy(), of_.reverse_kl_disagreement(la, lb).numpy())
    bins = ops.read_bins(out["bins"], 19, 15, nbins)
    valid = (tgt != 255)
    wrong = (om.mean_prob_prediction([la, lb]) != tgt)
    parity.assert_ens_wrong_parity(bins, int((wrong & valid).sum()), int(valid.sum()), _lib, floor=1)
    assert bins.counter(_lib.CNT_MARG_AMBIG) == 0   # this seed has no tie of the mean probabilities
    # histogram counts exact w.r.t. the emitted MI map
    idx = om.mi_bin_index(mi.numpy(), nbins, float(np.float32(np.log(2.0))))
    pos = np.bincount(idx[(wrong & valid).numpy()], minlength=nbins)
    neg = np.bincount(idx[(~wrong & valid).numpy()], minlength=nbins)
    assert np.array_equal(bins.auroc_pos, pos) and np.array_equal(bins.auroc_neg, neg)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import finalize
    val, bound = finalize.auroc_from_histogram(bins.auroc_pos, bins.auroc_neg)
    assert abs(val - om.disagreement_auroc([la, lb], tgt)) <= bound + 1e-6


def test_ties_nan_and_degenerate(pkg):
    p, ops, _lib = pkg
    la = torch.zeros(1, 19, 4, 8)                      # all ties -> class 0, confidence 1/19
    la[0, 5, 0, 0] = float("nan")                      # NaN wins the argmax
    la[0, 3, 1, :] = 50.0                              # saturated confidence (1.0, top bin)
    la[0, 7, 2, 0] = la[0, 9, 2, 0] = 3.0              # tie between 7 and 9 -> 7
    tgt = torch.zeros(1, 4, 8, dtype=torch.int64)
    out = ops.score(la, None, tgt, want_pred=torch.int64, want_conf=True)
    assert torch.equal(out["pred"].cpu(), la.argmax(1))
    bins = ops.read_bins(out["bins"], 19, 15, 0)
    ref = om.ece(la, tgt)
    assert np.array_equal(bins.ece_count, ref["count"])
    assert bins.counter(_lib.CNT_NO_BIN) == 1            # the NaN pixel
    # AUROC with a single class present is 0.5 (metrics.py:430-431)
    ens = p.EnsembleDisagreementMetrics()
    lb = torch.zeros_like(la)
    assert ens.compute_disagreement_auroc([torch.zeros(1, 19, 4, 8), lb], tgt) == 0.5
    with pytest.raises(ValueError, match="Need at least 2 predictions"):
        ens.compute_disagreement_map([la])
    # labels outside [0,C) that are not ignore_index: the reference's index_add_ raises
    bad = torch.full((1, 4, 8), 40, dtype=torch.int64)
    with pytest.raises(IndexError):
        p.IoUMetrics(19).compute_iou(la, bad)
    # empty batch
    e = ops.score(torch.zeros(0, 19, 4, 4), None, torch.zeros(0, 4, 4, dtype=torch.int64))
    assert int(e["bins"].sum()) == 0


@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean"])
@pytest.mark.parametrize("temp", [None, 1.7, 0.37])
def test_fuse_forward_bit_exact(pkg, strategy, temp):
    """awx_fuse_forward (EnsembleModel.forward's tensor, no statistics) == the reference's fusion expression ==
    awx_score's fused map, on vector-friendly and odd shapes and on a depth-head shaped [B,1,H,W] pair."""
    p, ops, _lib = pkg
    raw_w = torch.tensor([0.3, 0.9])
    w = of_.member_weights(raw_w)
    code = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF, "mean": _lib.FUSE_MEAN}[strategy]
    tt = None if temp is None else torch.tensor([temp])
    for seed, (b, c, h, w_) in enumerate([(2, 19, 48, 64), (1, 19, 7, 9), (3, 5, 5, 3), (2, 1, 33, 20)]):
        if c == 1 and strategy == "max_confidence":
            continue   # the depth heads are fused by the weighted / mean expressions only (model.py:471-478)
        gen = torch.Generator().manual_seed(40 + seed)
        la = torch.randn(b, c, h, w_, generator=gen) * 3
        lb = torch.randn(b, c, h, w_, generator=gen) * 3
        want = of_.fuse_logits(la, lb, strategy, raw_w, tt)
        got = ops.fuse_forward(la, lb, code, float(w[0]), float(w[1]), temp)
        assert got.is_cuda and torch.equal(got.cpu(), want), (strategy, temp, (b, c, h, w_))
        via_score = ops.score(la, lb, strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=temp, want_fused=True)["fused"]
        assert torch.equal(got, via_score)
    # unaligned views take the scalar path
    la = torch.randn(1, 19, 8, 33, generator=torch.Generator().manual_seed(1)).cuda()
    lb = torch.randn(1, 19, 8, 33, generator=torch.Generator().manual_seed(2)).cuda()
    a_off, b_off = la.flatten()[1:].reshape(-1)[:19 * 8 * 32].view(1, 19, 8, 32), lb.flatten()[3:][:19 * 8 * 32].view(1, 19, 8, 32)
    want = of_.fuse_logits(a_off.cpu(), b_off.cpu(), strategy, raw_w, tt)
    assert torch.equal(ops.fuse_forward(a_off, b_off, code, float(w[0]), float(w[1]), temp).cpu(), want)


@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean"])
def test_ensemble_non_finite_logits_follow_the_reference(pkg, strategy):
    """+-inf / NaN logits in either member: the prediction map is the arg-max of the reference's own fusion
    expression (0 * inf = NaN in max_confidence's `mask*l1 + (1-mask)*l2` included; NaN wins torch's arg-max),
    whatever member ends up being picked, and finite neighbours are untouched."""
    p, ops, _lib = pkg
    gen = torch.Generator().manual_seed(5)
    la = torch.randn(1, 19, 16, 32, generator=gen) * 2
    lb = torch.randn(1, 19, 16, 32, generator=gen) * 2
    la[0, :, 0, :] += 3 * torch.nn.functional.one_hot(torch.arange(32) % 19, 19).T   # member 1 confident on row 0
    lb[0, :, 1, :] += 3 * torch.nn.functional.one_hot(torch.arange(32) % 19, 19).T   # member 2 confident on row 1
    inf = float("inf")
    for row in (0, 1, 2):
        la[0, 4, row, 3], lb[0, 6, row, 5] = -inf, -inf       # -inf in member 1 / member 2
        la[0, 2, row, 9], lb[0, 8, row, 11] = inf, inf          # +inf
        la[0, 1, row, 15], lb[0, 3, row, 17] = float("nan"), float("nan")
        la[0, 5, row, 21], lb[0, 5, row, 21] = -inf, inf        # both members at one pixel
    tgt = torch.randint(0, 19, (1, 16, 32), generator=gen)
    raw_w = torch.tensor([0.3, 0.9])
    w = of_.member_weights(raw_w)
    code = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF, "mean": _lib.FUSE_MEAN}[strategy]
    want = of_.fuse_logits(la, lb, strategy, raw_w, torch.tensor([1.7]))
    kw = dict(strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=1.7, auroc_bins=4096)
    out = ops.score(la, lb, tgt, want_pred=torch.int64, want_fused=True, **kw)
    assert torch.equal(out["pred"].cpu(), want.argmax(1))
    got = out["fused"].cpu()
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(got, nan=0.0), torch.nan_to_num(want, nan=0.0))
    # the bins-only kernel counts the same confusion matrix
    fast = ops.read_bins(ops.score(la, lb, tgt, **kw)["bins"], 19, 15, 4096)
    slow = ops.read_bins(out["bins"], 19, 15, 4096)
    assert np.array_equal(fast.confusion, slow.confusion)
    assert np.array_equal(fast.confusion, om.confusion_matrix(want, tgt, 19).numpy())


def test_uint8_wrap_quirk_and_prediction_maps(pkg, golden):
    p, ops, _lib = pkg
    g = golden("metrics")
    la = torch.from_numpy(g["c19_u8_la"])
    t8 = torch.from_numpy(g["c19_u8_target"])
    pred = la.argmax(1)
    for preds, tg in ((pred, t8), (pred, t8.long()), (pred.to(torch.uint8), t8.long())):
        want = om.confusion_matrix(preds, tg, 19).numpy()
        cm, cnt = ops.confusion(preds, tg, 19)
        assert np.array_equal(cm.cpu().numpy(), want)
    # uint8 predictions with uint8 targets: the reference's index_add_ rejects the Byte source
    with pytest.raises(RuntimeError, match="same scalar type"):
        om.confusion_matrix(pred.to(torch.uint8), t8, 19)
    with pytest.raises(RuntimeError, match="same scalar type"):
        p.IoUMetrics(19).compute_iou(pred.to(torch.uint8), t8)
    rob = p.RobustnessMetrics(19)
    assert rob.compute_miou(pred, t8) == om.iou(pred, t8, 19)["mean_iou"]
    assert rob.compute_miou(la, t8.long()) == om.iou(la, t8.long(), 19)["mean_iou"]
    wm = rob.compute_weather_specific_metrics({"fog": pred, "hail": pred}, {"fog": t8, "hail": t8})
    assert list(wm) == ["miou_fog"]


def test_accumulation_and_determinism(pkg):
    """bins accumulate across calls (streaming) and are bit-identical run to run."""
    p, ops, _lib = pkg
    la, lb, tgt = _rand_case(21, 4, 19, 64, 64)
    kw = dict(strategy=_lib.FUSE_WEIGHTED, w0=0.4, w1=0.6, temperature=1.3, auroc_bins=2048)
    whole = ops.score(la, lb, tgt, **kw)["bins"].cpu()
    again = ops.score(la, lb, tgt, **kw)["bins"].cpu()
    assert torch.equal(whole, again)
    acc = ops.new_bins(19, 15, 2048)
    for i in range(4):
        ops.score(la[i:i + 1], lb[i:i + 1], tgt[i:i + 1], bins=acc, **kw)
    assert torch.equal(acc.cpu(), whole)


def test_largest_bin_configurations_route_and_agree(pkg, monkeypatch):
    """64 ECE bins x 8192 AUROC bins leave no room for the TMA ring next to the histograms: the call must route
    to the register-resident kernel (not fail); mid-size configurations must agree between both kernels."""
    p, ops, _lib = pkg
    la, lb, tgt = _rand_case(33, 2, 19, 96, 128)
    for eb, ab in ((64, 8192), (64, 4096), (32, 8192), (15, 8192), (64, 0)):
        kw = dict(strategy=_lib.FUSE_WEIGHTED, w0=0.3, w1=0.7, temperature=1.2, ece_bins=eb, auroc_bins=ab)
        monkeypatch.delenv("AWX_SCORE_KERNEL", raising=False)
        auto = ops.read_bins(ops.score(la, lb, tgt, **kw)["bins"], 19, eb, ab)
        monkeypatch.setenv("AWX_SCORE_KERNEL", "v1")
        v1 = ops.read_bins(ops.score(la, lb, tgt, **kw)["bins"], 19, eb, ab)
        assert np.array_equal(auto.confusion, v1.confusion), (eb, ab)
        for k in (_lib.CNT_VALID, _lib.CNT_CORRECT, _lib.CNT_PIXELS, _lib.CNT_NO_BIN):
            assert auto.counter(k) == v1.counter(k), (eb, ab, k)
        amb = auto.counter(_lib.CNT_ECE_AMBIG) + v1.counter(_lib.CNT_ECE_AMBIG)
        assert np.abs(auto.ece_count - v1.ece_count).sum() <= 2 * amb, (eb, ab)
        assert auto.ece_count.sum() == v1.ece_count.sum()
        np.testing.assert_allclose(auto.ece_conf_sum, v1.ece_conf_sum, rtol=1e-6, atol=1e-3)
        if ab:
            assert np.abs(auto.auroc_pos - v1.auroc_pos).sum() + np.abs(auto.auroc_neg - v1.auroc_neg).sum() <= 4


def test_full_size_properties(pkg):
    """1024x2048 frames: conservation laws that hold at any size (no oracle needed)."""
    p, ops, _lib = pkg
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(3)
    b, c, h, w = 2, 19, 1024, 2048
    la = torch.randn(b, c, h, w, device=dev, generator=gen)
    lb = torch.randn(b, c, h, w, device=dev, generator=gen)
    tgt = torch.randint(0, c, (b, h, w), device=dev, generator=gen).to(torch.uint8)
    tgt[0, :7] = 255
    out = ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=0.5, w1=0.5, temperature=1.7, auroc_bins=4096,
                    want_pred=torch.uint8)
    bins = ops.read_bins(out["bins"], c, 15, 4096)
    n_valid = int((tgt != 255).sum())
    assert bins.counter(_lib.CNT_PIXELS) == b * h * w
    assert bins.counter(_lib.CNT_VALID) == n_valid
    assert bins.confusion.sum() == n_valid
    assert bins.ece_count.sum() + bins.counter(_lib.CNT_NO_BIN) == n_valid
    assert bins.auroc_pos.sum() + bins.auroc_neg.sum() == n_valid
    assert bins.auroc_pos.sum() == bins.counter(_lib.CNT_ENS_WRONG)
    # the ECE's accuracy term is the arg-max of the PROBABILITIES, pixel accuracy that of the logits: they can
    # differ only where the kernel reported a tie of the top probabilities
    assert abs(int(bins.ece_correct.sum()) - bins.counter(_lib.CNT_CORRECT)) <= bins.counter(_lib.CNT_EPRED_AMBIG)
    for k in (_lib.CNT_ECE_AMBIG, _lib.CNT_EPRED_AMBIG, _lib.CNT_MARG_AMBIG):
        assert bins.counter(k) <= parity.amb_bound(n_valid, 5e-6), k
    # against torch on the device for the integer parts
    fused = (0.5 * la + 0.5 * lb) / torch.tensor([1.7], device=dev)  # tensor divisor: true division on CUDA
    assert torch.equal(out["pred"].long(), fused.argmax(1))
    cm, _ = ops.confusion(out["pred"].long(), tgt, c)  # int64 predictions: same promotion as inside awx_score
    assert np.array_equal(cm.cpu().numpy(), bins.confusion)


@pytest.mark.parametrize("mode", ["single", "weighted", "mean"])
@pytest.mark.parametrize("temp", [None, 1.7])
@pytest.mark.parametrize("ldt", [torch.uint8, torch.int64])
def test_full_size_bins_only_kernels_match_generic(pkg, mode, temp, ldt):
    """The streaming (bins-only) kernels -- compile-time division mode, LDS-free ECE binning, 19 consumer
    warps for one member -- against the generic kernel on 1024x2048 frames with a tail tile per image
    (2 097 152 is a multiple of neither 480 nor 608): every integer bin identical, up to the reported
    ambiguous pixels; and BOTH against the oracle on the same 4.2 Mpixel: the true mismatch of the ECE bins and of
    the ensemble-wrong count must not exceed the reported ambiguous pixels, themselves < 2e-6 of the frame."""
    p, ops, _lib = pkg
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(11)
    b, c, h, w = 2, 19, 1024, 2048
    la = torch.randn(b, c, h, w, device=dev, generator=gen) * 2
    lb = torch.randn(b, c, h, w, device=dev, generator=gen) * 2
    tgt = torch.randint(0, c, (b, h, w), device=dev, generator=gen).to(ldt)
    tgt[1, -5:] = 255
    code = {"single": _lib.FUSE_SINGLE, "weighted": _lib.FUSE_WEIGHTED, "mean": _lib.FUSE_MEAN}[mode]
    nb = 0 if mode == "single" else 4096
    w = of_.member_weights(torch.tensor([0.0, 1.0]))
    kw = dict(strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=temp, auroc_bins=nb)
    second = None if mode == "single" else lb
    fast = ops.read_bins(ops.score(la, second, tgt, **kw)["bins"], c, 15, nb)
    slow = ops.read_bins(ops.score(la, second, tgt, want_pred=torch.uint8, want_conf=True, **kw)["bins"], c, 15, nb)
    assert np.array_equal(fast.confusion, slow.confusion)
    for k in (_lib.CNT_VALID, _lib.CNT_CORRECT, _lib.CNT_BAD_LABEL, _lib.CNT_ENS_WRONG, _lib.CNT_PIXELS, _lib.CNT_NO_BIN):
        assert fast.counter(k) == slow.counter(k), k
    amb = fast.counter(_lib.CNT_ECE_AMBIG) + slow.counter(_lib.CNT_ECE_AMBIG)
    eamb = fast.counter(_lib.CNT_EPRED_AMBIG) + slow.counter(_lib.CNT_EPRED_AMBIG)
    assert np.abs(fast.ece_count - slow.ece_count).sum() <= 2 * amb
    assert np.abs(fast.ece_correct - slow.ece_correct).sum() <= 2 * amb + eamb
    np.testing.assert_allclose(fast.ece_conf_sum, slow.ece_conf_sum, rtol=1e-6, atol=1e-3)
    # the oracle on the same frames
    la_c, lb_c, tgt_c = la.cpu(), lb.cpu(), tgt.cpu()
    strategy = {"single": None, "weighted": "weighted_average", "mean": "mean"}[mode]
    want = la_c if strategy is None else of_.fuse_logits(la_c, lb_c, strategy, torch.tensor([0.0, 1.0]),
                                                         None if temp is None else torch.tensor([temp]))
    if strategy is None and temp is not None:
        want = la_c / torch.tensor([temp])
    n_valid = int((tgt_c != 255).sum())
    ref = om.ece(want, tgt_c)
    assert np.array_equal(fast.confusion, om.confusion_matrix(want, tgt_c, c).numpy())
    cap = parity.genuine_ece_near_edges(want, tgt_c)
    assert cap <= 1e-5 * n_valid      # the data itself: a few pixels per million sit on an edge
    for b_ in (fast, slow):
        parity.assert_ece_parity(b_, ref, n_valid, _lib, cap=cap)
        if nb:
            wrong = int(((om.mean_prob_prediction([la_c, lb_c]) != tgt_c) & (tgt_c != 255)).sum())
            parity.assert_ens_wrong_parity(b_, wrong, n_valid, _lib)
    if nb:
        # the MI value of a pixel is the same arithmetic in both kernels; a handful of pixels may sit on a
        # histogram edge after the different instruction scheduling -- none are expected
        assert np.abs(fast.auroc_pos - slow.auroc_pos).sum() + np.abs(fast.auroc_neg - slow.auroc_neg).sum() <= 4


@pytest.mark.parametrize("tag,n", [("m3", 3), ("m4", 4)])
def test_member_lists_golden(pkg, golden, tag, n):
    """EnsembleDisagreementMetrics on lists of more than two members (awx_members_n) against the reference."""
    p, ops, _lib = pkg
    g = golden("prep")
    members = [torch.from_numpy(g[f"{tag}_member{k}"]) for k in range(n)]
    targets = torch.from_numpy(g[f"{tag}_targets"])
    ens = p.EnsembleDisagreementMetrics()
    _close(ens.compute_disagreement_map(members).cpu().numpy(), g[f"{tag}_mi"])
    _close(ens.compute_variance_map(members).cpu().numpy(), g[f"{tag}_var"], rtol=1e-5, atol=1e-7)
    val, bound = ens.compute_disagreement_auroc(members, targets, return_bound=True)
    assert abs(val - float(g[f"{tag}_auroc"])) <= bound + 1e-6 and bound < 2e-3
    # two members through the list kernel == the fused kernel's maps (same formulas, different code)
    two = ops.members_n(members[:2], want_mi=True, want_var=True)
    _close(two["mi"].cpu().numpy(), ops.score(members[0], members[1], strategy=_lib.FUSE_MEAN, want_mi=True)["mi"].cpu().numpy())
    _close(two["var"].cpu().numpy(), ops.member_variance(members[0], members[1]).cpu().numpy(), rtol=1e-5, atol=1e-7)
    with pytest.raises(ValueError, match="at least 2"):
        ens.compute_disagreement_map(members[:1])
