"""The CPU oracle against the LIVE reference, on fresh seeds (not the ones frozen in tests/golden/*.npz).

Runs only where ``/root/reference`` exists (the build container); on the GPU box, where it does not, every test
here is skipped and the committed golden vectors (tests/test_oracle_golden.py) carry the pin.  Each check
executes the reference's own code through tests/refshim.py and requires the oracle's restatement to return the
same bits on the same inputs and the same RNG state.
"""

import numpy as np
import pytest
import torch

import refshim
from oracle import weather as ow, metrics as om, fusion as of_, loss as ol, prep as op

pytestmark = pytest.mark.skipif(not refshim.available(), reason="the reference is only present in the build container")


@pytest.mark.parametrize("kind", ["fog", "rain", "snow", "night"])
@pytest.mark.parametrize("seed,intensity", [(101, None), (102, 0.35)])
def test_weather_live(kind, seed, intensity):
    pre = refshim.preprocessing()
    rng = np.random.RandomState(seed)
    image = rng.randint(0, 255, (44, 60, 3)).astype(np.uint8)
    want = pre.WeatherDegradationTransforms(seed=seed).apply_weather_effect(image.copy(), kind, intensity)
    state_ref = np.random.get_state()[1].copy()
    np.random.seed(seed)
    got = ow.apply(image.copy(), kind, intensity)
    assert np.array_equal(got, want)
    assert np.array_equal(np.random.get_state()[1], state_ref), "the oracle consumed the RNG differently"


def test_synthetic_depth_live():
    pre = refshim.preprocessing()
    want = pre.WeatherDegradationTransforms(seed=5)._generate_synthetic_depth(37, 53)
    np.random.seed(5)
    assert np.array_equal(ow.depth_from_noise(np.random.normal(0, 10, (37, 53))), want)


@pytest.mark.parametrize("label_dtype", [torch.int64, torch.uint8])
def test_metrics_live(label_dtype):
    met = refshim.metrics()
    gen = torch.Generator().manual_seed(77)
    c = 19
    la = torch.randn(2, c, 20, 28, generator=gen) * 2
    lb = torch.randn(2, c, 20, 28, generator=gen) * 2
    tgt = torch.randint(0, c, (2, 20, 28), generator=gen)
    tgt[torch.rand(2, 20, 28, generator=gen) < 0.05] = 255
    tgt = tgt.to(label_dtype)
    iou = met.IoUMetrics(c)
    r, o = iou.compute_iou(la, tgt), om.iou(la, tgt, c)
    assert r["mean_iou"] == o["mean_iou"] and np.array_equal(r["per_class_iou"], o["per_class_iou"])
    assert iou.compute_pixel_accuracy(la, tgt) == om.pixel_accuracy(la, tgt)
    d = met.ConfidenceCalibration().compute_ece(la, tgt, return_details=True)
    e = om.ece(la, tgt)
    assert d["ece"] == e["ece"] and d["bin_details"] == e["bin_details"]
    ens = met.EnsembleDisagreementMetrics()
    assert torch.equal(ens.compute_disagreement_map([la, lb]), om.mi_map([la, lb]))
    assert torch.equal(ens.compute_variance_map([la, lb]), om.variance_map([la, lb]))
    assert torch.equal(ens.compute_jensen_shannon_divergence(la, lb), of_.reverse_kl_disagreement(la, lb))
    assert ens.compute_disagreement_auroc([la, lb], tgt) == om.disagreement_auroc([la, lb], tgt)
    rob = met.RobustnessMetrics(c)
    assert rob.compute_robustness_degradation_ratio(0.5, 0.3) == om.degradation_ratio(0.5, 0.3)
    assert rob.compute_robustness_degradation_ratio(0.0, 0.3) == om.degradation_ratio(0.0, 0.3)


@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean"])
@pytest.mark.parametrize("ts", [True, False])
def test_fusion_live(strategy, ts):
    gen = torch.Generator().manual_seed(9)
    l1 = torch.randn(2, 19, 12, 20, generator=gen) * 3
    l2 = torch.randn(2, 19, 12, 20, generator=gen) * 3
    raw_w, temp = torch.tensor([-0.4, 0.7]), torch.tensor([2.3])
    ens = refshim.ensemble_with_fixed_members(l1, l2, strategy, ts, raw_weights=raw_w, temperature=temp)
    with torch.no_grad():
        want = ens(torch.zeros(2, 3, 12, 20))["segmentation"]
        dis = ens.get_ensemble_disagreement(torch.zeros(2, 3, 12, 20))
    assert torch.equal(of_.fuse_logits(l1, l2, strategy, raw_w, temp if ts else None), want)
    assert torch.equal(of_.reverse_kl_disagreement(l1, l2), dis)


@pytest.mark.parametrize("base,with_fd,with_depth", [("cross_entropy", True, True), ("focal", True, False),
                                                     ("cross_entropy", False, True)])
def test_loss_live(base, with_fd, with_depth):
    m = refshim.model()
    gen = torch.Generator().manual_seed(13)
    b, c, h, w = 2, 19, 10, 14
    logits = torch.randn(b, c, h, w, generator=gen)
    lab = torch.randint(0, c, (b, h, w), generator=gen)
    fd = torch.rand(b, h, w, generator=gen) if with_fd else None
    dp = torch.rand(b, 1, h, w, generator=gen) * 40
    dt = torch.rand(b, h, w, generator=gen) * 40

    def run(fn):
        x = logits.clone().requires_grad_(True)
        d = dp.clone().requires_grad_(True)
        preds, tg = {"segmentation": x}, {"label": lab}
        if with_depth:
            preds["depth"], tg["depth"] = d, dt
        out = fn(preds, tg, fd)
        out["total_loss"].backward()
        return out, x.grad, d.grad

    ro, rx, rd = run(m.FogDensityAwareLoss(base_loss=base))
    oo, ox, od = run(lambda p_, t_, f_: ol.fog_loss(p_, t_, f_, base_loss=base))
    for k in ("total_loss", "segmentation_loss"):
        assert torch.equal(ro[k].detach(), oo[k].detach()), k
    assert torch.equal(rx, ox)
    if with_depth:
        assert torch.equal(rd, od)


def test_calibration_and_prep_live():
    met = refshim.metrics()
    pre = refshim.preprocessing()
    gen = torch.Generator().manual_seed(21)
    logits = torch.randn(400, 19, generator=gen) * 3
    tgt = torch.randint(0, 19, (400,), generator=gen)
    assert met.ConfidenceCalibration().optimize_temperature(logits, tgt) == op.optimize_temperature(logits, tgt)
    rng = np.random.RandomState(4)
    image = rng.randint(0, 255, (32, 48, 3)).astype(np.uint8)
    depth = rng.rand(32, 48) * 80 + 1
    t = pre.WeatherDegradationTransforms(seed=0)
    assert np.array_equal(t.get_fog_density_map(image, depth), op.fog_density_map(image, depth))
    est = pre.DepthEstimationPreprocessor()
    assert np.array_equal(est.estimate_depth(image), op.estimate_depth(image))
