"""evaluate_model drop-in (scripts/evaluate.py:134-274) on the GPU against the oracle's restatement of
the reference's end-of-loop metric calls on the concatenated tensors."""

import numpy as np
import pytest
import torch

from oracle import metrics as om, fusion as of_

pytestmark = pytest.mark.gpu

C, H, W = 19, 48, 64
CONDS = ["clean", "fog", "rain", "snow", "night"]


class _Config:
    def __init__(self, d):
        self.d = d

    def get(self, k, default=None):
        return self.d.get(k, default)


def _loader(seed, n_batches=4, bsz=3, with_other=True):
    gen = torch.Generator().manual_seed(seed)
    rng = np.random.RandomState(seed)
    names = CONDS + (["hail"] if with_other else [])
    out = []
    for _ in range(n_batches):
        out.append({
            "image": torch.rand(bsz, 3, H, W, generator=gen),
            "label": torch.randint(0, C, (bsz, H, W), generator=gen),
            "weather_condition": [names[i] for i in rng.randint(0, len(names), bsz)],
            "_la": torch.randn(bsz, C, H, W, generator=gen) * 2,
            "_lb": torch.randn(bsz, C, H, W, generator=gen) * 2,
        })
    out[0]["weather_condition"][0] = "clean"   # the degradation ratios need a clean frame
    out[0]["label"][0, :3] = 255
    return out


class _Ensemble(torch.nn.Module):
    """Stands in for EnsembleModel: returns the member logits stashed in the batch."""

    def __init__(self, batches, strategy="weighted_average", ts=True):
        super().__init__()
        self.batches, self.i = batches, 0
        self.ensemble_strategy, self.temperature_scaling = strategy, ts
        self.segformer = torch.nn.Identity()   # evaluate.py:196: ensemble statistics only for models with this member
        self.ensemble_weights = torch.nn.Parameter(torch.tensor([0.3, 0.9]))
        self.temperature = torch.nn.Parameter(torch.tensor([1.7]))

    def forward(self, x):
        b = self.batches[self.i]
        self.i += 1
        la, lb = b["_la"].cuda(), b["_lb"].cuda()
        return {"segmentation": None, "segformer_seg": la, "deeplabv3plus_seg": lb}


class _Single(torch.nn.Module):
    def __init__(self, batches):
        super().__init__()
        self.batches, self.i = batches, 0

    def forward(self, x):
        b = self.batches[self.i]
        self.i += 1
        return {"segmentation": b["_la"].cuda()}


def _expected(batches, fuse):
    logits = torch.cat([fuse(b["_la"], b["_lb"]) for b in batches])
    tg = torch.cat([b["label"] for b in batches])
    weather = sum((b["weather_condition"] for b in batches), [])
    res = {"overall_miou": om.iou(logits, tg, C)["mean_iou"], "expected_calibration_error": om.ece(logits, tg)["ece"]}
    mious = {}
    for c in CONDS:
        idx = [i for i, wname in enumerate(weather) if wname == c]
        if idx:
            mious[c] = om.iou(logits[idx], tg[idx], C)["mean_iou"]
            res[f"miou_{c}"] = mious[c]
            res[f"ece_{c}"] = om.ece(logits[idx], tg[idx])["ece"]
    degr = []
    for c in CONDS[1:]:
        if c in mious and "clean" in mious:
            res[f"robustness_degradation_{c}"] = om.degradation_ratio(mious["clean"], mious[c])
            degr.append(res[f"robustness_degradation_{c}"])
    if degr:
        res["robustness_degradation_ratio"] = np.mean(degr)
    return res


def _check(res, want):
    for k, v in want.items():
        assert k in res, k
        if k.startswith("ece") or k == "expected_calibration_error":
            np.testing.assert_allclose(res[k], v, rtol=1e-5, atol=1e-8, err_msg=k)
        else:
            assert res[k] == v, (k, res[k], v)


@pytest.mark.parametrize("strategy,ts", [("weighted_average", True), ("mean", False), ("max_confidence", True)])
def test_evaluate_model_ensemble(strategy, ts):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import evaluate_model
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import RobustnessMetrics
    batches = _loader(3)
    model = _Ensemble(batches, strategy, ts)
    cfg = _Config({"data.weather_conditions": CONDS})
    res = evaluate_model(model, batches, RobustnessMetrics(C), torch.device("cuda"), cfg)
    raw_w, temp = torch.tensor([0.3, 0.9]), torch.tensor([1.7])
    want = _expected(batches, lambda a, b: of_.fuse_logits(a, b, strategy, raw_w, temp if ts else None))
    _check(res, want)
    tg = torch.cat([b["label"] for b in batches])
    exact = om.disagreement_auroc([torch.cat([b["_la"] for b in batches]), torch.cat([b["_lb"] for b in batches])], tg)
    assert abs(res["ensemble_disagreement_auroc"] - exact) <= 2e-3
    assert not any(k.startswith("miou___") or "hail" in k for k in res)


class _DDPLike(torch.nn.Module):
    """What DistributedDataParallel looks like from outside: the model sits in ``.module``."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        return self.module(x)


@pytest.mark.parametrize("tag", ["e_weighted", "e_u8", "e_maxconf", "e_mean_not", "e_wrapped", "e_single"])
def test_evaluate_model_matches_a_run_of_the_reference(golden, tag):
    """tests/golden/evaluate.npz holds result dicts of the REFERENCE's evaluate_model (scripts/evaluate.py:134-274)
    on its own EnsembleModel with trained-looking fusion weights ([0.3, 0.9]) and temperature 1.7, int64 / uint8
    labels, mixed weather, a condition the config does not list, the model behind a DDP-style wrapper, and a
    single-member model.  The drop-in must return the same keys; mIoU and degradation ratios ==, ECE to 1e-5,
    AUROC within the histogram bound."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import evaluate_model
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import EnsembleModel, RobustnessMetrics
    from trainer_fixture import evaluate_fixture_batches, ReplayMember
    g = golden("evaluate")
    _, _, _, c, _, _, _, ts, wrapped, single = (int(x) for x in g[f"{tag}_args"])
    batches = evaluate_fixture_batches(g, tag)
    dev = torch.device("cuda")
    if single:
        model = ReplayMember(batches, "_la", dev)
    else:
        model = EnsembleModel(num_classes=c, include_depth=False, ensemble_strategy=str(g[f"{tag}_strategy"]),
                              temperature_scaling=bool(ts), segformer=ReplayMember(batches, "_la", dev),
                              deeplabv3plus=ReplayMember(batches, "_lb", dev))
        with torch.no_grad():
            model.ensemble_weights.copy_(torch.tensor([0.3, 0.9]))
            if ts:
                model.temperature.copy_(torch.tensor([1.7]))
        model = model.to(dev)
        if wrapped:
            model = _DDPLike(model)
    res = evaluate_model(model, batches, RobustnessMetrics(c), dev, _Config({"data.weather_conditions": CONDS}))
    want = dict(zip((str(k) for k in g[f"{tag}_keys"]), g[f"{tag}_vals"]))
    assert sorted(res) == sorted(want)
    for k, v in want.items():
        if k.startswith("ece") or k == "expected_calibration_error":
            np.testing.assert_allclose(res[k], v, rtol=1e-5, atol=1e-8, err_msg=k)
        elif k == "ensemble_disagreement_auroc":
            assert abs(res[k] - v) <= 2e-3, (k, res[k], v)
        else:
            assert float(res[k]) == v, (k, res[k], v)


def test_evaluate_model_scores_the_models_own_fusion_when_it_cannot_restate_it():
    """A model whose ``outputs['segmentation']`` is NOT what its fusion attributes say (here: an extra bias) must be
    scored on its own fused logits -- the reference scores outputs['segmentation'] (evaluate.py:178-179) -- with the
    members feeding only the disagreement histogram; and a model without the fusion attributes likewise."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import evaluate_model
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import RobustnessMetrics
    batches = _loader(21)
    bias = torch.linspace(-1, 1, C).view(1, C, 1, 1)

    class _Odd(_Ensemble):
        def __init__(self, batches, with_attrs):
            super().__init__(batches)
            self.segformer = torch.nn.Identity()       # evaluate.py:196 keys the ensemble statistics on this attribute
            if not with_attrs:
                del self.ensemble_weights, self.temperature
                del self.ensemble_strategy, self.temperature_scaling

        def forward(self, x):
            out = super().forward(x)
            out["segmentation"] = (0.5 * out["segformer_seg"] + 0.5 * out["deeplabv3plus_seg"]) + bias.cuda()
            return out

    want = _expected(batches, lambda a, b: (0.5 * a + 0.5 * b) + bias)
    for with_attrs in (True, False):
        res = evaluate_model(_Odd(batches, with_attrs), batches, RobustnessMetrics(C), torch.device("cuda"),
                             _Config({"data.weather_conditions": CONDS}))
        _check(res, want)
        tg = torch.cat([b["label"] for b in batches])
        exact = om.disagreement_auroc([torch.cat([b["_la"] for b in batches]), torch.cat([b["_lb"] for b in batches])], tg)
        assert abs(res["ensemble_disagreement_auroc"] - exact) <= 2e-3


def test_evaluate_model_single_member():
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import evaluate_model
    batches = _loader(5, with_other=False)
    res = evaluate_model(_Single(batches), batches, None, torch.device("cuda"), _Config({"data.weather_conditions": CONDS}))
    _check(res, _expected(batches, lambda a, b: a))
    assert "ensemble_disagreement_auroc" not in res


def test_corrupt_score_single_call_equals_two_calls():
    """awx_corrupt_score (one C call per condition) == awx_corrupt followed by awx_score."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
        WeatherDegradationTransforms)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import StreamingEvaluator
    h, w, b = 64, 96, 3
    gen = torch.Generator().manual_seed(2)
    imgs = torch.randint(0, 255, (b, h, w, 3), generator=gen, dtype=torch.uint8).cuda()
    la = (torch.randn(b, C, h, w, generator=gen) * 2).cuda()
    lb = (torch.randn(b, C, h, w, generator=gen) * 2).cuda()
    lab = torch.randint(0, C, (b, h, w), generator=gen).to(torch.uint8).cuda()
    t = WeatherDegradationTransforms(seed=4)
    for kind in ("rain", "night"):
        draws = [t.draw(kind, h, w) for _ in range(b)]
        prm, fld, items = t.pack(draws, h, w)
        fld_d = None if fld is None else torch.from_numpy(fld).cuda()
        items_d = None if items is None else torch.from_numpy(items).cuda()
        ws = ops.corrupt_workspace(b, h, w)
        ev1 = StreamingEvaluator(C, ("x",), auroc_bins=512, ensemble_weights=(0.3, 0.9), temperature=1.7)
        ev2 = StreamingEvaluator(C, ("x",), auroc_bins=512, ensemble_weights=(0.3, 0.9), temperature=1.7)
        out1, out2 = torch.empty_like(imgs), torch.empty_like(imgs)
        ev1.update_corrupted("x", imgs, prm, fld_d, items_d, out1, ws, la, lb, lab)
        ops.corrupt(imgs, prm, fld_d, items_d, out=out2, workspace=ws)
        ev2.update("x", la, lb, lab)
        assert torch.equal(out1, out2) and torch.equal(ev1.bins, ev2.bins)
        assert int(ev1.bins.sum()) > 0


def test_chunked_updates_equal_one_update():
    """Scoring a batch in chunks gives the same bins as one launch: every count identical, the confidence sums
    identical in canonical form (hi/lo words are plain accumulators and split differently per launch)."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import StreamingEvaluator
    gen = torch.Generator().manual_seed(8)
    b, h, w = 6, 96, 160
    la = (torch.randn(b, C, h, w, generator=gen) * 2).cuda()
    lb = (torch.randn(b, C, h, w, generator=gen) * 2).cuda()
    lab = torch.randint(0, C, (b, h, w), generator=gen).to(torch.uint8).cuda()
    kw = dict(auroc_bins=4096, ensemble_weights=(0.3, 0.9), temperature=1.7)
    one, many = StreamingEvaluator(C, ("x",), **kw), StreamingEvaluator(C, ("x",), **kw)
    one.update("x", la, lb, lab)
    for k in range(0, b, 2):
        many.update("x", la[k:k + 2], lb[k:k + 2], lab[k:k + 2])
    assert torch.equal(one.canonical_bins(), many.canonical_bins())
    assert one.finalize() == many.finalize()


# ------------------------------------------------------------- trainer twin: validate_epoch (trainer.py:377-511)
def test_trainer_fog_density_maps_bit_exact(golden):
    """estimate_fog_density draws from the global torch CPU generator in the reference's order; the affine map
    runs on the device with the CPU's two roundings."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import estimate_fog_density
    g = golden("trainer")
    names = [str(x) for x in g["fd_names"]]
    torch.manual_seed(123)
    got = estimate_fog_density({"weather_condition": names, "image": torch.zeros(len(names), 3, 20, 28)})
    assert got.is_cuda and got.dtype == torch.float32
    np.testing.assert_array_equal(got.cpu().numpy(), g["fd_maps"])
    assert estimate_fog_density({"image": torch.zeros(1, 3, 4, 4)}) is None


@pytest.mark.parametrize("tag", ["v_depth", "v_nodepth", "v_focal", "v_plain"])
def test_validate_epoch_matches_reference(golden, tag):
    """The streaming validation pass against the reference's own validate_epoch (fixture made by running it):
    mIoU values identical (integer confusion matrices, same finalisation), loss means to fp32 rounding."""
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import FogDensityAwareLoss, RobustnessMetrics
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation import validate_epoch
    from trainer_fixture import ReplayModel, trainer_fixture_batches
    g = golden("trainer")
    seed, nb, bsz, c = (int(x) for x in g[f"{tag}_args"][:4])
    batches = trainer_fixture_batches(g, tag)
    dev = torch.device("cuda")
    if tag == "v_plain":
        loss_fn = torch.nn.CrossEntropyLoss()
    else:
        loss_fn = FogDensityAwareLoss(base_loss=str(g[f"{tag}_base"]))
        torch.manual_seed(1000 + seed)
    res = validate_epoch(ReplayModel(batches, dev), batches, loss_fn, RobustnessMetrics(c), dev)
    assert sorted(res) == [str(k) for k in g[f"{tag}_keys"]]
    for k, v in zip((str(k) for k in g[f"{tag}_keys"]), g[f"{tag}_vals"]):
        if "miou" in k or k == "val_samples":
            assert float(res[k]) == v, k
        else:
            assert float(res[k]) == pytest.approx(v, rel=2e-6), k
