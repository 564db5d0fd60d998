"""The CPU oracle against golden vectors produced by the reference itself.

Bit-exact (``==``) everywhere: the oracle calls the same library routines in the same
order as the reference, so on the same library versions the results are identical.
If the installed versions differ from those stored in the fixture the comparison
relaxes to the tolerances SURVEY.md section 8d states and says so.
"""

import json

import numpy as np
import pytest
import torch

from oracle import weather as ow, metrics as om, fusion as of_, loss as ol


def _same_versions(g):
    import cv2, scipy, sklearn
    have = {"numpy": np.__version__, "cv2": cv2.__version__, "scipy": scipy.__version__,
            "torch": torch.__version__, "sklearn": sklearn.__version__}
    return json.loads(str(g["versions"])) == have


def _eq(a, b, exact, atol=0.0, rtol=0.0):
    a, b = np.asarray(a), np.asarray(b)
    if exact:
        assert np.array_equal(a, b, equal_nan=True), f"max abs diff {np.abs(a.astype(np.float64) - b).max()}"
    else:
        np.testing.assert_allclose(a, b, atol=atol, rtol=rtol)


# ------------------------------------------------------------------------------ weather
@pytest.mark.parametrize("tag", ["s", "m"])
@pytest.mark.parametrize("kind", ["fog", "rain", "snow", "night"])
@pytest.mark.parametrize("seed,intensity", [(42, None), (43, 0.5), (44, 0.9)])
def test_weather_matches_reference(golden, tag, kind, seed, intensity):
    g = golden("weather")
    exact = _same_versions(g)
    np.random.seed(seed)
    got = ow.apply(g[f"{tag}_image"].copy(), kind, intensity)
    want = g[f"{tag}_{kind}_seed{seed}"]
    assert got.dtype == np.uint8 and got.shape == want.shape
    if exact:
        assert np.array_equal(got, want)
    else:
        assert np.abs(got.astype(int) - want.astype(int)).max() <= 1


@pytest.mark.parametrize("kind,intensity", [("rain", 0.8), ("snow", 0.7)])
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_weather_border_overlays(golden, kind, intensity, seed):
    g = golden("weather")
    np.random.seed(seed)
    got = ow.apply(g["b_image"].copy(), kind, intensity)
    want = g[f"b_{kind}_seed{seed}"]
    assert np.abs(got.astype(int) - want.astype(int)).max() <= (0 if _same_versions(g) else 1)


@pytest.mark.parametrize("tag,shape", [("s", (40, 56)), ("m", (96, 160))])
def test_synthetic_depth(golden, tag, shape):
    g = golden("weather")
    np.random.seed(11)
    d = ow.depth_from_noise(np.random.normal(0, 10, shape))
    _eq(d, g[f"{tag}_depth_seed11"], _same_versions(g), rtol=1e-12)


def test_clean_is_alias_and_unknown_raises():
    img = np.zeros((4, 4, 3), np.uint8)
    assert ow.apply(img, "clean") is img
    with pytest.raises(ValueError, match="Unknown weather type"):
        ow.apply(img, "hail")


# ------------------------------------------------------------------------------ metrics
CASES = [("c5", 5), ("c19_i64_ign", 19), ("c19_u8", 19)]


@pytest.mark.parametrize("tag,c", CASES)
def test_iou_and_accuracy(golden, tag, c):
    g = golden("metrics")
    exact = _same_versions(g)
    la = torch.from_numpy(g[f"{tag}_la"])
    tgt = torch.from_numpy(g[f"{tag}_target"])
    r = om.iou(la, tgt, c)
    _eq(r["mean_iou"], g[f"{tag}_mean_iou"], exact, rtol=1e-6)
    _eq(r["per_class_iou"], g[f"{tag}_per_class_iou"], exact, rtol=1e-6)
    assert np.array_equal(r["valid_classes"], g[f"{tag}_valid_classes"])
    assert om.pixel_accuracy(la, tgt) == float(g[f"{tag}_pixel_accuracy"])


@pytest.mark.parametrize("tag,c", CASES)
def test_ece(golden, tag, c):
    g = golden("metrics")
    exact = _same_versions(g)
    la = torch.from_numpy(g[f"{tag}_la"])
    tgt = torch.from_numpy(g[f"{tag}_target"])
    d = om.ece(la, tgt)
    _eq(d["ece"], g[f"{tag}_ece"], exact, rtol=1e-5)
    _eq(d["ece"], g[f"{tag}_ece_scalar"], exact, rtol=1e-5)
    for key in ("accuracy", "confidence", "proportion", "error", "bin_lower", "bin_upper"):
        _eq([x[key] for x in d["bin_details"]], g[f"{tag}_ece_{key}"], exact, rtol=1e-5, atol=1e-7)
    _eq(d["overall_accuracy"], g[f"{tag}_ece_overall_accuracy"], exact, rtol=1e-6)
    _eq(d["overall_confidence"], g[f"{tag}_ece_overall_confidence"], exact, rtol=1e-6)
    _eq(om.reliability_points(d["bin_details"])["bin_centers"], g[f"{tag}_rel_centers"], exact, rtol=1e-6)
    # integer bins are consistent with the reference's fp32 proportions
    n = int((tgt != 255).sum())
    np.testing.assert_allclose(d["count"] / n, g[f"{tag}_ece_proportion"], rtol=1e-6, atol=1e-9)


@pytest.mark.parametrize("tag,c", CASES)
def test_disagreement(golden, tag, c):
    g = golden("metrics")
    exact = _same_versions(g)
    la = torch.from_numpy(g[f"{tag}_la"])
    lb = torch.from_numpy(g[f"{tag}_lb"])
    tgt = torch.from_numpy(g[f"{tag}_target"])
    _eq(om.mi_map([la, lb]).numpy(), g[f"{tag}_mi"], exact, rtol=1e-5, atol=1e-6)
    _eq(om.variance_map([la, lb]).numpy(), g[f"{tag}_var"], exact, rtol=1e-5, atol=1e-7)
    _eq(of_.reverse_kl_disagreement(la, lb).numpy(), g[f"{tag}_js"], exact, rtol=1e-5, atol=1e-6)
    _eq(om.disagreement_auroc([la, lb], tgt), g[f"{tag}_auroc"], exact, rtol=1e-9)
    with pytest.raises(ValueError, match="Need at least 2 predictions"):
        om.mi_map([la])


def test_uint8_label_wrap_quirk(golden):
    """targets*C+pred wraps mod 256 for uint8 labels >= 14: the uint8 and
    int64 confusion matrices of the same data must differ, and the oracle keeps both."""
    g = golden("metrics")
    la = torch.from_numpy(g["c19_u8_la"])
    t8 = torch.from_numpy(g["c19_u8_target"])
    cm8 = om.confusion_matrix(la, t8, 19)
    cm64 = om.confusion_matrix(la, t8.long(), 19)
    assert cm8.sum() == cm64.sum() == t8.numel()
    assert not torch.equal(cm8, cm64)
    assert cm8[14:].sum() == 0 and cm64[14:].sum() > 0


def test_degradation_and_streaming_helpers(golden):
    g = golden("metrics")
    got = [om.degradation_ratio(a, b) for a, b in ((0.5, 0.4), (0.0, 0.3), (0.4, 0.5), (0.78, 0.65))]
    assert np.array_equal(np.array(got), g["degr"])
    # binned AUROC equals the exact one when every score has its own bin
    pos = np.array([0, 1, 0, 2]); neg = np.array([3, 0, 1, 0])
    a, bound = om.auroc_from_histogram(pos, neg)
    from sklearn.metrics import roc_auc_score
    y = [0, 0, 0, 1, 0, 1, 1]; s = [0, 0, 0, 1, 2, 3, 3]
    assert abs(a - roc_auc_score(y, s)) < 1e-12 and bound == 0.0
    e = om.ece_edges(15).numpy()
    idx = om.ece_bin_index(np.array([0.0, 1.0, e[3], np.nextafter(e[3], np.float32(1)), np.nan], np.float32), e)
    assert idx.tolist() == [-1, 14, 2, 3, -1]


# ------------------------------------------------------------------------------- fusion
@pytest.mark.parametrize("strategy", ["weighted_average", "max_confidence", "mean_anything"])
@pytest.mark.parametrize("ts", [True, False])
def test_fusion(golden, strategy, ts):
    g = golden("fusion")
    exact = _same_versions(g)
    l1, l2 = torch.from_numpy(g["l1"]), torch.from_numpy(g["l2"])
    d1, d2 = torch.from_numpy(g["d1"]), torch.from_numpy(g["d2"])
    raw_w, temp = torch.from_numpy(g["raw_w"]), torch.from_numpy(g["temp"])
    tag = f"{strategy}_{'T' if ts else 'noT'}"
    _eq(of_.fuse_logits(l1, l2, strategy, raw_w, temp if ts else None).numpy(), g[f"{tag}_seg"], exact, rtol=1e-6)
    _eq(of_.fuse_depth(d1, d2, strategy, raw_w).numpy(), g[f"{tag}_depth"], exact, rtol=1e-6)
    _eq(of_.reverse_kl_disagreement(l1, l2).numpy(), g[f"{tag}_dis"], exact, rtol=1e-5, atol=1e-6)


# --------------------------------------------------------------------------------- loss
VARIANTS = {
    "ce_fd_depth": ("cross_entropy", True, True, True),
    "ce_fd_nodepth": ("cross_entropy", True, False, False),
    "ce_nofd_nodepth": ("cross_entropy", False, False, False),
    "focal_fd_depth": ("focal", True, True, True),
    "ce_pathB": ("cross_entropy", False, True, True),
    "ce_fd_dpred_only": ("cross_entropy", True, True, False),
}


@pytest.mark.parametrize("tag", sorted(VARIANTS))
def test_loss_forward_backward(golden, tag):
    g = golden("loss")
    exact = _same_versions(g)
    base, use_fd, dpred, dtgt = VARIANTS[tag]
    lg = torch.from_numpy(g["logits"]).clone().requires_grad_(True)
    dp = torch.from_numpy(g["depth"]).clone().requires_grad_(True)
    pred = {"segmentation": lg}
    tgt = {"label": torch.from_numpy(g["label"])}
    if dpred:
        pred["depth"] = dp
    if dtgt:
        tgt["depth"] = torch.from_numpy(g["dtgt"])
    r = ol.fog_loss(pred, tgt, torch.from_numpy(g["fd"]) if use_fd else None, base_loss=base)
    r["total_loss"].backward()
    _eq(r["total_loss"].item(), g[f"{tag}_total"], exact, rtol=1e-6)
    _eq(r["segmentation_loss"].item(), g[f"{tag}_seg"], exact, rtol=1e-6)
    dl = r["depth_loss"]
    _eq(dl.item() if torch.is_tensor(dl) else dl, g[f"{tag}_depthloss"], exact, rtol=1e-6)
    _eq(lg.grad.numpy(), g[f"{tag}_dlogits"], exact, rtol=1e-5, atol=1e-9)
    want = g[f"{tag}_ddepth"]
    if want.size:
        _eq(dp.grad.numpy(), want, exact, rtol=1e-5, atol=1e-9)
    else:
        assert dp.grad is None


def test_fog_density_from_depth(golden):
    g = golden("loss")
    d = torch.from_numpy(g["depth"]).squeeze(1)
    _eq(ol.fog_density_from_depth(d).numpy(), g["fd_from_depth"], _same_versions(g), rtol=1e-6, atol=1e-7)


# ------------------------------------------------- producers / consumers either side (8f rows 2-4)
from oracle import prep as op  # noqa: E402


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_fog_density_map_matches_reference(golden, tag):
    g = golden("prep")
    exact = _same_versions(g)
    img = g[f"{tag}_image"].astype(np.float32) / 255.0
    _eq(op.fog_density_map(img, g[f"{tag}_depth"]), g[f"{tag}_fogmap"], exact, atol=1e-6, rtol=1e-5)
    # depth=None path: the synthetic depth comes from the global RNG exactly as in the reference
    np.random.seed(14)
    depth = ow.depth_from_noise(np.random.normal(0, 10, img.shape[:2]))
    _eq(op.fog_density_map(img, depth), g[f"{tag}_fogmap_seed14"], exact, atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_estimate_depth_matches_reference(golden, tag):
    g = golden("prep")
    _eq(op.estimate_depth(g[f"{tag}_image"]), g[f"{tag}_est_depth"], _same_versions(g), atol=1e-12)


@pytest.mark.parametrize("kind", ["fog", "rain", "snow", "night", "clean"])
def test_style_transfer_matches_reference(golden, kind):
    g = golden("prep")
    for src in ("ramp", "rnd"):
        got = op.style_transfer(g[f"style_{src}"].copy(), kind)
        assert got.dtype == np.uint8
        assert np.array_equal(got, g[f"style_{src}_{kind}"])


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_domain_adaptation_augmentation_matches_reference(golden, seed):
    g = golden("prep")
    np.random.seed(seed)
    got = op.domain_adaptation_augmentation(g["aug_frame"].copy(), style_transfer_prob=0.6)
    d = np.abs(got.astype(int) - g[f"aug_seed{seed}"].astype(int))
    assert d.max() <= (0 if _same_versions(g) else 1)


@pytest.mark.parametrize("tag", ["t19", "t5"])
def test_temperature_grid_matches_reference(golden, tag):
    g = golden("prep")
    logits, targets = torch.from_numpy(g[f"{tag}_logits"]), torch.from_numpy(g[f"{tag}_targets"])
    _eq(op.temperature_nll(logits, targets).double().numpy(), g[f"{tag}_nll"], _same_versions(g), rtol=1e-6)
    assert op.optimize_temperature(logits, targets) == float(g[f"{tag}_best_t"])


def test_normalize_chw_restatement():
    """Parity unpinned (albumentations absent): the restatement against first principles in fp64."""
    rng = np.random.RandomState(3)
    img = rng.randint(0, 256, (20, 28, 3)).astype(np.uint8)
    got = op.normalize_chw(img)
    assert got.dtype == np.float32 and got.shape == (3, 20, 28)
    want = (img.astype(np.float64) / 255.0 - np.array(op.IMAGENET_MEAN)) / np.array(op.IMAGENET_STD)
    np.testing.assert_allclose(got, want.transpose(2, 0, 1), rtol=0, atol=2e-6)


@pytest.mark.parametrize("tag,n", [("m3", 3), ("m4", 4)])
def test_member_lists_match_reference(golden, tag, n):
    g = golden("prep")
    exact = _same_versions(g)
    members = [torch.from_numpy(g[f"{tag}_member{k}"]) for k in range(n)]
    targets = torch.from_numpy(g[f"{tag}_targets"])
    _eq(om.mi_map(members).numpy(), g[f"{tag}_mi"], exact, atol=2e-6, rtol=1e-5)
    _eq(om.variance_map(members).numpy(), g[f"{tag}_var"], exact, atol=1e-7, rtol=1e-5)
    assert om.disagreement_auroc(members, targets) == float(g[f"{tag}_auroc"])


# ------------------------------------------------- trainer twin (8f row 1): validation pass + fog-density maps
from trainer_fixture import ReplayModel, trainer_fixture_batches  # noqa: E402


def test_trainer_fog_density_maps_match_reference(golden):
    g = golden("trainer")
    torch.manual_seed(123)
    got = op.estimate_fog_density([str(x) for x in g["fd_names"]], 20, 28)
    np.testing.assert_array_equal(got.numpy(), g["fd_maps"])


@pytest.mark.parametrize("tag", ["v_depth", "v_nodepth", "v_focal", "v_plain"])
def test_validate_epoch_matches_reference(golden, tag):
    g = golden("trainer")
    seed, nb, bsz, c = (int(x) for x in g[f"{tag}_args"][:4])
    batches = trainer_fixture_batches(g, tag)
    if tag == "v_plain":
        res = op.validate_epoch(ReplayModel(batches), batches, torch.nn.CrossEntropyLoss(), c, fog_aware=False)
    else:
        base = str(g[f"{tag}_base"])
        torch.manual_seed(1000 + seed)
        res = op.validate_epoch(ReplayModel(batches), batches,
                                lambda o, t, fd: ol.fog_loss(o, t, fd, base_loss=base), c)
    assert sorted(res) == [str(k) for k in g[f"{tag}_keys"]]
    want = dict(zip((str(k) for k in g[f"{tag}_keys"]), g[f"{tag}_vals"]))
    for k, v in want.items():
        if "miou" in k or k == "val_samples":
            assert float(res[k]) == v, k
        else:
            assert float(res[k]) == pytest.approx(v, rel=1e-6), k
