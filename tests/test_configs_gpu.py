"""The BASELINE.json configurations at their named sizes, against the oracle where it finishes in seconds and
through size-independent properties otherwise (configs[1] is bench.py's workload; its building blocks are
covered by test_score_gpu.py::test_full_size_* and test_corrupt_gpu.py::test_full_size_properties).

configs[0]  fog corruption + 19-class confusion matrix on 20 synthetic 512x1024 frames (the reference CPU path)
configs[2]  ensemble logit fusion 19x1024x2048: temperature softmax, disagreement map, ECE 15 bins, AUROC
configs[3]  fog-density-aware loss forward+backward with depth map, batch 8 at 1024x2048
"""

import numpy as np
import pytest
import torch

from oracle import weather as ow, metrics as om, fusion as of_, loss as ol
import parity

pytestmark = pytest.mark.gpu
C = 19


@pytest.fixture(scope="module")
def pkg():
    import adverse_weather_semantic_segmentation_robustness_benchmark_b200 as p
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib, ops
    _lib.load()
    return p, ops, _lib


def test_config0_fog_and_confusion_on_20_synthetic_frames(pkg):
    """Frames and labels exactly as the reference's synthetic dataset makes them (loader.py:206,231), fog through the
    drop-in transform with the reference's RNG order, confusion matrix / mIoU / accuracy of seeded predictions."""
    p, ops, _lib = pkg
    h, w, n = 512, 1024, 20
    np.random.seed(0)
    frames = [np.random.randint(0, 255, (h, w, 3), dtype=np.uint8) for _ in range(n)]
    labels = [np.random.randint(0, 19, (h, w), dtype=np.uint8) for _ in range(n)]
    t = p.WeatherDegradationTransforms(seed=42)
    got = [t.apply_weather_effect(f, "fog") for f in frames]
    np.random.seed(42)
    off = 0
    for i in range(n):
        want = ow.apply(frames[i], "fog")
        d = np.abs(got[i].astype(np.int16) - want.astype(np.int16))
        assert d.max() == 0, f"frame {i}: {int((d > 0).sum())} values differ"
        off += int((d > 0).sum())
    assert off == 0
    gen = torch.Generator().manual_seed(1)
    pred = torch.randint(0, C, (n, h, w), generator=gen)
    tgt = torch.from_numpy(np.stack(labels))
    iou = p.IoUMetrics(C)
    r = iou.compute_iou(pred, tgt)
    ref = om.iou(pred, tgt, C)
    assert r["mean_iou"] == ref["mean_iou"] and np.array_equal(r["per_class_iou"], ref["per_class_iou"])
    assert iou.compute_pixel_accuracy(pred, tgt) == om.pixel_accuracy(pred, tgt)
    cm, _ = ops.confusion(pred, tgt, C)
    assert np.array_equal(cm.cpu().numpy(), om.confusion_matrix(pred, tgt, C).numpy())


def test_config2_ensemble_fusion_full_frame(pkg):
    """One 19x1024x2048 frame per member: fused logits bit-exact, arg-max / confusion ==; ECE bins and the
    ensemble-wrong count == the oracle's up to the pixels the kernel reports as ambiguous, which must be fewer than
    2e-6 of the frame (tests/parity.py); MI map and AUROC within the stated tolerances."""
    p, ops, _lib = pkg
    gen = torch.Generator().manual_seed(42)
    la = torch.randn(1, C, 1024, 2048, generator=gen)
    gen.manual_seed(43)
    lb = torch.randn(1, C, 1024, 2048, generator=gen)
    gen.manual_seed(44)
    tgt = torch.randint(0, C, (1, 1024, 2048), generator=gen)
    tgt[0, :16] = 255
    raw_w, temp = torch.tensor([0.3, 0.9]), torch.tensor([1.7])
    w = of_.member_weights(raw_w)
    want = of_.fuse_logits(la, lb, "weighted_average", raw_w, temp)
    out = ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=float(w[0]), w1=float(w[1]), temperature=1.7,
                    auroc_bins=4096, want_pred=torch.int64, want_fused=True, want_mi=True, want_conf=True)
    assert torch.equal(out["fused"].cpu(), want)
    assert torch.equal(out["pred"].cpu(), want.argmax(1))
    bins = ops.read_bins(out["bins"], C, 15, 4096)
    assert np.array_equal(bins.confusion, om.confusion_matrix(want, tgt, C).numpy())
    ref = om.ece(want, tgt)
    valid = tgt != 255
    n_valid = int(valid.sum())
    cap = parity.genuine_ece_near_edges(want, tgt)
    assert cap <= 1e-5 * n_valid
    parity.assert_ece_parity(bins, ref, n_valid, _lib, cap=cap)
    assert bins.counter(_lib.CNT_CORRECT) == int(((want.argmax(1) == tgt) & valid).sum())
    wrong = int(((om.mean_prob_prediction([la, lb]) != tgt) & valid).sum())
    parity.assert_ens_wrong_parity(bins, wrong, n_valid, _lib)
    conf_ref, _ = om.confidence_and_prediction(want)
    assert (out["conf"].cpu() - conf_ref).abs().max() <= 2.4e-7
    mi_ref = om.mi_map([la, lb])
    assert ((out["mi"].cpu() - mi_ref).abs() - (1e-5 * mi_ref.abs() + 2e-6)).max() <= 0
    auroc, bound = p.EnsembleDisagreementMetrics().compute_disagreement_auroc([la, lb], tgt, return_bound=True)
    exact = om.disagreement_auroc([la, lb], tgt)
    assert abs(auroc - exact) <= bound + 1e-6 and bound < 2e-3
    # the streaming (bins-only) kernel on the same frame
    fast = ops.read_bins(ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=float(w[0]), w1=float(w[1]),
                                   temperature=1.7, auroc_bins=4096)["bins"], C, 15, 4096)
    assert np.array_equal(fast.confusion, bins.confusion)
    parity.assert_ece_parity(fast, ref, n_valid, _lib, cap=cap)
    parity.assert_ens_wrong_parity(fast, wrong, n_valid, _lib)
    assert fast.counter(_lib.CNT_CORRECT) == bins.counter(_lib.CNT_CORRECT)


def test_config3_loss_forward_backward_batch8(pkg):
    """B = 8 at 1024x2048 (16.8 Mpx): values against the oracle on a 2-frame slice rescaled, and identities that
    hold at any size: every pixel's logit gradient sums to zero over classes, the depth gradient is
    2 (pred - target) / N, gradients scale linearly with the upstream gradient, the run is deterministic."""
    p, ops, _lib = pkg
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_loss
    dev = torch.device("cuda")
    gen = torch.Generator(device=dev).manual_seed(9)
    b, h, w = 8, 1024, 2048
    logits = torch.randn(b, C, h, w, device=dev, generator=gen)
    label = torch.randint(0, C, (b, h, w), device=dev, generator=gen)
    fd = torch.rand(b, h, w, device=dev, generator=gen)
    dpred = torch.rand(b, 1, h, w, device=dev, generator=gen) * 50
    dtgt = torch.rand(b, h, w, device=dev, generator=gen) * 50
    n = float(b * h * w)
    sums, dlogits, ddepth, bad, _ = ops_loss.fogloss_raw(logits, label, fd, dpred, dtgt, 2.0, False, True)
    sums2, dlogits2, _, _, _ = ops_loss.fogloss_raw(logits, label, fd, dpred, dtgt, 2.0, False, True)
    assert int(bad.item()) == 0
    assert torch.equal(sums, sums2) and torch.equal(dlogits, dlogits2)
    assert float(dlogits.sum(dim=1).abs().max()) <= 1e-12 + 8e-7 / n * 100
    assert torch.allclose(ddepth, 2.0 * (dpred[:, 0] - dtgt) / n, rtol=1e-6, atol=0)
    # oracle on frames 0..1: the per-frame sums add up, so compare the partial sums of that slice
    sl = slice(0, 2)
    s_part, dl_part, _, _, _ = ops_loss.fogloss_raw(logits[sl], label[sl], fd[sl], dpred[sl], dtgt[sl], 2.0, False, True)
    lg = logits[sl].cpu().requires_grad_(True)
    dp = dpred[sl].cpu().requires_grad_(True)
    r = ol.fog_loss({"segmentation": lg, "depth": dp}, {"label": label[sl].cpu(), "depth": dtgt[sl].cpu()}, fd[sl].cpu())
    r["total_loss"].backward()
    n2 = float(2 * h * w)
    np.testing.assert_allclose(float(s_part[0]) / n2, float(r["segmentation_loss"].detach()), rtol=1e-5)
    np.testing.assert_allclose(float(s_part[1]) / n2, float(r["depth_loss"].detach()), rtol=1e-5)
    ex = (dl_part.cpu() - lg.grad).abs() - (1e-5 * lg.grad.abs() + 1e-9)
    assert float(ex.max()) <= 0
    # the full batch is the sum of its slices (fp64 sums: equal to ~1e-12 relative)
    acc = torch.zeros(2, dtype=torch.float64, device=dev)
    for i in range(0, b, 2):
        acc += ops_loss.fogloss_raw(logits[i:i + 2], label[i:i + 2], fd[i:i + 2], dpred[i:i + 2], dtgt[i:i + 2], 2.0, False, False)[0]
    assert torch.allclose(acc, sums, rtol=1e-12, atol=0)
    # public class, autograd scaling
    lgd = logits.requires_grad_(True)
    fn = p.FogDensityAwareLoss()
    out = fn({"segmentation": lgd, "depth": dpred}, {"label": label, "depth": dtgt}, fd)
    (3.0 * out["total_loss"]).backward()
    assert torch.allclose(lgd.grad, 3.0 * dlogits, rtol=1e-6, atol=0)
    np.testing.assert_allclose(float(out["segmentation_loss"].detach()), float(sums[0]) / n, rtol=1e-6)
