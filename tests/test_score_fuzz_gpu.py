"""Seeded fuzz of awx_score against the oracle: random shapes (vector and scalar paths, tail tiles), label
dtypes, ignore fractions, logit scales (near-saturated softmaxes), strategies, raw weights and temperatures.
For every case: arg-max map and confusion matrix ==, correct / valid counters ==, ECE counts == up to the
reported ambiguous pixels, MI map within 1e-5*|ref| + 2e-6, and the bins-only (streaming) kernel agrees
with the kernel that emits the maps."""

import numpy as np
import pytest
import torch

from oracle import metrics as om, fusion as of_
import parity

pytestmark = pytest.mark.gpu
C = 19


def _case(seed):
    rng = np.random.RandomState(seed)
    b = int(rng.randint(1, 4))
    h = int(rng.choice([7, 16, 33, 48, 64, 120]))
    w = int(rng.choice([9, 20, 35, 64, 96, 244]))
    if seed % 3 == 0:            # force the TMA kernel: H*W a multiple of 4
        w = (w + 3) // 4 * 4
    ldt = torch.uint8 if rng.rand() < 0.5 else torch.int64
    ignore = float(rng.choice([0.0, 0.05, 0.5]))
    scale = float(rng.choice([0.2, 1.0, 3.0, 12.0]))
    strategy = str(rng.choice(["weighted_average", "mean", "max_confidence"]))
    temp = rng.choice([None, 0.37, 1.0, 2.9])
    temp = None if temp is None else float(temp)
    raw_w = torch.tensor(rng.uniform(-1.5, 1.5, 2).astype(np.float32))
    gen = torch.Generator().manual_seed(seed)
    la = torch.randn(b, C, h, w, generator=gen) * scale
    lb = torch.randn(b, C, h, w, generator=gen) * scale + 0.5 * la
    tgt = torch.randint(0, C, (b, h, w), generator=gen).to(ldt)
    if ignore > 0:
        tgt[torch.rand(b, h, w, generator=gen) < ignore] = 255
    return la, lb, tgt, strategy, temp, raw_w


@pytest.mark.parametrize("seed", range(36))
def test_score_fuzz(seed):
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib
    la, lb, tgt, strategy, temp, raw_w = _case(seed)
    code = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF, "mean": _lib.FUSE_MEAN}[strategy]
    w = of_.member_weights(raw_w)
    want = of_.fuse_logits(la, lb, strategy, raw_w, None if temp is None else torch.tensor([temp]))
    kw = dict(strategy=code, w0=float(w[0]), w1=float(w[1]), temperature=temp, auroc_bins=4096)
    out = ops.score(la, lb, tgt, want_pred=torch.int64, want_mi=True, want_conf=True, **kw)
    bins = ops.read_bins(out["bins"], C, 15, 4096)
    pick_amb = bins.counter(_lib.CNT_PICK_AMBIG)
    valid = tgt != 255
    if pick_amb == 0:
        assert torch.equal(out["pred"].cpu(), want.argmax(1))
        assert np.array_equal(bins.confusion, om.confusion_matrix(want, tgt, C).numpy())
        assert bins.counter(_lib.CNT_CORRECT) == int(((want.argmax(1) == tgt) & valid).sum())
        ref = om.ece(want, tgt)
        parity.assert_ece_parity(bins, ref, int(valid.sum()), _lib, floor=2)
        conf_ref, _ = om.confidence_and_prediction(want)
        # 3 ulp at 1.0, plus the reference's own rounding of z = v/T before its softmax (~ulp(|z|max) relative):
        # the kernel keeps the exact quotient in the exponent, so it is the more accurate of the two
        assert (out["conf"].cpu() - conf_ref).abs().max() <= 3.6e-7 + 3e-8 * float(want.abs().max())
    assert bins.counter(_lib.CNT_VALID) == int(valid.sum()) and bins.counter(_lib.CNT_PIXELS) == tgt.numel()
    assert bins.counter(_lib.CNT_BAD_LABEL) == 0
    mi_ref = om.mi_map([la, lb])
    err = (out["mi"].cpu() - mi_ref).abs() - (1e-5 * mi_ref.abs() + 2e-6)
    assert float(err.max()) <= 0, f"MI excess {float(err.max()):.2e}"
    wrong = (om.mean_prob_prediction([la, lb]) != tgt) & valid
    # saturated softmaxes (logit scale 12) tie for real: the cap on the reported ties is their fp64 count
    parity.assert_ens_wrong_parity(bins, int(wrong.sum()), int(valid.sum()), _lib, cap=parity.genuine_marg_ties([la, lb], tgt))
    # bins only
    fast = ops.read_bins(ops.score(la, lb, tgt, **kw)["bins"], C, 15, 4096)
    assert np.array_equal(fast.confusion, bins.confusion)
    for k in (_lib.CNT_VALID, _lib.CNT_CORRECT, _lib.CNT_ENS_WRONG, _lib.CNT_NO_BIN):
        assert fast.counter(k) == bins.counter(k), k
    amb2 = bins.counter(_lib.CNT_ECE_AMBIG) + fast.counter(_lib.CNT_ECE_AMBIG)
    assert np.abs(fast.ece_count - bins.ece_count).sum() <= 2 * amb2
    assert np.abs(fast.auroc_pos - bins.auroc_pos).sum() + np.abs(fast.auroc_neg - bins.auroc_neg).sum() <= 2
