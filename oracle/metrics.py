"""Oracle: confusion-matrix IoU, pixel accuracy, ECE bins, disagreement maps, AUROC.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``P/evaluation/metrics.py`` (``P/`` = /root/reference/src/
adverse_weather_semantic_segmentation_robustness_benchmark/):

* confusion / IoU          :34-89      * pixel accuracy          :91-123
* ECE + bin details        :143-226    * reliability-diagram data :228-264
* MI disagreement map      :336-369    * variance map            :371-391
* disagreement AUROC       :393-438    * robustness bookkeeping  :544-563, :607-651

The second half (``*_bin_index`` / ``auroc_from_histogram``) restates the
*streaming* form the CUDA path uses -- bins instead of concatenated logits --
so tests can check "counts are exact given the emitted map" and
"histogram AUROC is within its stated bound of sklearn's exact value".
"""

from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from sklearn.metrics import roc_auc_score

IGNORE = 255


# ------------------------------------------------------------------ confusion / IoU
def confusion_matrix(pred: torch.Tensor, target: torch.Tensor, num_classes: int,
                     ignore_index: int = IGNORE) -> torch.Tensor:
    """int64 [C,C], rows = target, cols = prediction (metrics.py:50-71).

    ``target * C + pred`` is evaluated with torch type promotion: a uint8
    label tensor times a Python int stays uint8 and WRAPS mod 256 before the
    int64 prediction is added.  Kept as is.
    """
    if pred.dim() == 4:
        pred = pred.argmax(dim=1)
    pred = pred.reshape(-1)
    target = target.reshape(-1)
    keep = target != ignore_index
    pred = pred[keep]
    target = target[keep]
    flat = torch.zeros(num_classes * num_classes, dtype=torch.long)
    idx = target * num_classes + pred
    flat.index_add_(0, idx.long(), torch.ones_like(idx))
    return flat.view(num_classes, num_classes)


def iou_from_confusion(cm: torch.Tensor) -> dict:
    """Per-class IoU (fp32 from int64/int64) and mean over classes with union>0 (:74-89)."""
    inter = torch.diag(cm)
    union = cm.sum(dim=0) + cm.sum(dim=1) - inter
    valid = union > 0
    per_class = torch.zeros(cm.shape[0])
    per_class[valid] = inter[valid] / union[valid]
    return {
        "mean_iou": per_class[valid].mean().item(),
        "per_class_iou": per_class.numpy(),
        "valid_classes": valid.numpy(),
    }


def iou(pred: torch.Tensor, target: torch.Tensor, num_classes: int,
        ignore_index: int = IGNORE) -> dict:
    return iou_from_confusion(confusion_matrix(pred, target, num_classes, ignore_index))


def pixel_accuracy(pred: torch.Tensor, target: torch.Tensor, ignore_index: int = IGNORE) -> float:
    """correct / valid as a Python float; 0.0 when nothing is valid (:106-123)."""
    if pred.dim() == 4:
        pred = pred.argmax(dim=1)
    pred = pred.reshape(-1)
    target = target.reshape(-1)
    keep = target != ignore_index
    pred = pred[keep]
    target = target[keep]
    n = target.numel()
    return (pred == target).sum().item() / n if n > 0 else 0.0


# ------------------------------------------------------------------------------ ECE
def ece_edges(num_bins: int = 15) -> torch.Tensor:
    """fp32 bin boundaries exactly as torch.linspace produces them (:179)."""
    return torch.linspace(0, 1, num_bins + 1)


def confidence_and_prediction(logits: torch.Tensor):
    """max softmax probability and ITS argmax (over probabilities, not logits) (:161-162)."""
    return torch.max(F.softmax(logits, dim=1), dim=1)


def ece(logits: torch.Tensor, target: torch.Tensor, num_bins: int = 15) -> dict:
    """ECE with (lo, hi] bins (:164-224).  Always returns the detailed dict, plus
    the integer per-bin ``count`` / ``correct`` the streaming form must reproduce."""
    conf, pred = confidence_and_prediction(logits)
    conf = conf.reshape(-1)
    pred = pred.reshape(-1)
    target = target.reshape(-1)
    drop = target == IGNORE
    conf = conf[~drop]
    pred = pred[~drop]
    target = target[~drop]
    acc = (pred == target).float()
    edges = ece_edges(num_bins)
    total = 0.0
    details, counts, corrects = [], [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (conf > lo) & (conf <= hi)
        share = sel.float().mean()
        counts.append(int(sel.sum().item()))
        corrects.append(int((sel & (acc > 0)).sum().item()))
        if share > 0:
            a = acc[sel].mean()
            c = conf[sel].mean()
            gap = torch.abs(c - a)
            total = total + gap * share
            details.append({"bin_lower": lo.item(), "bin_upper": hi.item(),
                            "accuracy": a.item(), "confidence": c.item(),
                            "proportion": share.item(), "error": gap.item()})
        else:
            details.append({"bin_lower": lo.item(), "bin_upper": hi.item(),
                            "accuracy": 0.0, "confidence": 0.0,
                            "proportion": 0.0, "error": 0.0})
    return {
        "ece": total.item() if torch.is_tensor(total) else float(total),
        "bin_details": details,
        "overall_accuracy": acc.mean().item(),
        "overall_confidence": conf.mean().item(),
        "count": np.asarray(counts, dtype=np.int64),
        "correct": np.asarray(corrects, dtype=np.int64),
    }


def reliability_points(details: list) -> dict:
    """Non-empty bins as plotting arrays (:246-264)."""
    keep = [d for d in details if d["proportion"] > 0]
    return {
        "bin_centers": np.array([(d["bin_lower"] + d["bin_upper"]) / 2 for d in keep]),
        "bin_accuracies": np.array([d["accuracy"] for d in keep]),
        "bin_confidences": np.array([d["confidence"] for d in keep]),
        "bin_proportions": np.array([d["proportion"] for d in keep]),
    }


# ------------------------------------------------------------- disagreement and AUROC
def mi_map(members: list) -> torch.Tensor:
    """Mutual-information disagreement H(mean p) - mean_k H(p_k), eps inside the logs (:349-369)."""
    if len(members) < 2:
        raise ValueError("Need at least 2 predictions for disagreement computation")
    probs = torch.stack([F.softmax(m, dim=1) for m in members], dim=0)
    mean_p = probs.mean(dim=0)
    h_mean = -torch.sum(mean_p * torch.log(mean_p + 1e-8), dim=1)
    h_each = -torch.sum(probs * torch.log(probs + 1e-8), dim=2)
    return h_mean - h_each.mean(dim=0)


def variance_map(members: list) -> torch.Tensor:
    """Unbiased variance over members of the class probabilities, [B,C,H,W] (:384-391)."""
    probs = torch.stack([F.softmax(m, dim=1) for m in members], dim=0)
    return torch.var(probs, dim=0)


def mean_prob_prediction(members: list) -> torch.Tensor:
    """argmax of the member-averaged probabilities (:414-416)."""
    return torch.stack([F.softmax(m, dim=1) for m in members], dim=0).mean(dim=0).argmax(dim=1)


def disagreement_auroc(members: list, target: torch.Tensor) -> float:
    """Exact AUROC (sklearn) of MI disagreement vs ensemble error (:410-438)."""
    score = mi_map(members).reshape(-1).numpy()
    wrong = (mean_prob_prediction(members) != target).float().reshape(-1).numpy()
    keep = (target.reshape(-1) != IGNORE).numpy()
    score = score[keep]
    wrong = wrong[keep]
    if len(np.unique(wrong)) < 2:
        return 0.5
    try:
        return roc_auc_score(wrong, score)
    except ValueError:
        return 0.5


# -------------------------------------------------------------- robustness bookkeeping
def degradation_ratio(clean_miou: float, adverse_miou: float) -> float:
    """max(0, (clean-adv)/clean); 1.0 when clean is 0 (:559-563)."""
    if clean_miou == 0:
        return 1.0
    return max(0.0, (clean_miou - adverse_miou) / clean_miou)


# ------------------------------------------------------- streaming (binned) restatement
def ece_bin_index(conf: np.ndarray, edges: np.ndarray) -> np.ndarray:
    """Bin b such that edges[b] < conf <= edges[b+1]; -1 when in no bin (conf<=0, NaN, >1)."""
    conf = np.asarray(conf, dtype=np.float32)
    edges = np.asarray(edges, dtype=np.float32)
    out = np.full(conf.shape, -1, dtype=np.int64)
    for b in range(len(edges) - 1):
        out[(conf > edges[b]) & (conf <= edges[b + 1])] = b
    return out


def mi_bin_index(mi: np.ndarray, num_bins: int, hi: float) -> np.ndarray:
    """Linear bins over [0, hi): floor(fp32(mi * fp32(num_bins/hi))), clamped to [0, num_bins-1].
    NaN goes to bin 0.  Mirrors ``awx_score``'s AUROC histogram rule (include/awx.h)."""
    mi = np.asarray(mi, dtype=np.float32)
    scale = np.float32(np.float32(num_bins) / np.float32(hi))
    with np.errstate(invalid="ignore"):
        q = np.floor(mi * scale)
    q = np.where(np.isnan(q), 0, q)
    return np.clip(q, 0, num_bins - 1).astype(np.int64)


def auroc_from_histogram(pos: np.ndarray, neg: np.ndarray):
    """AUROC of the binned score (ties inside a bin count 1/2) and the bound on its
    distance from the unbinned AUROC: 0.5*sum_b pos_b*neg_b / (P*N)."""
    pos = np.asarray(pos, dtype=np.float64)
    neg = np.asarray(neg, dtype=np.float64)
    p, n = pos.sum(), neg.sum()
    if p == 0 or n == 0:
        return 0.5, 0.0
    neg_below = np.cumsum(neg) - neg
    u = np.sum(pos * neg_below) + 0.5 * np.sum(pos * neg)
    return float(u / (p * n)), float(0.5 * np.sum(pos * neg) / (p * n))
