"""Host-side finalisers: turn the integer bins the kernels accumulate into the numbers the
reference reports.  A few hundred scalars; runs on the CPU with the reference's own fp32
expressions so the last digits agree (evaluation/metrics.py:74-89, 183-224, 430-438).
"""

from __future__ import annotations

import numpy as np
import torch


def iou_from_confusion(cm: np.ndarray) -> dict:
    """rows = target, cols = prediction.  int64/int64 -> fp32 per class, fp32 mean over classes
    whose union is non-empty (metrics.py:74-89)."""
    cm_t = torch.from_numpy(np.ascontiguousarray(cm, dtype=np.int64))
    inter = torch.diag(cm_t)
    union = cm_t.sum(dim=0) + cm_t.sum(dim=1) - inter
    valid = union > 0
    per_class = torch.zeros(cm_t.shape[0])
    per_class[valid] = inter[valid] / union[valid]
    return {
        "mean_iou": per_class[valid].mean().item(),
        "per_class_iou": per_class.numpy(),
        "valid_classes": valid.numpy(),
    }


def ece_from_bins(count: np.ndarray, correct: np.ndarray, conf_sum: np.ndarray, n_valid: int,
                  edges: np.ndarray) -> dict:
    """ECE and per-bin details from integer bins (metrics.py:183-224).  Per-bin means are formed
    in float64 and rounded once to fp32, then combined in fp32 as the reference does."""
    f32 = np.float32
    ece = f32(0.0)
    details = []
    any_bin = False
    for b in range(len(count)):
        lo, hi = float(edges[b]), float(edges[b + 1])
        n = int(count[b])
        prop = f32(n / n_valid) if n_valid > 0 else f32(np.nan)
        if n > 0 and prop > 0:
            acc = f32(int(correct[b]) / n)
            conf = f32(float(conf_sum[b]) / n)
            err = f32(abs(conf - acc))
            ece = f32(ece + f32(err * prop))
            any_bin = True
            details.append({"bin_lower": lo, "bin_upper": hi, "accuracy": float(acc), "confidence": float(conf),
                            "proportion": float(prop), "error": float(err)})
        else:
            details.append({"bin_lower": lo, "bin_upper": hi, "accuracy": 0.0, "confidence": 0.0,
                            "proportion": 0.0, "error": 0.0})
    total_correct = int(np.sum(correct))
    total_conf = float(np.sum(conf_sum))
    return {
        "ece": float(ece) if any_bin else 0.0,
        "bin_details": details,
        "overall_accuracy": float(f32(total_correct / n_valid)) if n_valid > 0 else float("nan"),
        "overall_confidence": float(f32(total_conf / n_valid)) if n_valid > 0 else float("nan"),
    }


def auroc_from_histogram(pos: np.ndarray, neg: np.ndarray):
    """AUROC of a binned score (ties inside a bin count one half) and the bound on its distance
    from the AUROC of the unbinned score, 0.5 * sum_b pos_b*neg_b / (P*N).  Exact integer
    arithmetic (Python ints), so the value is independent of how the bins were sharded."""
    pos = [int(x) for x in pos]
    neg = [int(x) for x in neg]
    p, n = sum(pos), sum(neg)
    if p == 0 or n == 0:
        return 0.5, 0.0  # fewer than two classes present: the reference returns 0.5 (metrics.py:430-431)
    below = 0
    twice_u = 0
    ties = 0
    for pb, nb_ in zip(pos, neg):
        twice_u += 2 * pb * below + pb * nb_
        ties += pb * nb_
        below += nb_
    return twice_u / (2 * p * n), ties / (2 * p * n)


def degradation_ratio(clean_miou: float, adverse_miou: float) -> float:
    """metrics.py:559-563."""
    if clean_miou == 0:
        return 1.0
    return max(0.0, (clean_miou - adverse_miou) / clean_miou)
