"""Streaming, sharded form of the reference's evaluation driver (scripts/evaluate.py:134-274).

The reference concatenates every frame's full-resolution logits on the CPU and scores them at the
end; memory grows with frames x 19 x H x W x 4 B x 3 tensors.  Here each batch is scored once on
the device into per-condition integer bins (a few KB), ranks shard the frames, and ONE
``all_reduce(SUM)`` of the packed int64 buffer merges them.  Integer bins make the merged result
independent of the number of ranks and of the reduction order (bit-reproducible at 1/2/4/8 GPUs).
The result dict keeps the reference's keys: ``overall_miou``, ``miou_<w>``, ``ece_<w>``,
``expected_calibration_error``, ``ensemble_disagreement_auroc``, ``robustness_degradation_<w>``,
``robustness_degradation_ratio``.
"""

from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import _lib, ops
from . import finalize

DEFAULT_CONDITIONS = ("clean", "fog", "rain", "snow", "night")


class StreamingEvaluator:
    def __init__(self, num_classes: int = 19, conditions: Sequence[str] = DEFAULT_CONDITIONS,
                 ece_bins: int = 15, auroc_bins: int = ops.DEFAULT_AUROC_BINS,
                 strategy: str = "weighted_average", ensemble_weights: Sequence[float] = (0.5, 0.5),
                 temperature: Optional[float] = 1.0, ensemble: bool = True, ignore_index: int = 255,
                 bins_device: Optional[torch.device] = None) -> None:
        self.num_classes = num_classes
        self.conditions = tuple(conditions)
        self.ece_bins = ece_bins
        self.ensemble = ensemble
        self.auroc_bins = auroc_bins if ensemble else 0
        self.ignore_index = ignore_index
        self.temperature = temperature
        codes = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF}
        self.strategy = codes.get(strategy, _lib.FUSE_MEAN) if ensemble else _lib.FUSE_SINGLE
        w = torch.softmax(torch.tensor(list(ensemble_weights), dtype=torch.float32), dim=0)
        self.w0, self.w1 = float(w[0]), float(w[1])
        self.layout = _lib.bins_layout(num_classes, ece_bins, self.auroc_bins)
        self.words = int(self.layout.total_words)
        # one packed buffer [n_conditions, words]: a single collective merges everything
        # (bins_device="cpu" exists for the host-side merge/finalise logic and its gloo tests; update()
        # always needs the CUDA device)
        dev = ops.require_cuda() if bins_device is None else torch.device(bins_device)
        self.bins = torch.zeros((len(self.conditions), self.words), dtype=torch.int64, device=dev)

    def reset(self) -> None:
        self.bins.zero_()

    def update(self, condition: str, logits_a: torch.Tensor, logits_b: Optional[torch.Tensor],
               labels: torch.Tensor) -> None:
        """Score one batch of one condition: a single awx_score launch, nothing is materialised."""
        row = self.bins[self.conditions.index(condition)]
        ops.score(logits_a, logits_b if self.ensemble else None, labels, strategy=self.strategy,
                  w0=self.w0, w1=self.w1, temperature=self.temperature, ignore_index=self.ignore_index,
                  ece_bins=self.ece_bins, auroc_bins=self.auroc_bins, bins=row)

    def all_reduce(self, group=None) -> None:
        """Merge the bins of all ranks (NCCL over NVLink on GPUs; any backend that sums int64)."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.bins, op=dist.ReduceOp.SUM, group=group)

    # ------------------------------------------------------------------------------ finalise
    def _view(self, raw: np.ndarray) -> ops.Bins:
        return ops.Bins(raw, self.num_classes, self.ece_bins, self.auroc_bins)

    def _metrics_of(self, raw: np.ndarray) -> Dict[str, float]:
        b = self._view(raw)
        if b.counter(_lib.CNT_BAD_LABEL):
            raise IndexError("index out of range in self (labels outside the confusion matrix)")
        valid = b.counter(_lib.CNT_VALID)
        out = {
            "mean_iou": finalize.iou_from_confusion(b.confusion)["mean_iou"] if valid else float("nan"),
            "pixel_accuracy": b.counter(_lib.CNT_CORRECT) / valid if valid else 0.0,
            "expected_calibration_error": finalize.ece_from_bins(
                b.ece_count, b.ece_correct, b.ece_conf_sum, valid, ops.ece_edges(self.ece_bins).numpy())["ece"],
            "ece_ambiguous_pixels": b.counter(_lib.CNT_ECE_AMBIG),
            "pixels": b.counter(_lib.CNT_PIXELS),
        }
        if self.auroc_bins:
            out["ensemble_disagreement_auroc"], out["auroc_bound"] = finalize.auroc_from_histogram(
                b.auroc_pos, b.auroc_neg)
        return out

    def per_condition(self) -> Dict[str, Dict[str, float]]:
        host = self.bins.cpu().numpy()
        return {c: self._metrics_of(host[i]) for i, c in enumerate(self.conditions) if host[i].any()}

    def finalize(self) -> Dict[str, float]:
        """The reference's ``evaluate_model`` result dict (evaluate.py:214-272) from the merged bins."""
        host = self.bins.cpu().numpy()
        total = self._metrics_of(host.sum(axis=0))
        res: Dict[str, float] = {"overall_miou": total["mean_iou"]}
        mious = {}
        for i, c in enumerate(self.conditions):
            if not host[i].any():
                continue
            m = self._metrics_of(host[i])
            mious[c] = m["mean_iou"]
            res[f"miou_{c}"] = m["mean_iou"]
            res[f"ece_{c}"] = m["expected_calibration_error"]
        res["expected_calibration_error"] = total["expected_calibration_error"]
        if self.auroc_bins:
            res["ensemble_disagreement_auroc"] = total["ensemble_disagreement_auroc"]
        if "clean" in mious:
            degr = []
            for c in ("fog", "rain", "snow", "night"):
                if c in mious:
                    res[f"robustness_degradation_{c}"] = finalize.degradation_ratio(mious["clean"], mious[c])
                    degr.append(res[f"robustness_degradation_{c}"])
            if degr:
                res["robustness_degradation_ratio"] = np.mean(degr)
        return res
