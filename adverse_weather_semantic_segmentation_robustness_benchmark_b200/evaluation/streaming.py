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
        self._cfg_cache = {}

    def reset(self) -> None:
        self.bins.zero_()

    def update(self, condition: str, logits_a: torch.Tensor, logits_b: Optional[torch.Tensor],
               labels: torch.Tensor) -> None:
        """Score one batch of one condition: a single awx_score launch, nothing is materialised."""
        row = self.bins[self.conditions.index(condition)]
        ops.score(logits_a, logits_b if self.ensemble else None, labels, strategy=self.strategy,
                  w0=self.w0, w1=self.w1, temperature=self.temperature, ignore_index=self.ignore_index,
                  ece_bins=self.ece_bins, auroc_bins=self.auroc_bins, bins=row)

    def update_corrupted(self, condition: str, images: torch.Tensor, params, field, items, out: torch.Tensor,
                         workspace: torch.Tensor, logits_a: torch.Tensor, logits_b: Optional[torch.Tensor],
                         labels: torch.Tensor) -> None:
        """One awx_corrupt_score call: corrupt `images` (uint8 [B,H,W,3], device) into `out` for the model's
        next batch and score this batch's logits into the condition's bins.  Inputs must already be
        contiguous device tensors (labels uint8 or int64)."""
        row = self.bins[self.conditions.index(condition)]
        key = (logits_a.shape[1], labels.dtype)
        cfg = self._cfg_cache.get(key)
        if cfg is None:
            cfg = ops.score_config(logits_a.shape[1], self.strategy, self.w0, self.w1, self.temperature,
                                   ops.label_code(labels), self.ignore_index, self.ece_bins, self.auroc_bins)
            self._cfg_cache[key] = cfg
        ops.corrupt_score(images, params, field, items, out, workspace, logits_a,
                          logits_b if self.ensemble else None, labels, cfg, row)

    def canonical_bins(self) -> torch.Tensor:
        """The bins with every confidence sum in canonical form.  A sum is stored as hi * 2^32 + lo where both
        words are plain accumulators (each launch / CTA / rank adds its own split), so two buffers holding the
        same sums can differ word by word; here the carry of `lo` is folded into `hi`."""
        out = self.bins.clone()
        hi = slice(int(self.layout.ece_conf_hi), int(self.layout.ece_conf_hi) + self.ece_bins)
        lo = slice(int(self.layout.ece_conf_lo), int(self.layout.ece_conf_lo) + self.ece_bins)
        out[:, hi] += out[:, lo] >> 32
        out[:, lo] &= 0xFFFFFFFF
        return out

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
            # pixels whose integer outcome lies inside the reference's own fp32 rounding noise (include/awx.h)
            "ece_ambiguous_pixels": b.counter(_lib.CNT_ECE_AMBIG),
            "epred_ambiguous_pixels": b.counter(_lib.CNT_EPRED_AMBIG),
            "marg_ambiguous_pixels": b.counter(_lib.CNT_MARG_AMBIG),
            "pixels": b.counter(_lib.CNT_PIXELS),
        }
        if self.auroc_bins:
            out["ensemble_disagreement_auroc"], out["auroc_bound"] = finalize.auroc_from_histogram(
                b.auroc_pos, b.auroc_neg)
        return out

    def per_condition(self) -> Dict[str, Dict[str, float]]:
        host = self.bins.cpu().numpy()
        return {c: self._metrics_of(host[i]) for i, c in enumerate(self.conditions)
                if host[i].any() and not c.startswith("__")}

    def finalize(self) -> Dict[str, float]:
        """The reference's ``evaluate_model`` result dict (evaluate.py:214-272) from the merged bins."""
        host = self.bins.cpu().numpy()
        total = self._metrics_of(host.sum(axis=0))
        res: Dict[str, float] = {"overall_miou": total["mean_iou"]}
        mious = {}
        for i, c in enumerate(self.conditions):
            if not host[i].any() or c.startswith("__"):  # "__other__": frames counted in the overall numbers only
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


OTHER = "__other__"


def evaluate_model(model, test_loader, metrics=None, device=None, config=None, group=None) -> Dict[str, float]:
    """Drop-in for the reference's ``evaluate_model(model, test_loader, metrics, device, config)``
    (scripts/evaluate.py:134-274; trainer twin training/trainer.py:377-478), streaming and sharded.

    Same loop -- ``outputs = model(images)`` per batch, frames grouped by ``batch['weather_condition']`` --
    but nothing is concatenated: every run of consecutive frames with the same condition is scored by one
    ``awx_score`` launch into that condition's integer bins.  An ensemble model (outputs carry
    ``segformer_seg`` and ``deeplabv3plus_seg``) is fused inside the kernel with the model's own
    ``ensemble_strategy`` / ``ensemble_weights`` / ``temperature``; any other model is scored on
    ``outputs['segmentation']``.  Under ``torch.distributed`` every rank evaluates its shard of the loader
    and one ``all_reduce`` merges the bins.  ``metrics`` / ``device`` are accepted for signature
    compatibility (``metrics.num_classes`` and ``metrics.weather_conditions`` are honoured)."""
    get = (lambda k, d: config.get(k, d)) if config is not None and hasattr(config, "get") else (lambda k, d: d)
    conditions = list(get("data.weather_conditions", None) or getattr(metrics, "weather_conditions", None)
                      or DEFAULT_CONDITIONS)
    num_classes = int(getattr(metrics, "num_classes", None) or get("model.num_classes", 19))
    ev = None
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    with torch.no_grad():
        for batch in test_loader:
            images = batch["image"].to(device) if device is not None else batch["image"]
            labels = batch["label"].to(device) if device is not None else batch["label"]
            weather = list(batch.get("weather_condition", ["clean"] * images.size(0)))
            outputs = model(images)
            ensemble = "segformer_seg" in outputs and "deeplabv3plus_seg" in outputs
            if ev is None:
                if ensemble:
                    ts = bool(getattr(model, "temperature_scaling", False))
                    raw_w = getattr(model, "ensemble_weights", torch.ones(2) / 2).detach().float().cpu()
                    temp = float(model.temperature.detach().float().cpu()[0]) if ts else None
                    ev = StreamingEvaluator(num_classes, conditions + [OTHER], strategy=getattr(model, "ensemble_strategy", "mean"),
                                            ensemble_weights=raw_w.tolist(), temperature=temp, ensemble=True)
                else:
                    ev = StreamingEvaluator(num_classes, conditions + [OTHER], ensemble=False, temperature=None)
            la = outputs["segformer_seg"] if ensemble else outputs["segmentation"]
            lb = outputs["deeplabv3plus_seg"] if ensemble else None
            # runs of consecutive frames with the same condition are contiguous views: no copies
            i = 0
            while i < len(weather):
                j = i
                while j + 1 < len(weather) and weather[j + 1] == weather[i]:
                    j += 1
                cond = weather[i] if weather[i] in conditions else OTHER
                ev.update(cond, la[i:j + 1], None if lb is None else lb[i:j + 1], labels[i:j + 1])
                i = j + 1
    if was_training and hasattr(model, "train"):
        model.train()
    if ev is None:
        return {}
    ev.all_reduce(group)
    return ev.finalize()

