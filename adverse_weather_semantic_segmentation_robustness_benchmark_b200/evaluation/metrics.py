"""Drop-in metric classes: same names, signatures, return types and dict keys as the reference's
``evaluation/metrics.py``, with the per-pixel work done by libawx.so on the GPU.

Every method accepts host or device tensors (host tensors are uploaded), launches one fused
pass (``awx_score`` / ``awx_confusion``), reads back a few hundred integers and finalises them
on the host with the reference's own fp32 expressions.  Reference lines are cited per method.
"""

from __future__ import annotations

from typing import Any, Dict, List, Optional, Union

import numpy as np
import torch

from .. import _lib, ops
from . import finalize

__all__ = ["IoUMetrics", "ConfidenceCalibration", "EnsembleDisagreementMetrics", "RobustnessMetrics"]


def _as_logits(t: torch.Tensor) -> torch.Tensor:
    return ops.to_device(t, torch.float32)


class IoUMetrics:
    """Confusion-matrix IoU and pixel accuracy (reference: evaluation/metrics.py:15-123)."""

    def __init__(self, num_classes: int, ignore_index: int = 255) -> None:
        self.num_classes = num_classes
        self.ignore_index = ignore_index

    def _confusion(self, predictions: torch.Tensor, targets: torch.Tensor):
        """int64 [C,C] confusion (host) and counters, from logits [B,C,H,W] or a prediction map."""
        if predictions.dim() == 4 and predictions.shape[1] == self.num_classes:
            out = ops.score(_as_logits(predictions), None, targets, ignore_index=self.ignore_index, ece_bins=1)
            bins = ops.read_bins(out["bins"], self.num_classes, 1, 0)
            return (bins.confusion.copy(), bins.counter(_lib.CNT_VALID), bins.counter(_lib.CNT_CORRECT),
                    bins.counter(_lib.CNT_BAD_LABEL))
        if predictions.dim() == 4:
            # channel count differs from num_classes: the reference still indexes a
            # num_classes x num_classes matrix with t*num_classes + argmax -> go through the map
            predictions = ops.score(_as_logits(predictions), want_pred=torch.int64)["pred"]
        # `targets * C + predictions` must promote to int64 or the reference's index_add_ raises
        if torch.promote_types(targets.dtype, predictions.dtype) != torch.int64:
            raise RuntimeError("index_add_(): self (Long) and source "
                               f"({torch.promote_types(targets.dtype, predictions.dtype)}) must have the same scalar type")
        cm_d, cnt_d = ops.confusion(predictions, targets, self.num_classes, self.ignore_index)
        cnt = cnt_d.cpu().numpy()
        return cm_d.cpu().numpy(), int(cnt[_lib.CNT_VALID]), int(cnt[_lib.CNT_CORRECT]), int(cnt[_lib.CNT_BAD_LABEL])

    def compute_iou(self, predictions: torch.Tensor, targets: torch.Tensor) -> Dict[str, float]:
        """metrics.py:34-89.  Keys: mean_iou (float), per_class_iou (np.float32[C]), valid_classes (bool[C])."""
        cm, _, _, bad = self._confusion(predictions, targets)
        if bad:
            # the reference's index_add_ raises on these (labels outside [0,C) that are not ignore_index)
            raise IndexError(f"index out of range in self ({bad} pixels have a target/prediction pair outside the "
                             f"{self.num_classes}x{self.num_classes} confusion matrix)")
        return finalize.iou_from_confusion(cm)

    def compute_pixel_accuracy(self, predictions: torch.Tensor, targets: torch.Tensor) -> float:
        """metrics.py:91-123: correct / valid as a Python float, 0.0 when nothing is valid."""
        _, valid, correct, _ = self._confusion(predictions, targets)
        return correct / valid if valid > 0 else 0.0


class ConfidenceCalibration:
    """ECE / reliability data / temperature scaling (reference: evaluation/metrics.py:126-321)."""

    def __init__(self, num_bins: int = 15) -> None:
        self.num_bins = num_bins

    def compute_ece(self, predictions: torch.Tensor, targets: torch.Tensor,
                    return_details: bool = False) -> Union[float, Dict[str, Any]]:
        """metrics.py:143-226.  One pass: softmax max-prob, (lo,hi] bin, per-bin count / correct / conf sum."""
        logits = _as_logits(predictions)
        ncls = logits.shape[1]
        out = ops.score(logits, None, targets, ece_bins=self.num_bins)
        bins = ops.read_bins(out["bins"], ncls, self.num_bins, 0)
        res = finalize.ece_from_bins(bins.ece_count, bins.ece_correct, bins.ece_conf_sum,
                                     bins.counter(_lib.CNT_VALID), ops.ece_edges(self.num_bins).numpy())
        res["ambiguous_pixels"] = bins.counter(_lib.CNT_ECE_AMBIG)
        if return_details:
            return res
        return res["ece"]

    def compute_reliability_diagram_data(self, predictions: torch.Tensor, targets: torch.Tensor) -> Dict[str, np.ndarray]:
        """metrics.py:228-264."""
        details = self.compute_ece(predictions, targets, return_details=True)["bin_details"]
        keep = [d for d in details if d["proportion"] > 0]
        return {
            "bin_centers": np.array([(d["bin_lower"] + d["bin_upper"]) / 2 for d in keep]),
            "bin_accuracies": np.array([d["accuracy"] for d in keep]),
            "bin_confidences": np.array([d["confidence"] for d in keep]),
            "bin_proportions": np.array([d["proportion"] for d in keep]),
        }

    def temperature_scale(self, logits: torch.Tensor, temperature) -> torch.Tensor:
        """metrics.py:266-281: ``logits / temperature`` (true fp32 division).  Like the reference's eager division
        it is differentiable w.r.t. the logits and a tensor temperature: with grad mode on and either requiring
        grad, the quotient goes through the fusion autograd function as the mean of the logits with themselves,
        (x + x) / 2 / T = x / T bit for bit, whose backward kernel returns g / T and the sum for dL/dT."""
        wants_grad = torch.is_grad_enabled() and (
            (torch.is_tensor(logits) and logits.requires_grad)
            or (torch.is_tensor(temperature) and temperature.requires_grad))
        if wants_grad:
            from ..models.model import _FuseFn
            x = ops.to_device(logits, torch.float32)
            t = temperature if torch.is_tensor(temperature) else torch.tensor([float(temperature)])
            flat = x if x.dim() == 4 else x.reshape(1, 1, -1, 1)
            return _FuseFn.apply(flat, flat, None, t.reshape(-1)[:1], "mean").reshape(x.shape)
        temperature = float(temperature.detach().reshape(-1)[0]) if torch.is_tensor(temperature) else float(temperature)
        x = _as_logits(logits)
        squeeze = x.dim() != 4
        if squeeze:
            # [N, C] or other layouts: present as [1, C, N, 1] with classes on dim 1 is not the same
            # memory order; elementwise division does not care about the layout, so flatten.
            flat = x.reshape(1, 1, -1, 1)
            return ops.score(flat, temperature=temperature, want_fused=True)["fused"].reshape(x.shape)
        return ops.score(x, temperature=temperature, want_fused=True)["fused"]

    def optimize_temperature(self, logits: torch.Tensor, targets: torch.Tensor, max_iter: int = 50) -> float:
        """metrics.py:283-321: 100-point grid search of the NLL over T in linspace(0.1, 10)."""
        from .. import ops_loss
        return ops_loss.temperature_grid_search(logits, targets)


class EnsembleDisagreementMetrics:
    """Disagreement maps and AUROC (reference: evaluation/metrics.py:324-467)."""

    def __init__(self) -> None:
        pass

    @staticmethod
    def _two_members(predictions_list: List[torch.Tensor]):
        if len(predictions_list) < 2:
            raise ValueError("Need at least 2 predictions for disagreement computation")
        return _as_logits(predictions_list[0]), _as_logits(predictions_list[1])

    def compute_disagreement_map(self, predictions_list: List[torch.Tensor]) -> torch.Tensor:
        """metrics.py:336-369: H(mean p) - mean_k H(p_k) with the reference's log(p + 1e-8); [B,H,W]."""
        a, b = self._two_members(predictions_list)
        if len(predictions_list) > 2:  # the general list form (awx_members_n); two members ride along in awx_score
            return ops.members_n(predictions_list, want_mi=True)["mi"]
        return ops.score(a, b, strategy=_lib.FUSE_MEAN, want_mi=True)["mi"]

    def compute_variance_map(self, predictions_list: List[torch.Tensor]) -> torch.Tensor:
        """metrics.py:371-391: unbiased variance over members of the class probabilities; [B,C,H,W]."""
        a, b = self._two_members(predictions_list)
        if len(predictions_list) > 2:
            return ops.members_n(predictions_list, want_var=True)["var"]
        return ops.member_variance(a, b)

    def compute_disagreement_auroc(self, predictions_list: List[torch.Tensor], targets: torch.Tensor,
                                   error_threshold: float = 0.5, num_bins: int = ops.DEFAULT_AUROC_BINS,
                                   return_bound: bool = False):
        """metrics.py:393-438.  Streaming form: MI scores are histogrammed (``num_bins`` linear bins over
        [0, ln 2)) separately for wrong / right ensemble pixels and the AUROC is the exact
        Mann-Whitney statistic of the binned score; its distance from sklearn's value on the
        unbinned score is at most ``bound`` (returned with ``return_bound=True``)."""
        a, b = self._two_members(predictions_list)
        if len(predictions_list) > 2:
            out = ops.members_n(predictions_list, targets, auroc_bins=num_bins)
            value, bound = finalize.auroc_from_histogram(out["pos"].cpu().numpy(), out["neg"].cpu().numpy())
            return (value, bound) if return_bound else value
        out = ops.score(a, b, targets, strategy=_lib.FUSE_MEAN, auroc_bins=num_bins)
        bins = ops.read_bins(out["bins"], a.shape[1], 15, num_bins)
        value, bound = finalize.auroc_from_histogram(bins.auroc_pos, bins.auroc_neg)
        return (value, bound) if return_bound else value

    def compute_jensen_shannon_divergence(self, pred1: torch.Tensor, pred2: torch.Tensor) -> torch.Tensor:
        """metrics.py:440-467: 0.5*[KL(m||p) + KL(m||q)] exactly as F.kl_div(log p, m) defines it."""
        a, b = _as_logits(pred1), _as_logits(pred2)
        return ops.score(a, b, strategy=_lib.FUSE_MEAN, want_js=True)["js"]


class RobustnessMetrics:
    """Orchestration of the three metric families (reference: evaluation/metrics.py:470-651)."""

    def __init__(self, num_classes: int = 19, weather_conditions: List[str] = None) -> None:
        self.num_classes = num_classes
        self.weather_conditions = weather_conditions or ["clean", "fog", "rain", "snow", "night"]
        self.iou_metrics = IoUMetrics(num_classes)
        self.calibration_metrics = ConfidenceCalibration()
        self.ensemble_metrics = EnsembleDisagreementMetrics()

    def compute_miou(self, predictions: torch.Tensor, targets: torch.Tensor) -> float:
        """metrics.py:498-514."""
        return self.iou_metrics.compute_iou(predictions, targets)["mean_iou"]

    def compute_weather_specific_metrics(self, predictions_dict: Dict[str, torch.Tensor],
                                         targets_dict: Dict[str, torch.Tensor]) -> Dict[str, float]:
        """metrics.py:516-542: ``miou_<weather>`` for every configured condition present in both dicts."""
        metrics = {}
        for weather in self.weather_conditions:
            if weather in predictions_dict and weather in targets_dict:
                preds, tgts = predictions_dict[weather], targets_dict[weather]
                if len(preds) > 0 and len(tgts) > 0:
                    metrics[f"miou_{weather}"] = self.compute_miou(preds, tgts)
        return metrics

    def compute_robustness_degradation_ratio(self, clean_miou: float, adverse_miou: float) -> float:
        """metrics.py:544-563."""
        return finalize.degradation_ratio(clean_miou, adverse_miou)

    def compute_comprehensive_metrics(self, predictions: torch.Tensor, targets: torch.Tensor,
                                      ensemble_predictions: Optional[List[torch.Tensor]] = None,
                                      weather_condition: str = "clean") -> Dict[str, float]:
        """metrics.py:565-605.  The reference reads the logits six times and softmaxes them five
        times; here `predictions` is read once (confusion + accuracy + ECE from one launch) and the
        two members once (MI histogram)."""
        metrics: Dict[str, float] = {}
        if predictions.dim() == 4 and predictions.shape[1] == self.num_classes:
            out = ops.score(_as_logits(predictions), None, targets, ece_bins=self.calibration_metrics.num_bins)
            bins = ops.read_bins(out["bins"], self.num_classes, self.calibration_metrics.num_bins, 0)
            if bins.counter(_lib.CNT_BAD_LABEL):
                raise IndexError("index out of range in self")
            metrics["mean_iou"] = finalize.iou_from_confusion(bins.confusion)["mean_iou"]
            valid = bins.counter(_lib.CNT_VALID)
            metrics["pixel_accuracy"] = bins.counter(_lib.CNT_CORRECT) / valid if valid > 0 else 0.0
            metrics["expected_calibration_error"] = finalize.ece_from_bins(
                bins.ece_count, bins.ece_correct, bins.ece_conf_sum, valid,
                ops.ece_edges(self.calibration_metrics.num_bins).numpy())["ece"]
        else:
            metrics["mean_iou"] = self.iou_metrics.compute_iou(predictions, targets)["mean_iou"]
            metrics["pixel_accuracy"] = self.iou_metrics.compute_pixel_accuracy(predictions, targets)
            metrics["expected_calibration_error"] = self.calibration_metrics.compute_ece(predictions, targets)
        if ensemble_predictions and len(ensemble_predictions) >= 2:
            metrics["ensemble_disagreement_auroc"] = self.ensemble_metrics.compute_disagreement_auroc(
                ensemble_predictions, targets)
        metrics[f"miou_{weather_condition}"] = metrics["mean_iou"]
        return metrics

    def create_robustness_summary(self, weather_metrics: Dict[str, Dict[str, float]]) -> Dict[str, float]:
        """metrics.py:607-651."""
        summary: Dict[str, float] = {}
        clean_miou = weather_metrics.get("clean", {}).get("mean_iou", 0.0)
        adverse = ["fog", "rain", "snow", "night"]
        for weather in adverse:
            if weather in weather_metrics:
                summary[f"robustness_degradation_{weather}"] = self.compute_robustness_degradation_ratio(
                    clean_miou, weather_metrics[weather].get("mean_iou", 0.0))
        degradations = [summary[f"robustness_degradation_{w}"] for w in adverse
                        if f"robustness_degradation_{w}" in summary]
        if degradations:
            summary["robustness_degradation_ratio"] = np.mean(degradations)
        eces = [m.get("expected_calibration_error", 0.0) for m in weather_metrics.values()]
        if eces:
            summary["expected_calibration_error"] = np.mean(eces)
        aurocs = [m.get("ensemble_disagreement_auroc", 0.5) for m in weather_metrics.values()]
        if aurocs:
            summary["ensemble_disagreement_auroc"] = np.mean(aurocs)
        return summary
