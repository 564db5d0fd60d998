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
        # the buffer is always laid out WITH the AUROC histograms (they come last, include/awx.h): single-member
        # launches touch only the prefix, so every rank of a sharded evaluation holds the same number of words
        # whatever its shard contains, and one all_reduce fits all
        self.auroc_bins = auroc_bins
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
        self._scratch = None
        self._cfg_cache = {}

    def reset(self) -> None:
        self.bins.zero_()

    def update(self, condition: str, logits_a: torch.Tensor, logits_b: Optional[torch.Tensor],
               labels: torch.Tensor) -> None:
        """Score one batch of one condition: a single awx_score launch, nothing is materialised."""
        row = self.bins[self.conditions.index(condition)]
        ens = self.ensemble and logits_b is not None
        ops.score(logits_a, logits_b if ens else None, labels, strategy=self.strategy if ens else _lib.FUSE_SINGLE,
                  w0=self.w0, w1=self.w1, temperature=self.temperature if ens else None,
                  ignore_index=self.ignore_index, ece_bins=self.ece_bins, auroc_bins=self.auroc_bins if ens else 0,
                  bins=row)

    def update_members(self, condition: str, fused: torch.Tensor, logits_a: torch.Tensor, logits_b: torch.Tensor,
                       labels: torch.Tensor) -> None:
        """The general form of evaluate.py:178-201 for a model whose fusion this package cannot restate: the
        confusion matrix and the ECE bins come from the model's OWN ``outputs['segmentation']`` (one single-member
        launch), the disagreement histogram from the two members (a second launch into a scratch buffer, of which
        only the AUROC words and their counters are kept)."""
        idx = self.conditions.index(condition)
        row = self.bins[idx]
        ops.score(fused, None, labels, strategy=_lib.FUSE_SINGLE, temperature=None, ignore_index=self.ignore_index,
                  ece_bins=self.ece_bins, auroc_bins=0, bins=row)
        if self._scratch is None:
            self._scratch = torch.zeros(self.words, dtype=torch.int64, device=self.bins.device)
        self._scratch.zero_()
        ops.score(logits_a, logits_b, labels, strategy=_lib.FUSE_MEAN, temperature=None, ignore_index=self.ignore_index,
                  ece_bins=self.ece_bins, auroc_bins=self.auroc_bins, bins=self._scratch)
        lay = self.layout
        hist = slice(int(lay.auroc_pos), int(lay.auroc_pos) + 2 * self.auroc_bins)   # pos and neg are adjacent
        row[hist] += self._scratch[hist]
        for k in (_lib.CNT_ENS_WRONG, _lib.CNT_MARG_AMBIG):
            row[int(lay.counters) + k] += self._scratch[int(lay.counters) + k]

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
        if self.auroc_bins and (self.ensemble or int(b.auroc_pos.sum() + b.auroc_neg.sum()) > 0):
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
        if "ensemble_disagreement_auroc" in total:
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


def _unwrap(model):
    """The module inside DistributedDataParallel / DataParallel style wrappers (``.module``)."""
    seen = 0
    while isinstance(getattr(model, "module", None), torch.nn.Module) and seen < 4:
        model = model.module
        seen += 1
    return model


def _fusion_of(core):
    """(strategy, raw ensemble weights, temperature | None) of an EnsembleModel-like module, or None when it does
    not carry the reference's fusion attributes (models/model.py:385-426) -- nothing is ever defaulted."""
    if not all(hasattr(core, a) for a in ("ensemble_strategy", "ensemble_weights", "temperature_scaling")):
        return None
    ts = bool(core.temperature_scaling)
    if ts and not hasattr(core, "temperature"):
        return None
    raw_w = core.ensemble_weights.detach().float().cpu().reshape(-1)
    if raw_w.numel() != 2:
        return None
    temp = float(core.temperature.detach().float().cpu().reshape(-1)[0]) if ts else None
    return str(core.ensemble_strategy), raw_w.tolist(), temp


def evaluate_model(model, test_loader, metrics=None, device=None, config=None, group=None) -> Dict[str, float]:
    """Drop-in for the reference's ``evaluate_model(model, test_loader, metrics, device, config)``
    (scripts/evaluate.py:134-274; trainer twin training/trainer.py:377-478), streaming and sharded.

    Same loop -- ``outputs = model(images)`` per batch, frames grouped by ``batch['weather_condition']`` --
    but nothing is concatenated: every run of consecutive frames with the same condition is scored on the device
    into that condition's integer bins.  What is scored is what the reference scores: ``outputs['segmentation']``
    for mIoU / ECE and, for an ensemble (``hasattr(model, 'segformer')`` and member logits in the outputs,
    evaluate.py:196), the two members for the disagreement AUROC.  When the (unwrapped) model carries the
    reference's fusion attributes AND re-fusing the first batch's members in the kernel reproduces its
    ``outputs['segmentation']`` bit for bit, the fusion is done inside the scoring kernel (one launch, 152 B/px
    instead of 229); otherwise the model's own fused logits are scored and the members only feed the histogram.
    Under ``torch.distributed`` every rank evaluates its shard of the loader -- an empty shard included -- and one
    ``all_reduce`` merges the bins.  ``metrics.num_classes`` / ``metrics.weather_conditions`` are honoured."""
    from ..models.model import _STRATEGY, _fuse_forward
    get = (lambda k, d: config.get(k, d)) if config is not None and hasattr(config, "get") else (lambda k, d: d)
    conditions = list(get("data.weather_conditions", None) or getattr(metrics, "weather_conditions", None)
                      or DEFAULT_CONDITIONS)
    num_classes = int(getattr(metrics, "num_classes", None) or get("model.num_classes", 19))
    core = _unwrap(model)
    fusion = _fusion_of(core)
    dev = torch.device(device) if device is not None else ops.require_cuda()
    if dev.type != "cuda":
        raise RuntimeError("evaluate_model scores on a CUDA device (libawx.so has no CPU path); got device=%s" % dev)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    strategy, raw_w, temp = fusion if fusion is not None else ("mean", [0.5, 0.5], None)
    # built BEFORE the loop: a rank whose shard is empty still owns (zero) bins of the common size for the all_reduce
    ev = StreamingEvaluator(num_classes, conditions + [OTHER], strategy=strategy, ensemble_weights=raw_w,
                            temperature=temp, ensemble=True, bins_device=dev)
    ev.ensemble_seen = False
    in_kernel = None     # decided on the first ensemble batch
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    with torch.no_grad():
        for batch in test_loader:
            images = batch["image"].to(dev)
            labels = batch["label"].to(dev)
            weather = list(batch.get("weather_condition", ["clean"] * images.size(0)))
            outputs = model(images)
            seg = outputs.get("segmentation")
            # evaluate.py:196: ensemble statistics only for a model that HAS a `segformer` member
            members = hasattr(model, "segformer") and "segformer_seg" in outputs and "deeplabv3plus_seg" in outputs
            la = lb = None
            if members:
                la, lb = outputs["segformer_seg"].to(dev), outputs["deeplabv3plus_seg"].to(dev)
                ev.ensemble_seen = True
                if in_kernel is None:
                    in_kernel = False
                    if fusion is not None:
                        w = torch.softmax(torch.tensor(raw_w, dtype=torch.float32), dim=0)
                        code = _STRATEGY.get(strategy, _lib.FUSE_MEAN)
                        in_kernel = seg is None or bool(torch.equal(
                            _fuse_forward(la, lb, code, float(w[0]), float(w[1]), temp), seg.to(dev).float()))
            if seg is None and not (members and in_kernel):
                raise KeyError("model outputs carry no 'segmentation' logits and the model has no fusion attributes "
                               "(ensemble_strategy / ensemble_weights / temperature) to rebuild them from")
            seg = None if seg is None else seg.to(dev)
            # runs of consecutive frames with the same condition are contiguous views: no copies
            i = 0
            while i < len(weather):
                j = i
                while j + 1 < len(weather) and weather[j + 1] == weather[i]:
                    j += 1
                cond = weather[i] if weather[i] in conditions else OTHER
                sl = slice(i, j + 1)
                if members and in_kernel:
                    ev.update(cond, la[sl], lb[sl], labels[sl])
                elif members:
                    ev.update_members(cond, seg[sl], la[sl], lb[sl], labels[sl])
                else:
                    ev.update(cond, seg[sl], None, labels[sl])
                i = j + 1
    if was_training and hasattr(model, "train"):
        model.train()
    ev.all_reduce(group)
    ev.ensemble = False    # the AUROC key appears iff some rank scored member logits (its histogram is non-empty)
    host = ev.bins.cpu()
    if not bool(host.any()):
        return {}          # nothing evaluated anywhere (the reference's torch.cat of an empty list raises here)
    return ev.finalize()
