"""Streaming form of the trainer's validation pass (training/trainer.py:377-478) and of the per-batch
fog-density maps it feeds the loss (training/trainer.py:480-511).

The reference keeps every frame's argmax map and labels on the CPU and builds the confusion matrices at the end
of the epoch; here each batch is scored once on the device into per-weather integer bins (``awx_score`` on the
fused ``outputs['segmentation']``), the loss terms come from one ``awx_fogloss`` launch per batch, and the only
device->host traffic of the epoch is one stack of loss scalars and the packed bins.  Result keys are the
reference's: ``val_loss, val_seg_loss, val_depth_loss, val_samples, val_miou, val_miou_<weather>``.
"""

from __future__ import annotations

from typing import Any, Dict, Optional, Sequence

import torch

from .. import ops
from .streaming import DEFAULT_CONDITIONS, OTHER, StreamingEvaluator

# trainer.py:497-509: torch.rand(h, w) * scale (+ offset) per frame; anything that is not fog / rain / snow
# gets the "clean" map rand * 0.1 (no offset: the reference does not add one)
FOG_DENSITY_AFFINE = {"fog": (0.5, 0.5), "rain": (0.3, 0.2), "snow": (0.3, 0.2)}
CLEAN_SCALE = 0.1


def estimate_fog_density(batch: Dict[str, Any], device: Optional[torch.device] = None) -> Optional[torch.Tensor]:
    """``AdverseWeatherTrainer._estimate_fog_density`` (trainer.py:480-511).

    The uniform draws stay on the host, from the global torch CPU generator and in the reference's order (one
    ``torch.rand(h, w)`` per frame), so a seeded run sees the reference's maps bit for bit; they are drawn
    straight into one pinned [B,h,w] buffer and cross to the device once.  The per-frame affine map is applied
    there as a multiply followed by an add (two roundings, as ``rand * a + b`` has on the CPU)."""
    weather_conditions = batch.get("weather_condition", [])
    if not len(weather_conditions):
        return None
    n = len(weather_conditions)
    h, w = batch["image"].shape[2:]
    dev = ops.require_cuda() if device is None else torch.device(device)
    try:
        host = torch.empty((n, h, w), dtype=torch.float32, pin_memory=True)
    except RuntimeError:
        host = torch.empty((n, h, w), dtype=torch.float32)
    scale = torch.empty(n, dtype=torch.float32)
    offset = torch.zeros(n, dtype=torch.float32)
    for i, weather in enumerate(weather_conditions):
        torch.rand(h, w, out=host[i])
        a, b = FOG_DENSITY_AFFINE.get(weather, (CLEAN_SCALE, 0.0))
        scale[i], offset[i] = a, b
    out = host.to(dev, non_blocking=True)
    out.mul_(scale.to(dev).view(n, 1, 1))
    out.add_(offset.to(dev).view(n, 1, 1))   # + 0.0 leaves the non-negative clean maps unchanged
    return out


def validate_epoch(model, val_loader, loss_fn, metrics=None, device=None, group=None,
                   weather_conditions: Sequence[str] = DEFAULT_CONDITIONS) -> Dict[str, float]:
    """Drop-in for ``AdverseWeatherTrainer.validate_epoch`` (trainer.py:377-478) as a function of the trainer's
    members (``self.model, self.val_loader, self.loss_fn, self.metrics, self.device``).

    ``loss_fn`` is this package's ``FogDensityAwareLoss`` (then the fog-density maps of ``estimate_fog_density``
    are drawn per batch, as the trainer does) or any callable ``loss_fn(logits, labels) -> scalar tensor``.
    Under ``torch.distributed`` every rank validates its shard of the loader; the bins and the loss sums are
    merged by one ``all_reduce`` each."""
    from ..models.model import FogDensityAwareLoss

    num_classes = int(getattr(metrics, "num_classes", None) or 19)
    conditions = list(weather_conditions)
    dev = ops.require_cuda() if device is None else torch.device(device)
    ev = StreamingEvaluator(num_classes, conditions + [OTHER], auroc_bins=0, ensemble=False, temperature=None,
                            bins_device=dev)
    fog_aware = isinstance(loss_fn, FogDensityAwareLoss)
    was_training = getattr(model, "training", False)
    if hasattr(model, "eval"):
        model.eval()
    terms, sizes = [], []       # per batch: [total, seg, depth] device scalars (read back once) and the batch size
    with torch.no_grad():
        for batch in val_loader:
            images = batch["image"].to(dev)
            labels = batch["label"].to(dev)
            weather = list(batch.get("weather_condition", ["clean"] * images.size(0)))
            depths = batch.get("depth")
            outputs = model(images)
            targets = {"label": labels}
            if depths is not None:
                targets["depth"] = depths.to(dev)
            if fog_aware:
                fog_density = estimate_fog_density(batch, dev)
                ld = loss_fn(outputs, targets, fog_density)
                dl = ld["depth_loss"]
                if not isinstance(dl, torch.Tensor):
                    dl = torch.full((), float(dl), dtype=torch.float32, device=dev)
                terms.append(torch.stack([ld["total_loss"].float().reshape(()), ld["segmentation_loss"].float().reshape(()),
                                          dl.float().reshape(())]))
            else:
                seg = loss_fn(outputs["segmentation"], labels).float().reshape(())
                terms.append(torch.stack([seg, seg, torch.zeros((), dtype=torch.float32, device=dev)]))
            sizes.append(images.size(0))
            logits = outputs["segmentation"]
            i = 0
            while i < len(weather):   # runs of consecutive frames with the same weather: contiguous views, one launch
                j = i
                while j + 1 < len(weather) and weather[j + 1] == weather[i]:
                    j += 1
                ev.update(weather[i] if weather[i] in conditions else OTHER, logits[i:j + 1], None, labels[i:j + 1])
                i = j + 1
    if was_training and hasattr(model, "train"):
        model.train()
    # the reference adds loss.item() * batch_size batch by batch into Python floats (trainer.py:436-444)
    sums = [0.0, 0.0, 0.0]
    if terms:
        host = torch.stack(terms).cpu()
        for row, bs in zip(host.tolist(), sizes):
            for k in range(3):
                sums[k] += row[k] * bs
    samples = sum(sizes)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        packed = torch.tensor(sums + [float(samples)], dtype=torch.float64, device=ev.bins.device)
        dist.all_reduce(packed, op=dist.ReduceOp.SUM, group=group)
        packed = packed.tolist()
        sums, samples = packed[:3], int(round(packed[3]))
    ev.all_reduce(group)
    res: Dict[str, float] = {"val_loss": sums[0], "val_seg_loss": sums[1], "val_depth_loss": sums[2],
                             "val_samples": samples}
    for key in ("val_loss", "val_seg_loss", "val_depth_loss"):
        res[key] /= res["val_samples"]   # ZeroDivisionError on an empty loader, as in the reference
    host_bins = ev.bins.cpu().numpy()
    res["val_miou"] = ev._metrics_of(host_bins.sum(axis=0))["mean_iou"]
    for i, c in enumerate(conditions):
        if host_bins[i].any():
            res[f"val_miou_{c}"] = ev._metrics_of(host_bins[i])["mean_iou"]
    return res
