"""Oracle: fog-density-aware pixel loss (forward; backward through torch autograd).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows ``P/models/model.py:560-677``
(``P/`` = /root/reference/src/adverse_weather_semantic_segmentation_robustness_benchmark/):
per-pixel cross entropy (or focal, :619-642), times ``1 + s * fog_density``
(:584-587), optional fog density estimated from predicted depth (:593-597,
:644-677), optional depth MSE (:600-604), means and weighted total (:610-611).
Labels are NOT filtered for 255 -- the reference lets torch raise on them.
"""

from __future__ import annotations

import torch
import torch.nn.functional as F


def focal_from_ce(ce: torch.Tensor, alpha: float = 1.0, gamma: float = 2.0) -> torch.Tensor:
    """alpha * (1 - exp(-ce))**gamma * ce (model.py:638-642)."""
    return alpha * (1 - torch.exp(-ce)) ** gamma * ce


def fog_density_from_depth(depth: torch.Tensor) -> torch.Tensor:
    """Heuristic density in [0,1] from a [B,H,W] depth map (model.py:657-677)."""
    unit = (depth - depth.min()) / (depth.max() - depth.min() + 1e-8)
    density = unit * 0.7
    gx = torch.abs(depth[:, :, 1:] - depth[:, :, :-1])
    gy = torch.abs(depth[:, 1:, :] - depth[:, :-1, :])
    gx = F.pad(gx, (0, 1, 0, 0), mode="replicate")
    gy = F.pad(gy, (0, 0, 0, 1), mode="replicate")
    mag = torch.sqrt(gx ** 2 + gy ** 2 + 1e-8)
    density = density - (mag > mag.mean()) * 0.3
    return torch.clamp(density, 0, 1)


def fog_loss(predictions: dict, targets: dict, fog_density=None, *,
             base_loss: str = "cross_entropy", depth_weight: float = 0.5,
             fog_sensitivity: float = 2.0, depth_loss_weight: float = 0.1) -> dict:
    """Returns {'total_loss','segmentation_loss','depth_loss'} (model.py:577-617)."""
    logits = predictions["segmentation"]
    labels = targets["label"].long()
    ce = F.cross_entropy(logits, labels, reduction="none")
    per_pixel = focal_from_ce(ce) if base_loss == "focal" else ce
    if fog_density is not None:
        per_pixel = per_pixel * (1.0 + fog_sensitivity * fog_density)
    depth_term = 0.0
    if "depth" in predictions and depth_weight > 0:
        d = predictions["depth"].squeeze(1)
        if fog_density is None:
            per_pixel = per_pixel * (1.0 + fog_sensitivity * fog_density_from_depth(d))
        if "depth" in targets:
            depth_term = F.mse_loss(d, targets["depth"], reduction="none").mean()
    seg = per_pixel.mean()
    return {"total_loss": seg + depth_loss_weight * depth_term,
            "segmentation_loss": seg, "depth_loss": depth_term}
