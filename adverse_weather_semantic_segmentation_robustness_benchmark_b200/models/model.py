"""Drop-in ``EnsembleModel`` (fusion part) and ``FogDensityAwareLoss`` (reference: models/model.py:377-677).

The SegFormer / DeepLabV3+ backbones stay ordinary PyTorch producers of logits (north-star); this
module owns what happens to their outputs: logit fusion, temperature scaling, the disagreement
map and the fog-density-aware loss -- all computed by libawx.so.
"""

from __future__ import annotations

import logging
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib, ops, ops_loss

logger = logging.getLogger(__name__)

_STRATEGY = {"weighted_average": _lib.FUSE_WEIGHTED, "max_confidence": _lib.FUSE_MAXCONF}


def _reference_backbones(num_classes: int, include_depth: bool):
    """The stock members, if the reference package is importable; otherwise the caller injects its own."""
    try:
        from adverse_weather_semantic_segmentation_robustness_benchmark.models.model import (  # type: ignore
            SegFormerModel, DeepLabV3PlusModel)
    except Exception as exc:  # pragma: no cover - depends on the environment
        raise RuntimeError(
            "EnsembleModel needs two logits producers: pass segformer=/deeplabv3plus= modules returning "
            "{'segmentation': [B,C,H,W][, 'depth': [B,1,H,W]]}, or install the reference package for its "
            f"stock backbones ({exc})") from exc
    return (SegFormerModel(num_classes=num_classes, include_depth=include_depth),
            DeepLabV3PlusModel(num_classes=num_classes, include_depth=include_depth))


class EnsembleModel(nn.Module):
    """Two-member ensemble with learnable fusion weights and temperature (models/model.py:377-513).

    Same constructor arguments, parameters (``ensemble_weights[2]``, ``temperature[1]``), output
    keys and strategies as the reference; ``segformer`` / ``deeplabv3plus`` may be injected."""

    def __init__(self, num_classes: int = 19, include_depth: bool = True,
                 ensemble_strategy: str = "weighted_average", temperature_scaling: bool = True,
                 segformer: Optional[nn.Module] = None, deeplabv3plus: Optional[nn.Module] = None) -> None:
        super().__init__()
        self.num_classes = num_classes
        self.include_depth = include_depth
        self.ensemble_strategy = ensemble_strategy
        self.temperature_scaling = temperature_scaling
        if segformer is None or deeplabv3plus is None:
            segformer, deeplabv3plus = _reference_backbones(num_classes, include_depth)
        self.segformer = segformer
        self.deeplabv3plus = deeplabv3plus
        self.ensemble_weights = nn.Parameter(torch.ones(2) / 2)
        if self.temperature_scaling:
            self.temperature = nn.Parameter(torch.ones(1))
        logger.info(f"Initialized ensemble model with {ensemble_strategy} strategy (libawx fusion)")

    # -- fusion -------------------------------------------------------------------------------
    def _fusion_scalars(self):
        """Host copies of softmax(ensemble_weights) (fp32, model.py:444) and the temperature."""
        w = F.softmax(self.ensemble_weights.detach().float().cpu(), dim=0)
        temp = float(self.temperature.detach().float().cpu()[0]) if self.temperature_scaling else None
        return float(w[0]), float(w[1]), temp

    def _wants_graph(self, a: torch.Tensor, b: torch.Tensor, with_weights: bool, with_temperature: bool) -> bool:
        """The reference's eager expression always carries autograd to the members, ``ensemble_weights`` and
        ``temperature`` (models/model.py:443-462), in train() and eval() mode alike: take the differentiable path
        whenever grad mode is on and ANY of them requires grad (frozen backbones with a trainable temperature --
        post-hoc calibration in eval mode -- included)."""
        if not torch.is_grad_enabled():
            return False
        params = []
        if with_weights:
            params.append(self.ensemble_weights)
        if with_temperature and self.temperature_scaling:
            params.append(self.temperature)
        return a.requires_grad or b.requires_grad or any(p.requires_grad for p in params)

    def fuse(self, seg1: torch.Tensor, seg2: torch.Tensor) -> torch.Tensor:
        """Fused, temperature-scaled logits (model.py:443-462), bit-exact w.r.t. torch eager."""
        if self._wants_graph(seg1, seg2, self.ensemble_strategy == "weighted_average", True):
            return _FuseFn.apply(seg1, seg2, self.ensemble_weights,
                                 self.temperature if self.temperature_scaling else None,
                                 self.ensemble_strategy)
        w0, w1, temp = self._fusion_scalars()
        code = _STRATEGY.get(self.ensemble_strategy, _lib.FUSE_MEAN)
        return _fuse_forward(seg1, seg2, code, w0, w1, temp)

    def fuse_depth(self, d1: torch.Tensor, d2: torch.Tensor) -> torch.Tensor:
        """model.py:471-478: weighted for weighted_average, plain mean otherwise; no temperature."""
        strategy = "weighted_average" if self.ensemble_strategy == "weighted_average" else "mean"
        if self._wants_graph(d1, d2, strategy == "weighted_average", False):
            return _FuseFn.apply(d1, d2, self.ensemble_weights, None, strategy)
        w0, w1, _ = self._fusion_scalars()
        code = _STRATEGY.get(strategy, _lib.FUSE_MEAN)
        return ops.fuse_forward(d1, d2, code, w0, w1, None)

    def forward(self, x: torch.Tensor) -> Dict[str, torch.Tensor]:
        o1 = self.segformer(x)
        o2 = self.deeplabv3plus(x)
        results = {
            "segmentation": self.fuse(o1["segmentation"], o2["segmentation"]),
            "segformer_seg": o1["segmentation"],
            "deeplabv3plus_seg": o2["segmentation"],
        }
        if self.include_depth:
            results.update({
                "depth": self.fuse_depth(o1["depth"], o2["depth"]),
                "segformer_depth": o1["depth"],
                "deeplabv3plus_depth": o2["depth"],
            })
        return results

    def get_ensemble_disagreement(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:488-513: 0.5*[KL(m||p)+KL(m||q)] of the two members' softmaxes, [B,H,W], no grad."""
        with torch.no_grad():
            o1 = self.segformer(x)
            o2 = self.deeplabv3plus(x)
            return ops.score(o1["segmentation"], o2["segmentation"], strategy=_lib.FUSE_MEAN, want_js=True)["js"]


def _fuse_forward(a, b, code, w0, w1, temp):
    """The fused logits alone.  weighted_average / mean are element-wise (awx_fuse_forward, one streaming pass);
    max_confidence needs both member softmaxes per pixel, which the TMA-staged score kernel already computes
    faster than a plain per-pixel kernel does (its fused map, statistics discarded)."""
    if code == _lib.FUSE_MAXCONF:
        return ops.score(a, b, strategy=code, w0=w0, w1=w1, temperature=temp, want_fused=True)["fused"]
    return ops.fuse_forward(a, b, code, w0, w1, temp)


class _FuseFn(torch.autograd.Function):
    """Differentiable fusion for training: forward through awx_fuse_forward, backward through awx_fuse_backward (the two
    scaled copies of the incoming gradient and the three dot products in one pass); only the 2-element softmax
    Jacobian of the raw weights and the temperature's scalar are formed from those sums afterwards."""

    @staticmethod
    def forward(ctx, a, b, raw_w, temperature, strategy):
        w0 = w1 = 0.5
        if raw_w is not None:
            w = F.softmax(raw_w.detach().float().cpu(), dim=0)
            w0, w1 = float(w[0]), float(w[1])
        temp = None if temperature is None else float(temperature.detach().float().reshape(-1)[0])
        code = _STRATEGY.get(strategy, _lib.FUSE_MEAN)
        fused = _fuse_forward(a, b, code, w0, w1, temp)
        ctx.save_for_backward(a, b)
        ctx.meta = (w0, w1, temp, code, raw_w, temperature)
        return fused

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        w0, w1, temp, code, raw_w, temperature = ctx.meta
        ga, gb, dots = ops.fuse_backward(g, a, b, code, w0, w1, temp, ctx.needs_input_grad[0], ctx.needs_input_grad[1])
        gw = gt = None
        need_w = raw_w is not None and ctx.needs_input_grad[2] and code == _lib.FUSE_WEIGHTED
        need_t = temperature is not None and ctx.needs_input_grad[3]
        if need_w or need_t:
            d = dots.cpu()   # three fp64 scalars
            if need_w:
                w = torch.tensor([w0, w1], dtype=torch.float64)
                # softmax Jacobian: dL/draw_i = w_i * (dot_i - sum_j w_j dot_j)
                gw = (w * (d[:2] - (w * d[:2]).sum())).to(raw_w.dtype).to(raw_w.device)
            if need_t:
                gt = (-d[2] / temp).reshape(temperature.shape).to(temperature.dtype).to(temperature.device)
        if ga is not None:
            ga = ga.to(a.dtype).to(a.device)
        if gb is not None:
            gb = gb.to(b.dtype).to(b.device)
        return ga, gb, gw, gt, None


class FogDensityAwareLoss(nn.Module):
    """Fog-density-aware loss (models/model.py:516-677): per-pixel CE/focal times (1 + s*fog_density),
    plus depth MSE; forward AND gradients come from one awx_fogloss launch."""

    def __init__(self, base_loss: str = "cross_entropy", depth_weight: float = 0.5,
                 fog_sensitivity: float = 2.0, depth_loss_weight: float = 0.1) -> None:
        super().__init__()
        self.base_loss = base_loss
        self.depth_weight = depth_weight
        self.fog_sensitivity = fog_sensitivity
        self.depth_loss_weight = depth_loss_weight
        logger.info(f"Initialized FogDensityAwareLoss with {base_loss} base loss (libawx)")

    def forward(self, predictions: Dict[str, torch.Tensor], targets: Dict[str, torch.Tensor],
                fog_density: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        seg_pred = predictions["segmentation"]
        label = targets["label"]
        use_depth = "depth" in predictions and self.depth_weight > 0
        pred_depth = predictions["depth"].squeeze(1) if use_depth else None
        if use_depth and fog_density is None:
            # model.py:593-597: density estimated from the predicted depth (global min/max and mean
            # gradient magnitude); autograd flows into the depth head through it
            fog_density = self._estimate_fog_density_from_depth(pred_depth)
        depth_tgt = targets["depth"] if (use_depth and "depth" in targets) else None
        seg, dep = ops_loss.fog_loss_terms(seg_pred, pred_depth if depth_tgt is not None else None, label,
                                           fog_density, depth_tgt, self.fog_sensitivity,
                                           self.base_loss == "focal")
        depth_loss = dep if depth_tgt is not None else 0.0
        total = seg + self.depth_loss_weight * depth_loss
        return {"total_loss": total, "segmentation_loss": seg, "depth_loss": depth_loss}

    def _estimate_fog_density_from_depth(self, depth: torch.Tensor) -> torch.Tensor:
        """model.py:644-677: density from the predicted depth (global min / max, mean gradient magnitude), forward
        and gradient in libawx (awx_depth_density_fwd / _bwd) so that path B trains the depth head as the
        reference does."""
        return ops_loss.depth_density(depth)
