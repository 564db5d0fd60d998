"""Autograd-capable wrappers of the loss / fusion kernels (awx_fogloss, awx_scale_inplace, awx_score)."""

from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib, ops
from .ops import _ptr, _stream, device_scoped


def _loss_workspace() -> torch.Tensor:
    n = _lib.load().awx_fogloss_workspace_bytes()
    return torch.empty(int(n), dtype=torch.uint8, device=ops.require_cuda())


@device_scoped
def fogloss_raw(logits: torch.Tensor, labels: torch.Tensor, fog_density: Optional[torch.Tensor],
                depth_pred: Optional[torch.Tensor], depth_tgt: Optional[torch.Tensor],
                fog_sensitivity: float, focal: bool, want_grads: bool, want_dfog: bool = False):
    """One launch: returns (sums fp64[2] device, dlogits|None, ddepth|None).  Raises IndexError for
    labels outside [0,C) as torch's cross_entropy does."""
    lib = _lib.load()
    x = ops.to_device(logits, torch.float32)
    bsz, ncls, h, w = x.shape
    lab = ops.normalise_labels(labels)
    if lab.numel() != bsz * h * w:
        raise ValueError(f"labels have {lab.numel()} elements, expected {bsz * h * w}")
    fd = None if fog_density is None else ops.to_device(fog_density, torch.float32)
    dp = None if depth_pred is None else ops.to_device(depth_pred, torch.float32)
    dt = None if depth_tgt is None else ops.to_device(depth_tgt, torch.float32)
    for name, t in (("fog_density", fd), ("depth prediction", dp), ("depth target", dt)):
        if t is not None and t.numel() != bsz * h * w:
            raise ValueError(f"{name} has {t.numel()} elements, expected {bsz * h * w}")
    dev = x.device
    sums = torch.zeros(2, dtype=torch.float64, device=dev)
    bad = torch.zeros(1, dtype=torch.int64, device=dev)
    dlogits = torch.empty_like(x) if want_grads else None
    ddepth = torch.empty((bsz, h, w), dtype=torch.float32, device=dev) if (want_grads and dp is not None and dt is not None) else None
    dfog = torch.empty((bsz, h, w), dtype=torch.float32, device=dev) if (want_dfog and fd is not None) else None
    ws = _loss_workspace()
    rc = lib.awx_fogloss(_ptr(x), _ptr(lab), ops.label_code(lab), _ptr(fd), _ptr(dp), _ptr(dt),
                         float(fog_sensitivity), 1 if focal else 0, bsz, ncls, h * w,
                         _ptr(sums), _ptr(dlogits), _ptr(ddepth), _ptr(dfog), _ptr(bad), _ptr(ws), _stream())
    _lib.check(rc, "awx_fogloss")
    return sums, dlogits, ddepth, bad, dfog


@device_scoped
def scale_inplace(x: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    """x *= scale (device scalar fp32) through awx_scale_inplace."""
    lib = _lib.load()
    s = ops.to_device(scale.reshape(1), torch.float32)
    _lib.check(lib.awx_scale_inplace(_ptr(x), x.numel(), _ptr(s), _stream()), "awx_scale_inplace")
    return x


class _FogLossFn(torch.autograd.Function):
    """(seg_mean, depth_mean) = f(logits, depth_pred); gradients were produced by the forward launch
    (unscaled); backward only multiplies by the incoming scalar grads."""

    @staticmethod
    def forward(ctx, logits, depth_pred, fog_density, labels, depth_tgt, sens, focal):
        fog_grad = fog_density is not None and fog_density.requires_grad
        need = logits.requires_grad or (depth_pred is not None and depth_pred.requires_grad) or fog_grad
        sums, dlogits, ddepth, bad, dfog = fogloss_raw(logits, labels, fog_density, depth_pred, depth_tgt, sens, focal,
                                                       need, fog_grad)
        ctx.dfog = dfog
        ctx.fog_shape = None if fog_density is None else fog_density.shape
        n = float(logits.shape[0] * logits.shape[2] * logits.shape[3])
        host = torch.cat([sums, bad.double()]).cpu()      # one D2H: the loss values and the bad-label count
        if int(host[2]) != 0:
            raise IndexError(f"Target out of bounds ({int(host[2])} labels outside [0,{logits.shape[1]}))")
        seg = (sums[0] / n).to(torch.float32)
        dep = (sums[1] / n).to(torch.float32)
        ctx.dlogits, ctx.ddepth = dlogits, ddepth
        ctx.depth_shape = None if depth_pred is None else depth_pred.shape
        ctx.logits_dtype = logits.dtype
        ctx.devices = (logits.device, None if depth_pred is None else depth_pred.device,
                       None if fog_density is None else fog_density.device)
        return seg, dep

    @staticmethod
    def backward(ctx, g_seg, g_dep):
        dlogits, ddepth = ctx.dlogits, ctx.ddepth
        gl = gd = None
        if dlogits is not None and ctx.needs_input_grad[0]:
            gl = scale_inplace(dlogits, g_seg).to(ctx.logits_dtype).to(ctx.devices[0])
        if ddepth is not None and ctx.needs_input_grad[1]:
            gd = scale_inplace(ddepth, g_dep).reshape(ctx.depth_shape).to(ctx.devices[1])
        gf = None
        if ctx.dfog is not None and ctx.needs_input_grad[2]:
            gf = scale_inplace(ctx.dfog, g_seg).reshape(ctx.fog_shape).to(ctx.devices[2])
        ctx.dlogits = ctx.ddepth = ctx.dfog = None
        return gl, gd, gf, None, None, None, None


class _DepthDensityFn(torch.autograd.Function):
    """fog density estimated from a predicted depth map (model.py:644-677): awx_depth_density_fwd / _bwd."""

    @staticmethod
    @device_scoped
    def forward(ctx, depth):
        lib = _lib.load()
        d = ops.to_device(depth.detach(), torch.float32)
        if d.dim() != 3:
            raise ValueError(f"depth must be [B,H,W], got {tuple(d.shape)}")
        b, h, w = d.shape
        out = torch.empty_like(d)
        ws = torch.empty(int(lib.awx_depth_density_workspace_bytes()), dtype=torch.uint8, device=d.device)
        _lib.check(lib.awx_depth_density_fwd(_ptr(d), _ptr(out), b, h, w, _ptr(ws), _stream()), "awx_depth_density_fwd")
        ctx.save_for_backward(d)
        ctx.src = (depth.device, depth.dtype)
        return out

    @staticmethod
    @device_scoped
    def backward(ctx, g):
        lib = _lib.load()
        (d,) = ctx.saved_tensors
        b, h, w = d.shape
        gg = ops.to_device(g, torch.float32)
        gd = torch.empty_like(d)
        ws = torch.empty(int(lib.awx_depth_density_workspace_bytes()), dtype=torch.uint8, device=d.device)
        _lib.check(lib.awx_depth_density_bwd(_ptr(d), _ptr(gg), _ptr(gd), b, h, w, _ptr(ws), _stream()), "awx_depth_density_bwd")
        return gd.to(ctx.src[1]).to(ctx.src[0])


def depth_density(depth: torch.Tensor) -> torch.Tensor:
    """Differentiable fog density from a [B,H,W] depth map; the result lives on the CUDA device."""
    return _DepthDensityFn.apply(depth)


def fog_loss_terms(logits, depth_pred, labels, fog_density, depth_tgt, sens: float, focal: bool):
    return _FogLossFn.apply(logits, depth_pred, fog_density, labels, depth_tgt, sens, focal)


def temperature_grid_search(logits: torch.Tensor, targets: torch.Tensor) -> float:
    """ConfidenceCalibration.optimize_temperature (evaluation/metrics.py:283-321): T in
    linspace(0.1, 10, 100) minimising the mean NLL over non-ignored rows; first minimum wins.
    One awx_temperature_nll pass over the logits evaluates the whole grid.  Rows are
    ``logits.view(-1, C)`` exactly as the reference flattens them (for an NCHW tensor that is not a
    pixel's class vector; pass [N, C] logits for the per-pixel meaning)."""
    from . import ops_prep
    temps = torch.linspace(0.1, 10.0, 100)
    sums, n_valid, n_bad = ops_prep.temperature_nll(logits, targets, temps)
    if n_bad:
        raise IndexError("Target out of bounds (labels outside [0, C) other than the ignored 255)")
    best_t, best_nll = 1.0, float("inf")
    for temp, total in zip(temps, sums):
        # F.cross_entropy returns the fp32 mean; no valid rows -> nan, which never compares smaller
        nll = float(np.float32(total / n_valid)) if n_valid else float("nan")
        if nll < best_nll:
            best_nll, best_t = nll, temp.item()
    return best_t
