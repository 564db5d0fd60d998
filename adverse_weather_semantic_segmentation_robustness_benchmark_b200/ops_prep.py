"""Tensor-level wrappers over the C ABI for the rows either side of the hot path:
Normalize + CHW, style transfer, local contrast / fog-density map, depth estimation, temperature grid.
Same rules as ops.py: torch owns memory and the stream, libawx.so does the arithmetic, no fallback."""

from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .ops import _ptr, _stream, device_scoped, label_code, normalise_labels, require_cuda, to_device

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)
_DT = {torch.float32: _lib.F32, torch.float64: _lib.F64, torch.bfloat16: _lib.BF16, torch.uint8: _lib.U8}


def normalize_params(mean: Sequence[float], std: Sequence[float], max_pixel_value: float = 255.0):
    """fp32 mean*max and 1/(std*max) exactly as albumentations forms them (host scalars)."""
    m = np.array(mean, dtype=np.float32)
    m *= max_pixel_value
    s = np.array(std, dtype=np.float32)
    s *= max_pixel_value
    return np.ascontiguousarray(m), np.ascontiguousarray(np.reciprocal(s, dtype=np.float32))


@device_scoped
def normalize_chw(images: torch.Tensor, mean: Sequence[float] = IMAGENET_MEAN, std: Sequence[float] = IMAGENET_STD,
                  max_pixel_value: float = 255.0, out_dtype: torch.dtype = torch.float32,
                  out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """uint8 [B,H,W,3] -> fp32 / bf16 [B,3,H,W]: Normalize(mean, std) + ToTensorV2 (loader.py:196-199)."""
    lib = _lib.load()
    dev = require_cuda()
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError(f"images must be uint8 [B,H,W,3], got {images.dtype} {tuple(images.shape)}")
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise TypeError("out_dtype must be torch.float32 or torch.bfloat16")
    images = to_device(images)
    b, h, w, _ = images.shape
    if out is None:
        out = torch.empty((b, 3, h, w), dtype=out_dtype, device=dev)
    m, r = normalize_params(mean, std, max_pixel_value)
    rc = lib.awx_normalize_chw(_ptr(images), _ptr(out), _DT[out.dtype], b, h, w, m.ctypes.data_as(C.c_void_p),
                               r.ctypes.data_as(C.c_void_p), _stream())
    _lib.check(rc, "awx_normalize_chw")
    return out


STYLE = {"fog": (0.8, 30.0, None), "rain": (1.2, -10.0, 1.1), "snow": (0.9, 20.0, None), "night": (0.4, -20.0, 1.3)}


@device_scoped
def style_transfer(images: torch.Tensor, weather_type: str, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """cv2.convertScaleAbs(alpha, beta) + blue-channel gain on uint8 [...,3] frames (loader.py:364-385)."""
    lib = _lib.load()
    if images.dtype != torch.uint8 or images.shape[-1] != 3:
        raise ValueError("images must be uint8 [...,3]")
    images = to_device(images)
    if weather_type not in STYLE:
        return images
    alpha, beta, gain = STYLE[weather_type]
    if out is None:
        out = torch.empty_like(images)
    rc = lib.awx_style_transfer(_ptr(images), _ptr(out), images.numel() // 3, alpha, beta, 0.0 if gain is None else gain,
                                0 if gain is None else 1, _stream())
    _lib.check(rc, "awx_style_transfer")
    return out


# ---------------------------------------------------------------------- percentile scalars
def percentile_indices(n: int, q: float, dtype=np.float32) -> Tuple[int, int, np.floating]:
    """Index arithmetic of np.percentile(a, q) (method 'linear') for a 1-D array of `n` values of
    `dtype`: (previous index, next index, gamma).  NumPy divides q by dtype(100) and forms the virtual
    index (n - 1) * q in the ARRAY's precision (numpy/lib/_function_base_impl.py, _quantile)."""
    qq = np.true_divide(q, dtype(100)) if np.issubdtype(dtype, np.floating) else np.true_divide(q, 100)
    vi = np.asanyarray((n - 1) * qq)
    prev = int(np.floor(vi))
    nxt = min(prev + 1, n - 1)
    prev = min(max(prev, 0), n - 1)
    gamma = np.asanyarray(np.asanyarray(vi - np.intp(prev)), dtype=vi.dtype)
    return prev, nxt, gamma[()]


def lerp(a, b, t):
    """numpy's _lerp on scalars: a + (b-a)*t, or b - (b-a)*(1-t) when t >= 0.5."""
    diff = np.subtract(b, a)
    r = np.add(a, diff * t)
    if t >= 0.5:
        r = np.subtract(b, diff * (1 - t)).astype(r.dtype)
    return r


@device_scoped
def fog_density_map(images: torch.Tensor, depth: torch.Tensor) -> torch.Tensor:
    """get_fog_density_map (preprocessing.py:250-288) for a batch: images [B,H,W,3] (uint8, or float in
    [0,1] as the reference's signature says), depth [B,H,W] fp64/fp32 -> fog density [B,H,W] (depth's dtype).
    Device: gray -> 5x5 contrast -> the two order statistics of the 95th percentile; host: NumPy's lerp of
    those two scalars per frame; device: depth maximum + final blend."""
    lib = _lib.load()
    dev = require_cuda()
    if images.dim() != 4 or images.shape[-1] != 3 or images.dtype not in (torch.uint8, torch.float32, torch.float64):
        raise ValueError("images must be [B,H,W,3] uint8 / float32 / float64")
    images = to_device(images)
    depth = to_device(depth)
    if depth.dtype not in (torch.float32, torch.float64):
        depth = depth.to(torch.float64)
    b, h, w, _ = images.shape
    if tuple(depth.shape) != (b, h, w):
        raise ValueError(f"depth must be [B,H,W] = {(b, h, w)}, got {tuple(depth.shape)}")
    n = h * w
    lo, hi, gamma = percentile_indices(n, 95, np.float32)
    contrast = torch.empty((b, h, w), dtype=torch.float32, device=dev)
    stats = torch.empty((b, 2), dtype=torch.float32, device=dev)
    ws = torch.empty(max(int(lib.awx_fog_density_workspace_bytes(b)), 16), dtype=torch.uint8, device=dev)
    rc = lib.awx_local_contrast(_ptr(images), _DT[images.dtype], _ptr(contrast), b, h, w, lo, hi, _ptr(stats), _ptr(ws), _stream())
    _lib.check(rc, "awx_local_contrast")
    st = stats.cpu().numpy()  # 8 bytes per frame: the two order statistics
    # max_contrast + 1e-8 in fp32 (np.float32 + python float stays fp32)
    denom = np.array([lerp(st[i, 0], st[i, 1], gamma) + 1e-8 for i in range(b)], dtype=np.float32)
    denom_d = torch.from_numpy(denom).to(dev)
    out = torch.empty_like(depth)
    rc = lib.awx_fog_density_finish(_ptr(contrast), _ptr(depth), _DT[depth.dtype], _ptr(denom_d), _ptr(out), b, n, _ptr(ws), _stream())
    _lib.check(rc, "awx_fog_density_finish")
    return out


@device_scoped
def local_contrast(images: torch.Tensor) -> torch.Tensor:
    """The fp32 [B,H,W] contrast map alone (intermediate of fog_density_map)."""
    lib = _lib.load()
    dev = require_cuda()
    images = to_device(images)
    b, h, w, _ = images.shape
    contrast = torch.empty((b, h, w), dtype=torch.float32, device=dev)
    stats = torch.empty((b, 2), dtype=torch.float32, device=dev)
    ws = torch.empty(max(int(lib.awx_fog_density_workspace_bytes(b)), 16), dtype=torch.uint8, device=dev)
    rc = lib.awx_local_contrast(_ptr(images), _DT[images.dtype], _ptr(contrast), b, h, w, 0, 0, _ptr(stats), _ptr(ws), _stream())
    _lib.check(rc, "awx_local_contrast")
    return contrast


@device_scoped
def estimate_depth(images: torch.Tensor, weights: np.ndarray) -> torch.Tensor:
    """DepthEstimationPreprocessor._geometric_depth_estimation (preprocessing.py:332-367): uint8 [B,H,W,3]
    -> fp64 [B,H,W].  `weights`: scipy's sigma=2 Gaussian taps (host, fp64)."""
    lib = _lib.load()
    dev = require_cuda()
    if images.dtype != torch.uint8 or images.dim() != 4 or images.shape[-1] != 3:
        raise ValueError("images must be uint8 [B,H,W,3]")
    images = to_device(images)
    b, h, w, _ = images.shape
    out = torch.empty((b, h, w), dtype=torch.float64, device=dev)
    tmp = torch.empty_like(out)
    amax = torch.empty(max(b, 1), dtype=torch.int32, device=dev)
    wts = np.ascontiguousarray(weights, dtype=np.float64)
    rc = lib.awx_estimate_depth(_ptr(images), _ptr(out), _ptr(tmp), b, h, w, wts.ctypes.data_as(C.c_void_p), (len(wts) - 1) // 2,
                                _ptr(amax), _stream())
    _lib.check(rc, "awx_estimate_depth")
    return out


@device_scoped
def temperature_nll(logits: torch.Tensor, targets: torch.Tensor, temperatures: torch.Tensor, ignore_index: int = 255):
    """Sum over valid rows of cross_entropy(rows / T) for every T of the grid, in one pass.
    rows = logits.view(-1, C) exactly as the reference flattens (metrics.py:305).  Returns
    (fp64 sums [n_T], valid rows, labels outside [0,C))."""
    lib = _lib.load()
    dev = require_cuda()
    c = logits.size(1)
    logits = to_device(logits, torch.float32)
    targets = to_device(normalise_labels(targets)).reshape(-1)
    rows = logits.numel() // c
    if targets.numel() != rows:
        raise ValueError(f"{targets.numel()} targets for {rows} rows of {c} logits")
    temps = np.ascontiguousarray(temperatures.detach().cpu().numpy(), dtype=np.float32)
    nt = len(temps)
    sums = torch.zeros(nt + 2, dtype=torch.float64, device=dev)
    ws = torch.empty(max(int(lib.awx_temperature_workspace_bytes(nt)), 16), dtype=torch.uint8, device=dev)
    rc = lib.awx_temperature_nll(_ptr(logits), _ptr(targets), label_code(targets), rows, c, ignore_index,
                                 temps.ctypes.data_as(C.c_void_p), nt, _ptr(sums), _ptr(ws), _stream())
    _lib.check(rc, "awx_temperature_nll")
    host = sums.cpu().numpy()
    return host[:nt], int(host[nt]), int(host[nt + 1])
