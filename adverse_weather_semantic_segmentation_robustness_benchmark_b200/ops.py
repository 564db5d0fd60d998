"""Tensor-level wrappers over the C ABI (include/awx.h).

PyTorch is plumbing here: it owns device memory and the current stream; every function below
passes raw ``data_ptr()``s to libawx.so.  Nothing falls back to torch arithmetic.
"""

from __future__ import annotations

import ctypes as C
import functools
import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib

LN2 = math.log(2.0)
DEFAULT_AUROC_BINS = 4096


def require_cuda() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError(
            "adverse_weather_semantic_segmentation_robustness_benchmark_b200 needs a CUDA device: "
            "its arithmetic lives in libawx.so (sm_100a) and there is no CPU fallback.")
    return torch.device("cuda", torch.cuda.current_device())


def _tensors_in(args, kwargs):
    for v in list(args) + list(kwargs.values()):
        if torch.is_tensor(v):
            yield v
        elif isinstance(v, (list, tuple)):
            for u in v:
                if torch.is_tensor(u):
                    yield u


def device_scoped(fn):
    """Run `fn` with the device of its CUDA operands current.  The library sizes its grids from cudaGetDevice(),
    the wrappers allocate outputs / workspaces on the current device and hand over ITS current stream, and host
    operands are uploaded to it -- so a call on tensors of cuda:1 while cuda:0 is current must switch first
    (otherwise kernels on device 0 would dereference device-1 pointers).  Operands on two different CUDA devices
    raise instead of silently going through peer access."""
    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        if not torch.cuda.is_available():
            return fn(*args, **kwargs)          # fn raises the "needs a CUDA device" error itself
        devs = {t.device for t in _tensors_in(args, kwargs) if t.is_cuda}
        if len(devs) > 1:
            raise ValueError(f"{fn.__name__}: operands live on different CUDA devices: {sorted(map(str, devs))}")
        if not devs or next(iter(devs)).index == torch.cuda.current_device():
            return fn(*args, **kwargs)
        with torch.cuda.device(next(iter(devs))):
            return fn(*args, **kwargs)
    return wrapper


def to_device(t: torch.Tensor, dtype: Optional[torch.dtype] = None) -> torch.Tensor:
    """Contiguous CUDA copy/view of `t` (host tensors are uploaded)."""
    dev = require_cuda()
    if not t.is_cuda:
        t = t.to(dev, non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def label_code(t: torch.Tensor) -> int:
    if t.dtype == torch.uint8:
        return _lib.LABEL_U8
    if t.dtype == torch.int64:
        return _lib.LABEL_I64
    raise TypeError(f"labels must be uint8 or int64, got {t.dtype}")


def normalise_labels(t: torch.Tensor) -> torch.Tensor:
    """uint8 stays uint8 (the reference's wrap quirk depends on it); every other integer type is
    widened to int64, which is what torch's promotion does in the reference's expressions."""
    if t.dtype in (torch.uint8, torch.int64):
        return to_device(t)
    if t.dtype in (torch.int8, torch.int16, torch.int32, torch.bool):
        return to_device(t, torch.int64)
    raise TypeError(f"labels must be an integer tensor, got {t.dtype}")


@functools.lru_cache(maxsize=None)
def _ece_edge_floats(num_bins: int) -> tuple:
    """The edges as Python floats, once per bin count (score_config runs on every launch)."""
    return tuple(float(e) for e in torch.linspace(0, 1, num_bins + 1).numpy())


def ece_edges(num_bins: int) -> torch.Tensor:
    """Bin boundaries exactly as the reference builds them (evaluation/metrics.py:179)."""
    return torch.linspace(0, 1, num_bins + 1)


@dataclass
class Bins:
    """Host view of the int64 bins buffer (layout: include/awx.h, AwxBinsLayout)."""
    raw: np.ndarray
    num_classes: int
    ece_bins: int
    auroc_bins: int

    def __post_init__(self):
        self.layout = _lib.bins_layout(self.num_classes, self.ece_bins, self.auroc_bins)

    def _seg(self, off, n):
        return self.raw[off:off + n]

    @property
    def confusion(self) -> np.ndarray:
        c = self.num_classes
        return self._seg(self.layout.confusion, c * c).reshape(c, c)

    @property
    def ece_count(self):
        return self._seg(self.layout.ece_count, self.ece_bins)

    @property
    def ece_correct(self):
        return self._seg(self.layout.ece_correct, self.ece_bins)

    @property
    def ece_conf_sum(self) -> np.ndarray:
        """Per-bin sum of confidences (float64), from the 2^-31 fixed-point words."""
        hi = self._seg(self.layout.ece_conf_hi, self.ece_bins)
        lo = self._seg(self.layout.ece_conf_lo, self.ece_bins)
        return np.array([(int(h) * (1 << 32) + int(l)) / float(1 << 31) for h, l in zip(hi, lo)], dtype=np.float64)

    @property
    def auroc_pos(self):
        return self._seg(self.layout.auroc_pos, self.auroc_bins)

    @property
    def auroc_neg(self):
        return self._seg(self.layout.auroc_neg, self.auroc_bins)

    def counter(self, which: int) -> int:
        return int(self.raw[self.layout.counters + which])


def new_bins(num_classes: int, ece_bins: int = 15, auroc_bins: int = 0, device=None) -> torch.Tensor:
    lay = _lib.bins_layout(num_classes, ece_bins, auroc_bins)
    return torch.zeros(lay.total_words, dtype=torch.int64, device=require_cuda() if device is None else device)


def score_config(num_classes: int, strategy: int, w0: float, w1: float, temperature: Optional[float], label_dtype: int,
                 ignore_index: int = 255, ece_bins: int = 15, auroc_bins: int = 0, auroc_hi: float = LN2):
    """AwxScoreConfig (host struct) with the reference's torch.linspace ECE edges."""
    cfg = _lib.ScoreConfig()
    cfg.num_classes, cfg.strategy = num_classes, strategy
    cfg.w0, cfg.w1 = float(w0), float(w1)
    cfg.use_temperature = 0 if temperature is None else 1
    cfg.temperature = 1.0 if temperature is None else float(temperature)
    cfg.label_dtype = label_dtype
    cfg.ignore_index = ignore_index
    cfg.ece_bins, cfg.auroc_bins, cfg.auroc_hi = ece_bins, auroc_bins, float(auroc_hi)
    cfg.ece_edges[:ece_bins + 1] = _ece_edge_floats(ece_bins)
    return cfg


@device_scoped
def score(logits_a: torch.Tensor, logits_b: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None, *,
          strategy: int = _lib.FUSE_SINGLE, w0: float = 0.5, w1: float = 0.5,
          temperature: Optional[float] = None, ignore_index: int = 255,
          ece_bins: int = 15, auroc_bins: int = 0, auroc_hi: float = LN2,
          bins: Optional[torch.Tensor] = None,
          want_pred: Optional[torch.dtype] = None, want_fused: bool = False, want_conf: bool = False,
          want_mi: bool = False, want_js: bool = False) -> dict:
    """One fused pass over [B,C,H,W] fp32 logits (one or two members).  Returns a dict with
    ``bins`` (device int64 buffer, accumulated into if given) and the requested maps."""
    lib = _lib.load()
    a = to_device(logits_a, torch.float32)
    if a.dim() != 4:
        raise ValueError(f"logits must be [B,C,H,W], got shape {tuple(a.shape)}")
    bsz, ncls, h, w = a.shape
    ens = strategy != _lib.FUSE_SINGLE
    b = None
    if ens:
        if logits_b is None:
            raise ValueError("an ensemble strategy needs two members")
        b = to_device(logits_b, torch.float32)
        if b.shape != a.shape:
            raise ValueError(f"member shapes differ: {tuple(a.shape)} vs {tuple(b.shape)}")
    lab = None
    if labels is not None:
        lab = normalise_labels(labels)
        if lab.numel() != bsz * h * w:
            raise ValueError(f"labels have {lab.numel()} elements, expected {bsz * h * w}")
    nb_auroc = auroc_bins if ens else 0
    cfg = score_config(ncls, strategy, w0, w1, temperature, _lib.LABEL_I64 if lab is None else label_code(lab),
                       ignore_index, ece_bins, nb_auroc, auroc_hi)
    dev = a.device
    out = {}
    if lab is not None:
        if bins is None:
            bins = new_bins(ncls, ece_bins, nb_auroc)
        out["bins"] = bins
    maps = _lib.ScoreMaps()
    if want_pred is not None:
        if want_pred not in (torch.uint8, torch.int64):
            raise TypeError("want_pred must be torch.uint8 or torch.int64")
        out["pred"] = torch.empty((bsz, h, w), dtype=want_pred, device=dev)
        maps.pred = out["pred"].data_ptr()
        maps.pred_dtype = _lib.PRED_U8 if want_pred == torch.uint8 else _lib.PRED_I64
    if want_fused:
        out["fused"] = torch.empty_like(a)
        maps.fused = out["fused"].data_ptr()
    if want_conf:
        out["conf"] = torch.empty((bsz, h, w), dtype=torch.float32, device=dev)
        maps.conf = out["conf"].data_ptr()
    if want_mi and ens:
        out["mi"] = torch.empty((bsz, h, w), dtype=torch.float32, device=dev)
        maps.mi = out["mi"].data_ptr()
    if want_js and ens:
        out["js"] = torch.empty((bsz, h, w), dtype=torch.float32, device=dev)
        maps.js = out["js"].data_ptr()
    rc = lib.awx_score(_ptr(a), _ptr(b), _ptr(lab), bsz, h * w, C.byref(cfg),
                       _ptr(bins) if lab is not None else None, C.byref(maps), _stream())
    _lib.check(rc, "awx_score")
    return out


def read_bins(bins: torch.Tensor, num_classes: int, ece_bins: int = 15, auroc_bins: int = 0) -> Bins:
    """Device -> host read of a bins buffer (synchronises the current stream)."""
    return Bins(bins.cpu().numpy(), num_classes, ece_bins, auroc_bins)


@device_scoped
def member_variance(logits_a: torch.Tensor, logits_b: torch.Tensor) -> torch.Tensor:
    lib = _lib.load()
    a = to_device(logits_a, torch.float32)
    b = to_device(logits_b, torch.float32)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError("member_variance needs two [B,C,H,W] tensors of equal shape")
    out = torch.empty_like(a)
    bsz, ncls, h, w = a.shape
    _lib.check(lib.awx_member_variance(_ptr(a), _ptr(b), _ptr(out), bsz, ncls, h * w, _stream()), "awx_member_variance")
    return out


@device_scoped
def members_n(members, labels: Optional[torch.Tensor] = None, *, ignore_index: int = 255, auroc_bins: int = 0,
              auroc_hi: Optional[float] = None, want_mi: bool = False, want_var: bool = False) -> dict:
    """awx_members_n: disagreement of a list of N >= 2 members ([B,C,H,W] fp32 each).  Returns a dict with the
    requested maps (``mi`` [B,H,W], ``var`` [B,C,H,W]) and, with labels, ``pos`` / ``neg`` (int64 [auroc_bins]) and
    ``counters`` (int64 [NUM_COUNTERS]).  The MI of N members lies in [0, ln N): that is the default histogram range."""
    lib = _lib.load()
    ms = [to_device(m, torch.float32) for m in members]
    if len(ms) < 2:
        raise ValueError("Need at least 2 predictions for disagreement computation")
    if any(m.shape != ms[0].shape or m.dim() != 4 for m in ms):
        raise ValueError("members must be [B,C,H,W] tensors of equal shape")
    bsz, ncls, h, w = ms[0].shape
    dev = ms[0].device
    lab = None
    if labels is not None:
        lab = normalise_labels(labels)
        if lab.numel() != bsz * h * w:
            raise ValueError(f"labels have {lab.numel()} elements, expected {bsz * h * w}")
    out = {}
    if want_mi:
        out["mi"] = torch.empty((bsz, h, w), dtype=torch.float32, device=dev)
    if want_var:
        out["var"] = torch.empty_like(ms[0])
    if lab is not None:
        out["pos"] = torch.zeros(max(auroc_bins, 1), dtype=torch.int64, device=dev)
        out["neg"] = torch.zeros(max(auroc_bins, 1), dtype=torch.int64, device=dev)
        out["counters"] = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    ptrs = (C.c_void_p * len(ms))(*[m.data_ptr() for m in ms])
    hi = float(auroc_hi) if auroc_hi is not None else math.log(len(ms))
    rc = lib.awx_members_n(ptrs, len(ms), _ptr(lab), _lib.LABEL_I64 if lab is None else label_code(lab), bsz, ncls, h * w,
                           ignore_index, auroc_bins if lab is not None else 0, hi, _ptr(out.get("pos")), _ptr(out.get("neg")),
                           _ptr(out.get("counters")), _ptr(out.get("mi")), _ptr(out.get("var")), _stream())
    _lib.check(rc, "awx_members_n")
    return out


@device_scoped
def fuse_forward(logits_a: torch.Tensor, logits_b: torch.Tensor, strategy: int, w0: float, w1: float,
                 temperature: Optional[float]) -> torch.Tensor:
    """awx_fuse_forward: the fused logits of two [B,C,H,W] members (EnsembleModel.forward), nothing else."""
    lib = _lib.load()
    a = to_device(logits_a, torch.float32)
    b = to_device(logits_b, torch.float32)
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError("fuse_forward needs two [B,C,H,W] tensors of equal shape")
    bsz, ncls, h, w = a.shape
    out = torch.empty_like(a)
    rc = lib.awx_fuse_forward(_ptr(a), _ptr(b), _ptr(out), bsz, ncls, h * w, strategy, float(w0), float(w1),
                              1.0 if temperature is None else float(temperature), 0 if temperature is None else 1,
                              _stream())
    _lib.check(rc, "awx_fuse_forward")
    return out


@device_scoped
def fuse_backward(grad_fused: torch.Tensor, logits_a: torch.Tensor, logits_b: torch.Tensor, strategy: int, w0: float,
                  w1: float, temperature: Optional[float], want_a: bool = True, want_b: bool = True):
    """awx_fuse_backward: (grad_a | None, grad_b | None, dots fp64[3] device) for the fusion's backward pass."""
    lib = _lib.load()
    g = to_device(grad_fused, torch.float32)
    a = to_device(logits_a, torch.float32)
    b = to_device(logits_b, torch.float32)
    if not (g.shape == a.shape == b.shape) or a.dim() != 4:
        raise ValueError("fuse_backward needs three [B,C,H,W] tensors of equal shape")
    bsz, ncls, h, w = a.shape
    ga = torch.empty_like(a) if want_a else None
    gb = torch.empty_like(b) if want_b else None
    dots = torch.zeros(3, dtype=torch.float64, device=a.device)
    ws = torch.empty(int(lib.awx_fuse_backward_workspace_bytes()), dtype=torch.uint8, device=a.device)
    rc = lib.awx_fuse_backward(_ptr(g), _ptr(a), _ptr(b), _ptr(ga), _ptr(gb), bsz, ncls, h * w, strategy, float(w0), float(w1),
                               1.0 if temperature is None else float(temperature), 0 if temperature is None else 1,
                               _ptr(dots), _ptr(ws), _stream())
    _lib.check(rc, "awx_fuse_backward")
    return ga, gb, dots


@device_scoped
def confusion(pred: torch.Tensor, labels: torch.Tensor, num_classes: int, ignore_index: int = 255):
    """(confusion int64 [C,C] device tensor, counters int64 [NUM_COUNTERS] device tensor) from prediction maps."""
    lib = _lib.load()
    lab = normalise_labels(labels).reshape(-1)
    if pred.dtype not in (torch.uint8, torch.int64):
        pred = pred.to(torch.int64)
    prd = to_device(pred).reshape(-1)
    if prd.numel() != lab.numel():
        raise ValueError(f"predictions ({prd.numel()}) and targets ({lab.numel()}) differ in size")
    dev = require_cuda()
    cm = torch.zeros(num_classes * num_classes, dtype=torch.int64, device=dev)
    cnt = torch.zeros(_lib.NUM_COUNTERS, dtype=torch.int64, device=dev)
    rc = lib.awx_confusion(_ptr(prd), _lib.PRED_U8 if prd.dtype == torch.uint8 else _lib.PRED_I64,
                           _ptr(lab), label_code(lab), prd.numel(), num_classes, ignore_index,
                           _ptr(cm), _ptr(cnt), _stream())
    _lib.check(rc, "awx_confusion")
    return cm.view(num_classes, num_classes), cnt


def corrupt_workspace(batch: int, height: int, width: int) -> torch.Tensor:
    n = _lib.load().awx_corrupt_workspace_bytes(batch, height, width)
    return torch.empty(max(int(n), 16), dtype=torch.uint8, device=require_cuda())


@device_scoped
def corrupt(images: torch.Tensor, params: np.ndarray, field: Optional[torch.Tensor] = None,
            items: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
            workspace: Optional[torch.Tensor] = None, norm_out: Optional[torch.Tensor] = None,
            norm_params=None, write_u8: bool = True):
    """awx_corrupt on a device-resident uint8 [B,H,W,3] batch.  `params`: host records
    (``_lib.CORRUPT_PARAMS_DTYPE``); `field`: device fp32/fp64 depth / noise values; `items`:
    device int32 [n,5] drops / flakes.
    With `norm_out` (fp32 / bf16 [B,3,H,W]) and `norm_params` = (mean*255, 1/(std*255)) fp32 triples, the
    dataset's Normalize + ToTensorV2 is fused into the kernels' epilogue (awx_corrupt_normalized);
    `write_u8=False` then skips the uint8 frame.  Returns `out` (or `norm_out` when write_u8 is False)."""
    lib = _lib.load()
    if not images.is_cuda or images.dtype != torch.uint8 or not images.is_contiguous():
        raise ValueError("images must be a contiguous CUDA uint8 tensor")
    b, h, w, ch = images.shape
    if ch != 3:
        raise ValueError("images must be [B,H,W,3]")
    if params.dtype != _lib.CORRUPT_PARAMS_DTYPE or len(params) != b:
        raise ValueError("params must be one CORRUPT_PARAMS_DTYPE record per frame")
    params = np.ascontiguousarray(params)
    if not write_u8 and norm_out is None:
        raise ValueError("write_u8=False needs norm_out")
    if out is None and write_u8:
        out = torch.empty_like(images)
    if workspace is None:
        workspace = corrupt_workspace(b, h, w)
    fdt = _lib.F64
    if field is not None:
        if field.dtype == torch.float32:
            fdt = _lib.F32
        elif field.dtype != torch.float64:
            raise TypeError("field must be float32 or float64")
        field = field.contiguous()
    n_items = 0
    if items is not None:
        if items.dtype != torch.int32:
            raise TypeError("items must be int32 [n,5]")
        items = items.contiguous()
        n_items = items.shape[0]
    if norm_out is None:
        rc = lib.awx_corrupt(_ptr(images), _ptr(out), b, h, w, params.ctypes.data_as(C.c_void_p), _ptr(field), fdt,
                             _ptr(items), n_items, _ptr(workspace), _stream())
        _lib.check(rc, "awx_corrupt")
        return out
    if norm_out.dtype not in (torch.float32, torch.bfloat16) or tuple(norm_out.shape) != (b, 3, h, w) \
            or not norm_out.is_cuda or not norm_out.is_contiguous():
        raise ValueError("norm_out must be a contiguous CUDA fp32 / bf16 [B,3,H,W] tensor")
    m, r = norm_params
    m = np.ascontiguousarray(m, dtype=np.float32)
    r = np.ascontiguousarray(r, dtype=np.float32)
    rc = lib.awx_corrupt_normalized(_ptr(images), _ptr(out) if write_u8 else None, _ptr(norm_out),
                                    _lib.F32 if norm_out.dtype == torch.float32 else _lib.BF16,
                                    m.ctypes.data_as(C.c_void_p), r.ctypes.data_as(C.c_void_p), b, h, w,
                                    params.ctypes.data_as(C.c_void_p), _ptr(field), fdt, _ptr(items), n_items,
                                    _ptr(workspace), _stream())
    _lib.check(rc, "awx_corrupt_normalized")
    return out if write_u8 else norm_out


@device_scoped
def corrupt_score(images: torch.Tensor, params: np.ndarray, field: Optional[torch.Tensor], items: Optional[torch.Tensor],
                  out: torch.Tensor, workspace: torch.Tensor, logits_a: torch.Tensor, logits_b: Optional[torch.Tensor],
                  labels: torch.Tensor, cfg, bins: torch.Tensor) -> None:
    """awx_corrupt_score: one C call per condition -- corrupt `images` into `out` and score the logits into
    `bins`.  Everything device resident, contiguous and pre-validated (this is the sweep driver's inner call);
    `cfg` comes from score_config()."""
    lib = _lib.load()
    b, h, w, _ = images.shape
    fdt = _lib.F32 if (field is not None and field.dtype == torch.float32) else _lib.F64
    rc = lib.awx_corrupt_score(_ptr(images), _ptr(out), h, w, params.ctypes.data_as(C.c_void_p), _ptr(field), fdt,
                               _ptr(items), 0 if items is None else items.shape[0], _ptr(workspace),
                               _ptr(logits_a), _ptr(logits_b), _ptr(labels), b, C.byref(cfg), _ptr(bins), None, _stream())
    _lib.check(rc, "awx_corrupt_score")

