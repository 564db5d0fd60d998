"""ctypes binding of libawx.so (the C ABI declared in include/awx.h).

There is no CPU fallback: if the library cannot be loaded (or rebuilt from csrc/ with nvcc when
it is missing or stale) every entry point of this package raises ``RuntimeError``.
"""

from __future__ import annotations

import ctypes as C
import fcntl
import os
import threading

import numpy as np

from . import build as _build

MAX_CLASSES = 64
MAX_ECE_BINS = 64
MAX_AUROC_BINS = 8192

FUSE_SINGLE, FUSE_WEIGHTED, FUSE_MAXCONF, FUSE_MEAN = 0, 1, 2, 3
LABEL_U8, LABEL_I64 = 0, 1
PRED_U8, PRED_I64 = 0, 1
CLEAN, FOG, RAIN, SNOW, NIGHT = 0, 1, 2, 3, 4
F32, F64, BF16, U8 = 0, 1, 2, 3
KIND_CODES = {"clean": CLEAN, "fog": FOG, "rain": RAIN, "snow": SNOW, "night": NIGHT}

NUM_COUNTERS = 16
(CNT_VALID, CNT_CORRECT, CNT_BAD_LABEL, CNT_ECE_AMBIG, CNT_ENS_WRONG, CNT_PICK_AMBIG, CNT_NO_BIN, CNT_PIXELS,
 CNT_MARG_AMBIG, CNT_EPRED_AMBIG) = range(10)


class ScoreConfig(C.Structure):
    _fields_ = [
        ("num_classes", C.c_int32), ("strategy", C.c_int32),
        ("w0", C.c_float), ("w1", C.c_float),
        ("temperature", C.c_float), ("use_temperature", C.c_int32),
        ("label_dtype", C.c_int32), ("ignore_index", C.c_int32),
        ("ece_bins", C.c_int32), ("auroc_bins", C.c_int32),
        ("auroc_hi", C.c_float),
        ("ece_edges", C.c_float * (MAX_ECE_BINS + 1)),
    ]


class ScoreMaps(C.Structure):
    _fields_ = [
        ("pred", C.c_void_p), ("pred_dtype", C.c_int32), ("reserved", C.c_int32),
        ("fused", C.c_void_p), ("conf", C.c_void_p), ("mi", C.c_void_p), ("js", C.c_void_p),
    ]


class BinsLayout(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "confusion", "ece_count", "ece_correct", "ece_conf_hi", "ece_conf_lo",
        "auroc_pos", "auroc_neg", "counters", "total_words")]


# numpy mirror of AwxCorruptParams (one record per image, uploaded as raw bytes)
CORRUPT_PARAMS_DTYPE = np.dtype([
    ("kind", np.int32), ("blur_k", np.int32),
    ("d0", np.float64), ("d1", np.float64),
    ("f0", np.float32), ("f1", np.float32),
    ("taps", np.float32, (4,)),
    ("field_offset", np.int64),
    ("item_begin", np.int32), ("item_count", np.int32),
], align=True)

_SIGNATURES = {
    "awx_version": (C.c_int, []),
    "awx_last_error": (C.c_char_p, []),
    "awx_launch_count": (C.c_int64, []),
    "awx_device_info": (C.c_int, [C.POINTER(C.c_int)] * 3),
    "awx_bins_layout": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, C.POINTER(BinsLayout)]),
    "awx_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                            C.POINTER(ScoreConfig), C.c_void_p, C.POINTER(ScoreMaps), C.c_void_p]),
    "awx_member_variance": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_void_p]),
    "awx_members_n": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_int32,
                                C.c_int32, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                C.c_void_p]),
    "awx_confusion": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32,
                                C.c_void_p, C.c_void_p]),
    "awx_corrupt_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int32, C.c_int32]),
    "awx_corrupt": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                              C.c_int32, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "awx_corrupt_normalized": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64,
                                         C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64,
                                         C.c_void_p, C.c_void_p]),
    "awx_corrupt_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                    C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                    C.POINTER(ScoreConfig), C.c_void_p, C.POINTER(ScoreMaps), C.c_void_p]),
    "awx_synth_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_double, C.c_void_p, C.c_int32, C.c_void_p]),
    "awx_fogloss": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                              C.c_int32, C.c_int64, C.c_int32, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "awx_fogloss_workspace_bytes": (C.c_size_t, []),
    "awx_fuse_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64, C.c_int32, C.c_float,
                                   C.c_float, C.c_float, C.c_int32, C.c_void_p]),
    "awx_fuse_backward_workspace_bytes": (C.c_size_t, []),
    "awx_fuse_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int64,
                                    C.c_int32, C.c_float, C.c_float, C.c_float, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "awx_depth_density_workspace_bytes": (C.c_size_t, []),
    "awx_depth_density_fwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]),
    "awx_depth_density_bwd": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p]),
    "awx_scale_inplace": (C.c_int, [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p]),
    "awx_normalize_chw": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                    C.c_void_p, C.c_void_p]),
    "awx_style_transfer": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_float, C.c_float, C.c_double, C.c_int32,
                                     C.c_void_p]),
    "awx_temperature_workspace_bytes": (C.c_size_t, [C.c_int32]),
    "awx_temperature_nll": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                      C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]),
    "awx_fog_density_workspace_bytes": (C.c_size_t, [C.c_int64]),
    "awx_local_contrast": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int64,
                                     C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]),
    "awx_fog_density_finish": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                         C.c_void_p, C.c_void_p]),
    "awx_estimate_depth": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_void_p,
                                     C.c_int32, C.c_void_p, C.c_void_p]),
}

_lock = threading.Lock()
_lib = None


def exported_symbols() -> list:
    """Every entry point include/awx.h declares (checked by the CPU test-suite)."""
    return sorted(_SIGNATURES)


def _ensure_built() -> str:
    path = _build.LIB_PATH
    if _build.is_current():
        return path
    # stale or missing: rebuild under an inter-process lock (several ranks may import at once)
    lock_path = os.path.join(_build.PKG_DIR, ".build.lock")
    with open(lock_path, "w") as lf:
        fcntl.flock(lf, fcntl.LOCK_EX)
        try:
            if not _build.is_current():
                _build.build()
        finally:
            fcntl.flock(lf, fcntl.LOCK_UN)
    return path


def load():
    """Load (building first if needed) and return the ctypes handle; raises RuntimeError on failure."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        try:
            # dev knob: AWX_LIB points at an experimental build of csrc/ (tools/build_variants.py)
            path = os.environ.get("AWX_LIB") or _ensure_built()
            lib = C.CDLL(path)
        except Exception as exc:  # no fallback: fail loudly
            raise RuntimeError(
                "libawx.so (the CUDA implementation of this package) is not available and could not be "
                f"built: {exc}. Run `python -m adverse_weather_semantic_segmentation_robustness_benchmark_b200.build`."
            ) from exc
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
        return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().awx_last_error().decode("utf-8", "replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def bins_layout(num_classes: int, ece_bins: int, auroc_bins: int) -> BinsLayout:
    lay = BinsLayout()
    check(load().awx_bins_layout(num_classes, ece_bins, auroc_bins, C.byref(lay)), "awx_bins_layout")
    return lay
