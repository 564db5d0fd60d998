"""Loader and driver of ``oracle/_ref/`` -- the reference's OWN hot-path code (byte-compiled by oracle/build_ref.py).

TEST / BASELINE INFRASTRUCTURE (see oracle/__init__.py): used by ``bench.py``'s ``cpu_baseline`` leg and
``--impl reference`` arm, and by tests.  Nothing here is imported by the product package.

``reference_step`` runs, per weather condition and per frame, exactly what the reference's evaluation does on the
CPU: ``WeatherDegradationTransforms.apply_weather_effect`` (data/preprocessing.py:61-92), ``EnsembleModel.forward``
with two logit-replaying members injected (models/model.py:428-486: its own fusion code, unmodified), and
``RobustnessMetrics.compute_comprehensive_metrics`` (evaluation/metrics.py:565-605: IoU, pixel accuracy, ECE,
disagreement AUROC).
"""

from __future__ import annotations

import importlib.machinery
import importlib.util
import json
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
EXT = ".refbin"
_cache: dict = {}


def available() -> bool:
    """True when oracle/_ref holds the three byte-compiled modules for THIS interpreter."""
    try:
        with open(os.path.join(REF_DIR, "MANIFEST.json")) as fh:
            meta = json.load(fh)
    except Exception:
        return False
    if meta.get("magic") != importlib.util.MAGIC_NUMBER.hex():
        return False
    return all(os.path.exists(os.path.join(REF_DIR, n + EXT)) for n in ("preprocessing", "model", "metrics"))


def module(name: str):
    """One of 'preprocessing', 'model', 'metrics' (the reference's module object)."""
    if name not in _cache:
        if name == "model" and "segmentation_models_pytorch" not in sys.modules:
            # not installed in this image; only needed by the backbones, which the benchmark replaces by injected
            # logit producers
            sys.modules["segmentation_models_pytorch"] = types.ModuleType("segmentation_models_pytorch")
        path = os.path.join(REF_DIR, name + EXT)
        loader = importlib.machinery.SourcelessFileLoader("_awx_ref_" + name, path)
        spec = importlib.util.spec_from_loader("_awx_ref_" + name, loader)
        mod = importlib.util.module_from_spec(spec)
        loader.exec_module(mod)
        _cache[name] = mod
    return _cache[name]


def ensemble(l1, l2, strategy="weighted_average", raw_weights=(0.5, 0.5), temperature=1.0):
    """The reference's EnsembleModel around two modules that return the given logits: its
    __init__ would download backbones, so the object is assembled by hand; forward() is the reference's own."""
    import torch
    from torch import nn
    m = module("model")

    class _Fixed(nn.Module):
        def __init__(self, seg):
            super().__init__()
            self.seg = seg

        def forward(self, x):
            return {"segmentation": self.seg}

    ens = m.EnsembleModel.__new__(m.EnsembleModel)
    nn.Module.__init__(ens)
    ens.num_classes = l1.shape[1]
    ens.include_depth = False
    ens.ensemble_strategy = strategy
    ens.temperature_scaling = temperature is not None
    ens.segformer, ens.deeplabv3plus = _Fixed(l1), _Fixed(l2)
    ens.ensemble_weights = nn.Parameter(torch.tensor(list(raw_weights), dtype=torch.float32))
    if temperature is not None:
        ens.temperature = nn.Parameter(torch.tensor([float(temperature)]))
    return ens.eval()


def reference_step(frames, conditions, raw_weights, temperature, num_classes=19, seed=42) -> float:
    """`frames`: list of (image uint8 [H,W,3] ndarray, labels [1,H,W] tensor, logits_a, logits_b [1,C,H,W]).
    Every frame goes through every condition: corrupt -> fuse -> comprehensive metrics.  Returns the pixels scored."""
    import torch
    pre, met = module("preprocessing"), module("metrics")
    transforms = pre.WeatherDegradationTransforms(seed=seed)
    robust = met.RobustnessMetrics(num_classes=num_classes)
    pixels = 0.0
    with torch.no_grad():
        for kind in conditions:
            for img, lab, la, lb in frames:
                transforms.apply_weather_effect(img, kind)
                fused = ensemble(la, lb, "weighted_average", raw_weights, temperature)(None)["segmentation"]
                robust.compute_comprehensive_metrics(fused, lab, [la, lb], kind)
                pixels += float(img.shape[0] * img.shape[1])
    return pixels
