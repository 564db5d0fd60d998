"""Import shim for the LIVE reference (only available in the build container).

Used by tests/golden/make_golden.py and tests/test_oracle_live.py.  Never used on
the GPU box (``/root/reference`` does not exist there) and never by the product.
Recipe: SURVEY.md Appendix A.
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("AWX_REFERENCE_ROOT", "/root/reference")
PKG = "adverse_weather_semantic_segmentation_robustness_benchmark"
_SRC = os.path.join(REF_ROOT, "src")
_PKG_DIR = os.path.join(_SRC, PKG)


def available() -> bool:
    return os.path.isdir(_PKG_DIR)


def _load_by_path(name: str, rel: str):
    spec = importlib.util.spec_from_file_location(name, os.path.join(_PKG_DIR, rel))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


_cache: dict = {}


def preprocessing():
    """data/preprocessing.py loaded by file path (its package __init__ needs albumentations)."""
    if "pre" not in _cache:
        _cache["pre"] = _load_by_path("_awx_ref_preprocessing", "data/preprocessing.py")
    return _cache["pre"]


def metrics():
    if "met" not in _cache:
        _cache["met"] = _load_by_path("_awx_ref_metrics", "evaluation/metrics.py")
    return _cache["met"]


def model():
    """models/model.py with segmentation_models_pytorch stubbed (not installed here)."""
    if "mod" not in _cache:
        if "segmentation_models_pytorch" not in sys.modules:
            sys.modules["segmentation_models_pytorch"] = types.ModuleType("segmentation_models_pytorch")
        _cache["mod"] = _load_by_path("_awx_ref_model", "models/model.py")
    return _cache["mod"]


def loader():
    """data/loader.py imported through the reference package with albumentations stubbed (not installed
    here; only its names are needed at import time -- SURVEY.md Appendix A.2)."""
    if "loader" not in _cache:
        if "albumentations" not in sys.modules:
            alb = types.ModuleType("albumentations")
            for name in ("Compose", "HorizontalFlip", "RandomBrightnessContrast", "Normalize"):
                setattr(alb, name, type(name, (), {"__init__": lambda self, *a, **k: None}))
            albp = types.ModuleType("albumentations.pytorch")
            albp.ToTensorV2 = type("ToTensorV2", (), {"__init__": lambda self, *a, **k: None})
            alb.pytorch = albp
            sys.modules["albumentations"] = alb
            sys.modules["albumentations.pytorch"] = albp
        if _SRC not in sys.path:
            sys.path.insert(0, _SRC)
        import importlib
        _cache["loader"] = importlib.import_module(PKG + ".data.loader")
    return _cache["loader"]


def trainer():
    """training/trainer.py imported through the reference package (needs the loader and model shims)."""
    if "trainer" not in _cache:
        loader()
        if "segmentation_models_pytorch" not in sys.modules:
            sys.modules["segmentation_models_pytorch"] = types.ModuleType("segmentation_models_pytorch")
        import importlib
        _cache["trainer"] = importlib.import_module(PKG + ".training.trainer")
    return _cache["trainer"]


def evaluate_script():
    """scripts/evaluate.py (the reference's evaluation driver) loaded by path.  Its plotting imports (matplotlib,
    seaborn: not installed here, used only by generate_evaluation_report) are stubbed; everything
    ``evaluate_model`` touches is the reference's own code."""
    if "evaluate" not in _cache:
        loader()
        model()
        for name in ("matplotlib", "matplotlib.pyplot", "seaborn"):
            if name not in sys.modules:
                try:
                    __import__(name)
                except Exception:
                    sys.modules[name] = types.ModuleType(name)
        if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
            sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
        if "segmentation_models_pytorch" not in sys.modules:
            sys.modules["segmentation_models_pytorch"] = types.ModuleType("segmentation_models_pytorch")
        spec = importlib.util.spec_from_file_location("_awx_ref_evaluate", os.path.join(REF_ROOT, "scripts", "evaluate.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _cache["evaluate"] = mod
    return _cache["evaluate"]


def ensemble_with_fixed_members(l1, l2, strategy="weighted_average", temperature_scaling=True,
                                raw_weights=None, temperature=None, d1=None, d2=None):
    """The reference's EnsembleModel with two tiny producers injected, so that its own
    forward()/get_ensemble_disagreement() code runs unmodified on chosen logits."""
    import torch
    from torch import nn

    m = model()

    class _Fixed(nn.Module):
        def __init__(self, seg, depth):
            super().__init__()
            self.seg, self.depth = seg, depth

        def forward(self, x):
            out = {"segmentation": self.seg}
            if self.depth is not None:
                out["depth"] = self.depth
            return out

    ens = m.EnsembleModel.__new__(m.EnsembleModel)
    nn.Module.__init__(ens)
    ens.num_classes = l1.shape[1]
    ens.include_depth = d1 is not None
    ens.ensemble_strategy = strategy
    ens.temperature_scaling = temperature_scaling
    ens.segformer = _Fixed(l1, d1)
    ens.deeplabv3plus = _Fixed(l2, d2)
    ens.ensemble_weights = nn.Parameter(torch.ones(2) / 2 if raw_weights is None else raw_weights.clone())
    if temperature_scaling:
        ens.temperature = nn.Parameter(torch.ones(1) if temperature is None else temperature.clone())
    return ens
