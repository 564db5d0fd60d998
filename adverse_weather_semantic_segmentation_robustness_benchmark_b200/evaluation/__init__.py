"""Evaluation metrics (mirror of the reference's ``evaluation`` package, evaluation/__init__.py:3-14)."""

from .metrics import (
    IoUMetrics,
    ConfidenceCalibration,
    EnsembleDisagreementMetrics,
    RobustnessMetrics,
)

from .streaming import StreamingEvaluator, evaluate_model
from .validation import validate_epoch, estimate_fog_density

__all__ = [
    "StreamingEvaluator",
    "evaluate_model",
    "validate_epoch",
    "estimate_fog_density",
    "IoUMetrics",
    "ConfidenceCalibration",
    "EnsembleDisagreementMetrics",
    "RobustnessMetrics",
]
