"""Evaluation metrics (mirror of the reference's ``evaluation`` package, evaluation/__init__.py:3-14)."""

from .metrics import (
    IoUMetrics,
    ConfidenceCalibration,
    EnsembleDisagreementMetrics,
    RobustnessMetrics,
)

__all__ = [
    "IoUMetrics",
    "ConfidenceCalibration",
    "EnsembleDisagreementMetrics",
    "RobustnessMetrics",
]
