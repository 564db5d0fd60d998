"""Model-side pieces on the hot path (mirror of the reference's ``models`` package for those names)."""

from .model import EnsembleModel, FogDensityAwareLoss

__all__ = ["EnsembleModel", "FogDensityAwareLoss"]
