"""Data-side pieces on the hot path (mirror of the reference's ``data`` package for those names)."""

from .preprocessing import WeatherDegradationTransforms, WeatherDraw, DepthEstimationPreprocessor
from .loader import WeatherAugmentationPipeline, normalize_to_tensor

__all__ = ["WeatherDegradationTransforms", "WeatherDraw", "DepthEstimationPreprocessor", "WeatherAugmentationPipeline",
           "normalize_to_tensor"]
