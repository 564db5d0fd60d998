"""Data-side pieces on the hot path (mirror of the reference's ``data`` package for those names)."""

from .preprocessing import WeatherDegradationTransforms, WeatherDraw

__all__ = ["WeatherDegradationTransforms", "WeatherDraw"]
