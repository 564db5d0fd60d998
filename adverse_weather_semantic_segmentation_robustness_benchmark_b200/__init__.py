"""B200-native implementation of the per-pixel robustness-evaluation hot path of
A-SHOJAEI/adverse-weather-semantic-segmentation-robustness-benchmark.

Same public class names as the reference package (src/.../__init__.py:14-41) for the classes on
the hot path; the arithmetic runs in libawx.so (hand-written sm_100a CUDA behind the C ABI in
include/awx.h).  Importing the package does not need a GPU; calling into it does.
"""

__version__ = "0.1.0"

from .evaluation.metrics import (  # noqa: F401
    IoUMetrics,
    ConfidenceCalibration,
    EnsembleDisagreementMetrics,
    RobustnessMetrics,
)

from .data.preprocessing import WeatherDegradationTransforms  # noqa: F401,E402
from .models.model import EnsembleModel, FogDensityAwareLoss  # noqa: F401,E402

__all__ = [
    "WeatherDegradationTransforms",
    "EnsembleModel",
    "FogDensityAwareLoss",
    "IoUMetrics",
    "ConfidenceCalibration",
    "EnsembleDisagreementMetrics",
    "RobustnessMetrics",
]
