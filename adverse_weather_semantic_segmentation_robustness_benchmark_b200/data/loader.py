"""Drop-in pieces of the reference's ``data/loader.py`` that sit either side of the corruption kernels:
``WeatherAugmentationPipeline`` (loader.py:300-387) and the Normalize + ToTensorV2 epilogue of the dataset's
transform pipeline (loader.py:196-199).  Stochastic choices stay on the host with the reference's RNG calls;
the per-pixel arithmetic runs in libawx.so."""

from __future__ import annotations

import logging
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from .. import ops_prep
from .preprocessing import WeatherDegradationTransforms

logger = logging.getLogger(__name__)


def normalize_to_tensor(images, mean: Sequence[float] = ops_prep.IMAGENET_MEAN, std: Sequence[float] = ops_prep.IMAGENET_STD,
                        out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """Normalize(mean, std) + ToTensorV2 of the reference's transform pipeline (loader.py:196-199) for a
    uint8 [B,H,W,3] (or [H,W,3]) batch: fp32 / bf16 [B,3,H,W] on the device, ready for the backbones."""
    img = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
    single = img.dim() == 3
    if single:
        img = img[None]
    out = ops_prep.normalize_chw(img, mean, std, out_dtype=out_dtype)
    return out[0] if single else out


class WeatherAugmentationPipeline:
    """Advanced weather augmentation pipeline for domain adaptation (reference: loader.py:300-387)."""

    def __init__(self, weather_intensities: Dict[str, float] = None, style_transfer_prob: float = 0.3, **kwargs) -> None:
        self.weather_intensities = weather_intensities or {"fog": 0.7, "rain": 0.5, "snow": 0.6, "night": 0.8}
        self.style_transfer_prob = style_transfer_prob
        self.weather_transforms = WeatherDegradationTransforms()
        logger.info("Initialized WeatherAugmentationPipeline")

    def apply_domain_adaptation_augmentation(self, image: np.ndarray, target_weather: str = None) -> np.ndarray:
        """loader.py:331-360: weather degradation at the configured intensity, then (with probability
        ``style_transfer_prob``) the colour-space style transfer.  RNG calls in the reference's order."""
        if target_weather is None:
            target_weather = np.random.choice(list(self.weather_intensities.keys()))
        augmented = self.weather_transforms.apply_weather_effect(
            image, target_weather, intensity=self.weather_intensities[target_weather])
        if np.random.random() < self.style_transfer_prob:
            augmented = self._apply_style_transfer(augmented, target_weather)
        return augmented

    def _apply_style_transfer(self, image: np.ndarray, weather_type: str) -> np.ndarray:
        """loader.py:362-387: cv2.convertScaleAbs(alpha, beta) and the blue-channel gain (awx_style_transfer)."""
        if weather_type not in ops_prep.STYLE:
            return image
        if image.dtype != np.uint8:
            raise TypeError("libawx restyles uint8 frames; got dtype %s" % image.dtype)
        out = ops_prep.style_transfer(torch.from_numpy(np.ascontiguousarray(image)), weather_type)
        return out.cpu().numpy()
