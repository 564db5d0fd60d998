"""Oracle: the producers / consumers either side of the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, under
``/root/reference/src/adverse_weather_semantic_segmentation_robustness_benchmark/`` (``P/``):

* get_fog_density_map                         P/data/preprocessing.py:250-288
* DepthEstimationPreprocessor.estimate_depth  P/data/preprocessing.py:304-367
* WeatherAugmentationPipeline._apply_style_transfer   P/data/loader.py:364-387
* Normalize(mean, std) + ToTensorV2           P/data/loader.py:196-199 (albumentations; see below)
* ConfidenceCalibration.optimize_temperature  P/evaluation/metrics.py:283-321
* Trainer._estimate_fog_density               P/training/trainer.py:480-511
* Trainer.validate_epoch                      P/training/trainer.py:377-478

OpenCV / SciPy / NumPy / torch arithmetic is CALLED as the reference calls it.  ``normalize_chw``
is the exception: albumentations is not installed in the build container (and not vendored in the
reference), so its published algorithm (albumentations.augmentations.functional.normalize, 1.x:
``img.astype(float32); img -= mean*max_pixel_value; img *= reciprocal(std*max_pixel_value)``, then
ToTensorV2's HWC->CHW transpose) is restated -- PARITY UNPINNED for that one function.
"""

from __future__ import annotations

import numpy as np
import cv2
import torch
import torch.nn.functional as F
from scipy.ndimage import gaussian_filter

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


# ------------------------------------------------------------------------ fog density map
def fog_density_map(image: np.ndarray, depth: np.ndarray) -> np.ndarray:
    """image: float HWC in [0,1]; depth: [H,W] (fp64 from the synthetic-depth generator).  :265-288"""
    gray = cv2.cvtColor((image * 255).astype(np.uint8), cv2.COLOR_RGB2GRAY)
    gray = gray.astype(np.float32) / 255.0
    kernel = np.ones((5, 5), np.float32) / 25
    local_mean = cv2.filter2D(gray, -1, kernel)
    local_variance = cv2.filter2D((gray - local_mean) ** 2, -1, kernel)
    local_contrast = np.sqrt(local_variance)
    max_contrast = np.percentile(local_contrast, 95)
    fog_density = 1.0 - (local_contrast / (max_contrast + 1e-8))
    normalized_depth = depth / np.max(depth)
    fog_density = fog_density * (0.3 + 0.7 * normalized_depth)
    return np.clip(fog_density, 0, 1)


def local_contrast(image: np.ndarray) -> np.ndarray:
    """The intermediate of fog_density_map before the global percentile (fp32 [H,W]).  uint8 frames are
    taken as they are (an extension: the reference's own conversion only makes sense for floats)."""
    u8 = image if image.dtype == np.uint8 else (image * 255).astype(np.uint8)
    gray = cv2.cvtColor(u8, cv2.COLOR_RGB2GRAY)
    gray = gray.astype(np.float32) / 255.0
    kernel = np.ones((5, 5), np.float32) / 25
    local_mean = cv2.filter2D(gray, -1, kernel)
    return np.sqrt(cv2.filter2D((gray - local_mean) ** 2, -1, kernel))


# ------------------------------------------------------------------------ depth estimation
def estimate_depth(image_u8: np.ndarray) -> np.ndarray:
    """uint8 RGB HWC -> fp64 [H,W] in [0,1].  :332-367"""
    h, w = image_u8.shape[:2]
    gray = cv2.cvtColor(image_u8, cv2.COLOR_RGB2GRAY)
    sky_mask = np.zeros((h, w), dtype=np.float32)
    sky_mask[:h // 3, :] = 1.0
    road_mask = np.zeros((h, w), dtype=np.float32)
    road_mask[h // 2:, :] = 1.0
    y_coords = np.arange(h)[:, np.newaxis] / h
    depth = np.tile(y_coords * 0.8 + 0.2, (1, w))
    depth[sky_mask > 0] = 1.0
    depth[road_mask > 0] *= 0.5
    texture = cv2.Laplacian(gray, cv2.CV_64F)
    texture_strength = np.abs(texture) / (np.max(np.abs(texture)) + 1e-8)
    depth = np.clip(depth + (-0.3 * texture_strength), 0, 1)
    return gaussian_filter(depth, sigma=2)


# -------------------------------------------------------------------------- style transfer
STYLE = {"fog": (0.8, 30, None), "rain": (1.2, -10, 1.1), "snow": (0.9, 20, None), "night": (0.4, -20, 1.3)}


def style_transfer(image_u8: np.ndarray, weather_type: str) -> np.ndarray:
    """cv2.convertScaleAbs(alpha, beta) (+ blue-channel gain, truncated on assignment).  :364-385"""
    if weather_type not in STYLE:
        return image_u8
    alpha, beta, gain = STYLE[weather_type]
    image = cv2.convertScaleAbs(image_u8, alpha=alpha, beta=beta)
    if gain is not None:
        image[:, :, 2] = np.clip(image[:, :, 2] * gain, 0, 255)
    return image


# ------------------------------------------------------------------ Normalize + ToTensorV2
def normalize_chw(image_u8: np.ndarray, mean=IMAGENET_MEAN, std=IMAGENET_STD, max_pixel_value: float = 255.0) -> np.ndarray:
    """uint8 HWC -> fp32 CHW (published albumentations algorithm; parity unpinned, see header)."""
    m = np.array(mean, dtype=np.float32)
    m *= max_pixel_value
    s = np.array(std, dtype=np.float32)
    s *= max_pixel_value
    denominator = np.reciprocal(s, dtype=np.float32)
    img = image_u8.astype(np.float32)
    img -= m
    img *= denominator
    return np.ascontiguousarray(img.transpose(2, 0, 1))


# ------------------------------------------------------------------ temperature grid search
def temperature_grid() -> torch.Tensor:
    return torch.linspace(0.1, 10.0, 100)


def temperature_nll(logits: torch.Tensor, targets: torch.Tensor) -> torch.Tensor:
    """The 100 NLL values optimize_temperature compares (fp32 tensor).  :302-316"""
    logits_flat = logits.view(-1, logits.size(1))
    targets_flat = targets.view(-1)
    valid = targets_flat != 255
    logits_flat = logits_flat[valid]
    targets_flat = targets_flat[valid]
    return torch.stack([F.cross_entropy(logits_flat / t, targets_flat) for t in temperature_grid()])


def optimize_temperature(logits: torch.Tensor, targets: torch.Tensor) -> float:
    """:283-321 (note: logits.view(-1, C) of an NCHW tensor is the reference's own flattening)."""
    best_t, best = 1.0, float("inf")
    for t, nll in zip(temperature_grid(), temperature_nll(logits, targets)):
        if nll < best:
            best, best_t = nll, t.item()
    return best_t


# ------------------------------------------------------------------ trainer fog-density maps
FOG_DENSITY_AFFINE = {"fog": (0.5, 0.5), "rain": (0.3, 0.2), "snow": (0.3, 0.2)}


def estimate_fog_density(weather_conditions, h: int, w: int) -> torch.Tensor:
    """torch.rand(h, w) * a + b per frame from the global torch CPU generator.  :497-509"""
    out = torch.zeros(len(weather_conditions), h, w)
    for i, weather in enumerate(weather_conditions):
        a, b = FOG_DENSITY_AFFINE.get(weather, (0.1, None))
        r = torch.rand(h, w)
        out[i] = r * a + b if b is not None else r * a
    return out


def validate_epoch(model, val_loader, loss_fn, num_classes: int = 19, fog_aware: bool = True) -> dict:
    """The trainer's validation pass (:377-478) on the CPU: per batch the loss (``loss_fn(outputs, targets,
    fog_density) -> dict`` with the maps of ``estimate_fog_density`` when ``fog_aware``, else
    ``loss_fn(logits, labels) -> scalar``), sums of ``loss.item() * batch_size`` in Python floats; at the end
    mIoU of the concatenated argmax maps, overall and for the five fixed weather names (:393-394, :466-476)."""
    from . import metrics as om
    sums = {"val_loss": 0.0, "val_seg_loss": 0.0, "val_depth_loss": 0.0, "val_samples": 0}
    fixed = ("clean", "fog", "rain", "snow", "night")
    preds, tgts = [], []
    by_weather = {k: ([], []) for k in fixed}
    for batch in val_loader:
        images, labels = batch["image"], batch["label"]
        weather = batch.get("weather_condition", ["clean"] * images.size(0))
        outputs = model(images)
        targets = {"label": labels}
        if batch.get("depth") is not None:
            targets["depth"] = batch["depth"]
        if fog_aware:
            fd = estimate_fog_density(weather, images.shape[2], images.shape[3]) if len(weather) else None
            ld = loss_fn(outputs, targets, fd)
            loss, seg, dep = ld["total_loss"], ld["segmentation_loss"], ld["depth_loss"]
        else:
            seg = loss_fn(outputs["segmentation"], labels)
            loss, dep = seg, 0.0
        n = images.size(0)
        sums["val_loss"] += loss.item() * n
        sums["val_seg_loss"] += seg.item() * n
        sums["val_depth_loss"] += (dep.item() if isinstance(dep, torch.Tensor) else dep) * n
        sums["val_samples"] += n
        p = outputs["segmentation"].argmax(dim=1)
        preds.append(p)
        tgts.append(labels)
        for i, wname in enumerate(weather):
            if wname in by_weather:
                by_weather[wname][0].append(p[i:i + 1])
                by_weather[wname][1].append(labels[i:i + 1])
    for k in ("val_loss", "val_seg_loss", "val_depth_loss"):
        sums[k] /= sums["val_samples"]
    sums["val_miou"] = om.iou(torch.cat(preds), torch.cat(tgts), num_classes)["mean_iou"]
    for wname, (pl, tl) in by_weather.items():
        if pl:
            sums[f"val_miou_{wname}"] = om.iou(torch.cat(pl), torch.cat(tl), num_classes)["mean_iou"]
    return sums


# ------------------------------------------------------------ domain-adaptation augmentation
DEFAULT_INTENSITIES = {"fog": 0.7, "rain": 0.5, "snow": 0.6, "night": 0.8}


def domain_adaptation_augmentation(image_u8: np.ndarray, intensities=None, style_transfer_prob: float = 0.3,
                                   target_weather=None) -> np.ndarray:
    """P/data/loader.py:331-360 with the reference's RNG order: weather choice, the corruption's own
    draws, then the style-transfer coin."""
    from . import weather as ow
    intensities = intensities or DEFAULT_INTENSITIES
    if target_weather is None:
        target_weather = np.random.choice(list(intensities.keys()))
    out = ow.apply(image_u8, target_weather, intensities[target_weather])
    if np.random.random() < style_transfer_prob:
        out = style_transfer(out, target_weather)
    return out
