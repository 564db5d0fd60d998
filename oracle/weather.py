"""Oracle: weather corruptions (fog / rain / snow / night) as draw + apply pairs.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
``/root/reference/src/adverse_weather_semantic_segmentation_robustness_benchmark/
data/preprocessing.py`` (abbreviated ``P/data/preprocessing.py`` below):

* dispatcher                     P/data/preprocessing.py:61-92
* fog                            :94-123   (+ synthetic depth :227-248)
* rain                           :125-168
* snow                           :170-202
* night                          :204-225

Every ``draw_*`` consumes the process-global legacy NumPy RNG in exactly the
order the reference does, so ``np.random.seed(s); apply(...)`` here and
``np.random.seed(s); ref.apply_weather_effect(...)`` see the same stream.
Every ``*_apply`` is deterministic given the drawn parameters.
"""

from __future__ import annotations

import numpy as np
import cv2
from scipy.ndimage import gaussian_filter

KINDS = ("clean", "fog", "rain", "snow", "night")

# Parameter tables, P/data/preprocessing.py:33-57
FOG_BETA = (0.005, 0.05)
FOG_A = (0.7, 1.0)
FOG_DEPTH_SCALE = 100.0
RAIN_DROPS = (100, 500)
RAIN_THICKNESS = (1, 3)
RAIN_ANGLE = (-15, 15)
RAIN_COLOR = (0.8, 0.9, 1.0)
SNOW_FLAKES = (50, 200)
SNOW_RADII = (2, 8)
SNOW_BLUR = (3, 7)
NIGHT_REDUCTION = (0.2, 0.6)
NIGHT_SHIFT = (0.8, 0.85, 1.2)
NIGHT_NOISE_STD = 5.0


def to_unit_float(image_u8: np.ndarray) -> np.ndarray:
    """uint8 HWC -> float32 in [0,1]  (P/data/preprocessing.py:81)."""
    return image_u8.astype(np.float32) / 255.0


def _to_u8(x: np.ndarray) -> np.ndarray:
    """clip to [0,1], scale by 255, TRUNCATE to uint8 (:123,:168,:202,:225)."""
    return (np.clip(x, 0, 1) * 255).astype(np.uint8)


# --------------------------------------------------------------------------- fog
def depth_from_noise(noise: np.ndarray) -> np.ndarray:
    """Synthetic depth: vertical ramp + noise, Gaussian sigma=2, floor 1.0 (:235-246)."""
    h = noise.shape[0]
    ramp = (np.arange(h)[:, np.newaxis] / h) * FOG_DEPTH_SCALE
    smooth = gaussian_filter(ramp + noise, sigma=2)
    return np.maximum(smooth, 1.0)


def draw_fog(h: int, w: int, intensity=None) -> dict:
    """RNG order: depth noise first (:104 -> :239), then the intensity (:107-108)."""
    noise = np.random.normal(0, 10, (h, w))
    if intensity is None:
        intensity = np.random.uniform(0.3, 0.9)
    return {"noise": noise, "intensity": intensity}


def fog_coefficients(intensity) -> tuple:
    """beta, A as the reference derives them (:110-114)."""
    beta = FOG_BETA[0] + intensity * (FOG_BETA[1] - FOG_BETA[0])
    airlight = FOG_A[0] + intensity * (FOG_A[1] - FOG_A[0])
    return beta, airlight


def fog_apply(image_u8: np.ndarray, depth: np.ndarray, intensity) -> np.ndarray:
    """Koschmieder blend in float64, truncated to uint8 (:116-123)."""
    img = to_unit_float(image_u8)
    beta, airlight = fog_coefficients(intensity)
    t = np.exp(-beta * depth)[..., np.newaxis]
    veil = airlight * np.ones_like(img)
    return _to_u8(img * t + veil * (1 - t))


# -------------------------------------------------------------------------- rain
def rain_drop_count(intensity) -> int:
    return int(RAIN_DROPS[0] + intensity * (RAIN_DROPS[1] - RAIN_DROPS[0]))


def draw_rain(h: int, w: int, intensity=None) -> dict:
    """Per drop: x, y, length, thickness, angle -> clipped end point (:144-156).

    Returns ``drops`` int32 [n,5] = (x0, y0, x1, y1, thickness).
    """
    if intensity is None:
        intensity = np.random.uniform(0.2, 0.8)
    n = rain_drop_count(intensity)
    drops = np.zeros((n, 5), dtype=np.int32)
    for i in range(n):
        x = np.random.randint(0, w)
        y = np.random.randint(0, h)
        length = np.random.randint(5, 20)
        thickness = np.random.choice(RAIN_THICKNESS)
        angle = np.random.uniform(*RAIN_ANGLE)
        ex = int(x + length * np.sin(np.radians(angle)))
        ey = int(y + length * np.cos(np.radians(angle)))
        ex = np.clip(ex, 0, w - 1)
        ey = np.clip(ey, 0, h - 1)
        drops[i] = (x, y, ex, ey, thickness)
    return {"intensity": intensity, "drops": drops}


def rain_apply(image_u8: np.ndarray, intensity, drops: np.ndarray) -> np.ndarray:
    """Haze, streaks (cv2.line), 3x3 Gaussian sigma 0.5, truncate (:133-168)."""
    img = to_unit_float(image_u8)
    haze = intensity * 0.3
    canvas = img * (1 - haze) + haze * 0.7
    for x0, y0, x1, y1, th in np.asarray(drops).tolist():
        cv2.line(canvas, (x0, y0), (x1, y1), list(RAIN_COLOR), th)
    canvas = cv2.GaussianBlur(canvas, (3, 3), 0.5)
    return _to_u8(canvas)


# -------------------------------------------------------------------------- snow
def snow_flake_count(intensity) -> int:
    return int(SNOW_FLAKES[0] + intensity * (SNOW_FLAKES[1] - SNOW_FLAKES[0]))


def draw_snow(h: int, w: int, intensity=None) -> dict:
    """Per flake x, y, radius; then the blur size (:189-199).

    Returns ``flakes`` int32 [n,3] = (x, y, r) and odd ``blur_k``.
    """
    if intensity is None:
        intensity = np.random.uniform(0.2, 0.7)
    n = snow_flake_count(intensity)
    flakes = np.zeros((n, 3), dtype=np.int32)
    for i in range(n):
        x = np.random.randint(0, w)
        y = np.random.randint(0, h)
        r = np.random.choice(SNOW_RADII)
        flakes[i] = (x, y, r)
    k = int(np.random.choice(SNOW_BLUR))
    if k % 2 == 0:
        k += 1
    return {"intensity": intensity, "flakes": flakes, "blur_k": k}


def snow_apply(image_u8: np.ndarray, intensity, flakes: np.ndarray, blur_k: int) -> np.ndarray:
    """Brightness boost, filled white discs, k x k Gaussian sigma 1, truncate (:178-202)."""
    img = to_unit_float(image_u8)
    canvas = np.clip(img + intensity * 0.2, 0, 1)
    for x, y, r in np.asarray(flakes).tolist():
        cv2.circle(canvas, (x, y), r, (1.0, 1.0, 1.0), -1)
    canvas = cv2.GaussianBlur(canvas, (int(blur_k), int(blur_k)), 1.0)
    return _to_u8(canvas)


# ------------------------------------------------------------------------- night
def draw_night(shape: tuple, intensity=None) -> dict:
    """RNG order: intensity, brightness-reduction draw, noise field (:207,:212,:222)."""
    if intensity is None:
        intensity = np.random.uniform(0.4, 0.8)
    reduction = np.random.uniform(*NIGHT_REDUCTION)
    noise = np.random.normal(0, NIGHT_NOISE_STD / 255.0, shape)
    return {"intensity": intensity, "reduction": reduction, "noise": noise}


def night_apply(image_u8: np.ndarray, intensity, reduction, noise: np.ndarray) -> np.ndarray:
    """Dim (fp32), per-channel shift (fp32, in place), add fp64 noise, truncate (:209-225)."""
    img = to_unit_float(image_u8)
    dimmed = img * (1 - intensity * reduction)
    for c, s in enumerate(NIGHT_SHIFT):
        dimmed[:, :, c] *= s
    return _to_u8(dimmed + noise * intensity * 0.5)


# -------------------------------------------------------------------- dispatcher
def apply(image_u8: np.ndarray, kind: str, intensity=None) -> np.ndarray:
    """draw + apply with the reference's RNG consumption order (:61-92)."""
    if kind == "clean":
        return image_u8
    h, w = image_u8.shape[:2]
    if kind == "fog":
        d = draw_fog(h, w, intensity)
        return fog_apply(image_u8, depth_from_noise(d["noise"]), d["intensity"])
    if kind == "rain":
        d = draw_rain(h, w, intensity)
        return rain_apply(image_u8, d["intensity"], d["drops"])
    if kind == "snow":
        d = draw_snow(h, w, intensity)
        return snow_apply(image_u8, d["intensity"], d["flakes"], d["blur_k"])
    if kind == "night":
        d = draw_night(image_u8.shape, intensity)
        return night_apply(image_u8, d["intensity"], d["reduction"], d["noise"])
    raise ValueError(f"Unknown weather type: {kind}")
