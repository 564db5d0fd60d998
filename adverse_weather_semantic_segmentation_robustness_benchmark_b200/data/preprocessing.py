"""Drop-in ``WeatherDegradationTransforms`` (reference: data/preprocessing.py:15-288).

Split the way the north-star asks: the STOCHASTIC parameters (fog depth noise and intensity, rain
drop positions / lengths / thickness / angle, snow flake positions / radii / blur size, night
brightness draw and noise field) are drawn on the HOST from the process-global legacy NumPy RNG in
exactly the reference's order, so a given seed yields the same parameters; the PER-PIXEL work
(u8->fp32, scattering blend, haze, streak / flake scan conversion, Gaussian blur, colour shift,
noise add, clip, truncate) runs in libawx.so on the GPU.  NumPy in -> NumPy out for the reference
signatures; ``corrupt_batch`` is the batched device-resident extension used by the sweep driver.
"""

from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass, field as dc_field
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _lib, ops, ops_prep

logger = logging.getLogger(__name__)

_KINDS = ("clean", "fog", "rain", "snow", "night")


def gaussian_taps(ksize: int, sigma: float) -> np.ndarray:
    """fp32 taps of cv2.getGaussianKernel(ksize, sigma, CV_32F): exp(-x^2/(2 sigma^2)) in double,
    normalised to sum 1, rounded to fp32 (OpenCV imgproc/smooth; checked against cv2 in the tests)."""
    x = np.arange(ksize, dtype=np.float64) - (ksize - 1) * 0.5
    k = np.exp(-(x * x) / (2.0 * sigma * sigma))
    return (k / k.sum()).astype(np.float32)


def scipy_gaussian_weights(sigma: float = 2.0, truncate: float = 4.0) -> np.ndarray:
    """fp64 weights of scipy.ndimage.gaussian_filter1d (order 0): radius int(truncate*sigma+0.5)."""
    radius = int(truncate * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    return phi / phi.sum()


@dataclass
class WeatherDraw:
    """Host-drawn parameters of one frame's corruption (what crosses the C ABI)."""
    kind: str
    intensity: float = 0.0
    depth_noise: Optional[np.ndarray] = None   # fog: N(0,10) field [H,W] fp64 (input of the depth filter)
    depth: Optional[np.ndarray] = None         # fog: filtered depth [H,W] (filled by the device filter or given)
    items: Optional[np.ndarray] = None         # rain: [n,5] x0,y0,x1,y1,thickness; snow: [n,5] x,y,r,0,0
    blur_k: int = 0
    reduction: float = 0.0                     # night: the uniform(0.2,0.6) draw
    noise: Optional[np.ndarray] = None         # night: N(0,5/255) field [H,W,3] fp64


class WeatherDegradationTransforms:
    """Weather degradation transforms for synthetic adverse conditions (GPU implementation)."""

    def __init__(self, seed: Optional[int] = None) -> None:
        # data/preprocessing.py:30-31: the seed goes to the process-global legacy RNG
        if seed is not None:
            np.random.seed(seed)
        self.fog_parameters = {"beta_range": (0.005, 0.05), "A_range": (0.7, 1.0), "depth_scale": 100.0}
        self.rain_parameters = {"intensity_range": (0.1, 0.8), "drop_size_range": (1, 3),
                                "angle_range": (-15, 15), "num_drops_range": (100, 500)}
        self.snow_parameters = {"intensity_range": (0.1, 0.7), "flake_size_range": (2, 8),
                                "num_flakes_range": (50, 200), "blur_kernel": (3, 7)}
        self.night_parameters = {"brightness_reduction": (0.2, 0.6),
                                 "color_shift": {"r": 0.8, "g": 0.85, "b": 1.2}, "noise_std": 5.0}
        logger.info("Initialized WeatherDegradationTransforms (libawx)")

    # ------------------------------------------------------------------ host draws (RNG replay)
    def draw(self, weather_type: str, height: int, width: int, intensity: Optional[float] = None) -> WeatherDraw:
        """Consume np.random exactly as the reference's ``_apply_*`` would (orders: SURVEY 8 a3-a7)."""
        if weather_type == "clean":
            return WeatherDraw("clean")
        if weather_type == "fog":          # :104 (-> :239) then :107-108
            noise = np.random.normal(0, 10, (height, width))
            if intensity is None:
                intensity = np.random.uniform(0.3, 0.9)
            return WeatherDraw("fog", float(intensity), depth_noise=noise)
        if weather_type == "rain":         # :127-156
            if intensity is None:
                intensity = np.random.uniform(0.2, 0.8)
            lo, hi = self.rain_parameters["num_drops_range"]
            n = int(lo + intensity * (hi - lo))
            items = np.zeros((n, 5), dtype=np.int32)
            for i in range(n):
                x = np.random.randint(0, width)
                y = np.random.randint(0, height)
                length = np.random.randint(5, 20)
                thickness = np.random.choice(self.rain_parameters["drop_size_range"])
                angle = np.random.uniform(*self.rain_parameters["angle_range"])
                ex = int(x + length * np.sin(np.radians(angle)))
                ey = int(y + length * np.cos(np.radians(angle)))
                items[i] = (x, y, min(max(ex, 0), width - 1), min(max(ey, 0), height - 1), thickness)
            return WeatherDraw("rain", float(intensity), items=items, blur_k=3)
        if weather_type == "snow":         # :172-199
            if intensity is None:
                intensity = np.random.uniform(0.2, 0.7)
            lo, hi = self.snow_parameters["num_flakes_range"]
            n = int(lo + intensity * (hi - lo))
            items = np.zeros((n, 5), dtype=np.int32)
            for i in range(n):
                x = np.random.randint(0, width)
                y = np.random.randint(0, height)
                items[i, :3] = (x, y, np.random.choice(self.snow_parameters["flake_size_range"]))
            k = int(np.random.choice(self.snow_parameters["blur_kernel"]))
            if k % 2 == 0:
                k += 1
            return WeatherDraw("snow", float(intensity), items=items, blur_k=k)
        if weather_type == "night":        # :206-222
            if intensity is None:
                intensity = np.random.uniform(0.4, 0.8)
            reduction = np.random.uniform(*self.night_parameters["brightness_reduction"])
            noise = np.random.normal(0, self.night_parameters["noise_std"] / 255.0, (height, width, 3))
            return WeatherDraw("night", float(intensity), reduction=float(reduction), noise=noise)
        raise ValueError(f"Unknown weather type: {weather_type}")

    # ------------------------------------------------------------------ parameter packing
    def pack(self, draws: Sequence[WeatherDraw], height: int, width: int, field_dtype=np.float64,
             gather_fields: bool = True):
        """AwxCorruptParams records + concatenated field / item arrays for a batch of draws.
        ``gather_fields=False``: the depth / noise fields already live on the device, laid out frame
        after frame; only their offsets (H*W per fog frame, H*W*3 per night frame) are recorded."""
        prm = np.zeros(len(draws), dtype=_lib.CORRUPT_PARAMS_DTYPE)
        fields: List[np.ndarray] = []
        items: List[np.ndarray] = []
        f_off = 0
        i_off = 0
        for i, d in enumerate(draws):
            prm[i]["kind"] = _lib.KIND_CODES[d.kind]
            if d.kind == "fog":
                b0, b1 = self.fog_parameters["beta_range"]
                a0, a1 = self.fog_parameters["A_range"]
                prm[i]["d0"] = b0 + d.intensity * (b1 - b0)
                # A * np.ones_like(fp32 image) rounds A to fp32 before the fp64 blend (:118)
                prm[i]["d1"] = np.float64(np.float32(a0 + d.intensity * (a1 - a0)))
                prm[i]["field_offset"] = f_off
                if gather_fields:
                    if d.depth is None:
                        raise ValueError("fog draw has no depth map; call synthetic_depth() first")
                    fields.append(np.ascontiguousarray(d.depth, dtype=field_dtype).reshape(-1))
                f_off += height * width
            elif d.kind == "night":
                prm[i]["d0"] = d.intensity
                prm[i]["f0"] = np.float32(1 - d.intensity * d.reduction)
                prm[i]["field_offset"] = f_off
                if gather_fields:
                    fields.append(np.ascontiguousarray(d.noise, dtype=field_dtype).reshape(-1))
                f_off += height * width * 3
            elif d.kind in ("rain", "snow"):
                if d.kind == "rain":
                    haze = d.intensity * 0.3
                    prm[i]["f0"] = np.float32(1 - haze)
                    prm[i]["f1"] = np.float32(haze * 0.7)
                    taps = gaussian_taps(3, 0.5)
                else:
                    prm[i]["f0"] = np.float32(d.intensity * 0.2)
                    taps = gaussian_taps(d.blur_k, 1.0)
                prm[i]["blur_k"] = d.blur_k
                half = taps[len(taps) // 2:]
                prm[i]["taps"][:len(half)] = half
                prm[i]["item_begin"] = i_off
                prm[i]["item_count"] = len(d.items)
                items.append(np.ascontiguousarray(d.items, dtype=np.int32))
                i_off += len(d.items)
        field = np.concatenate(fields) if fields else None
        item_arr = np.concatenate(items, axis=0) if items else None
        return prm, field, item_arr

    # ------------------------------------------------------------------ device entry points
    def synthetic_depth(self, noise: np.ndarray, out_dtype=torch.float64) -> torch.Tensor:
        """ramp + noise -> Gaussian sigma=2 (scipy 'reflect') -> max(.,1), on the device (:235-246)."""
        lib = _lib.load()
        dev = ops.require_cuda()
        nz = torch.from_numpy(np.ascontiguousarray(noise, dtype=np.float64)).to(dev)
        if nz.dim() == 2:
            nz = nz.unsqueeze(0)
        b, h, w = nz.shape
        out = torch.empty((b, h, w), dtype=out_dtype, device=dev)
        tmp = torch.empty((b, h, w), dtype=torch.float64, device=dev)
        wts = np.ascontiguousarray(scipy_gaussian_weights(2.0), dtype=np.float64)
        rc = lib.awx_synth_depth(C.c_void_p(nz.data_ptr()), C.c_void_p(out.data_ptr()),
                                 _lib.F64 if out_dtype == torch.float64 else _lib.F32,
                                 C.c_void_p(tmp.data_ptr()), b, h, w, float(self.fog_parameters["depth_scale"]),
                                 wts.ctypes.data_as(C.c_void_p), (len(wts) - 1) // 2,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
        _lib.check(rc, "awx_synth_depth")
        return out

    def corrupt_batch(self, images, draws: Sequence[WeatherDraw], field_dtype=np.float64,
                      out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 [B,H,W,3] (NumPy or torch, host or device) + one draw per frame -> uint8 device tensor."""
        dev = ops.require_cuda()
        img = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
        if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
            raise ValueError(f"images must be uint8 [B,H,W,3], got {img.dtype} {tuple(img.shape)}")
        img = ops.to_device(img)
        b, h, w, _ = img.shape
        if len(draws) != b:
            raise ValueError(f"{len(draws)} draws for {b} frames")
        for d in draws:
            if d.kind == "fog" and d.depth is None:
                d.depth = self.synthetic_depth(d.depth_noise)[0].cpu().numpy()
        prm, fld, items = self.pack(draws, h, w, field_dtype)
        fld_d = None if fld is None else torch.from_numpy(fld).to(dev)
        items_d = None if items is None else torch.from_numpy(items).to(dev)
        return ops.corrupt(img, prm, fld_d, items_d, out=out)

    def corrupt_batch_normalized(self, images, draws: Sequence[WeatherDraw], mean=ops_prep.IMAGENET_MEAN,
                                 std=ops_prep.IMAGENET_STD, out_dtype: torch.dtype = torch.float32,
                                 field_dtype=np.float64, keep_u8: bool = False):
        """corrupt_batch with the dataset's Normalize(mean, std) + ToTensorV2 (loader.py:196-199) fused into
        the corruption kernels: returns the fp32 / bf16 [B,3,H,W] tensor the backbones consume (and the
        uint8 frames too when ``keep_u8``), in one pass over HBM."""
        dev = ops.require_cuda()
        img = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
        if img.dtype != torch.uint8 or img.dim() != 4 or img.shape[-1] != 3:
            raise ValueError(f"images must be uint8 [B,H,W,3], got {img.dtype} {tuple(img.shape)}")
        img = ops.to_device(img)
        b, h, w, _ = img.shape
        if len(draws) != b:
            raise ValueError(f"{len(draws)} draws for {b} frames")
        for d in draws:
            if d.kind == "fog" and d.depth is None:
                d.depth = self.synthetic_depth(d.depth_noise)[0].cpu().numpy()
        prm, fld, items = self.pack(draws, h, w, field_dtype)
        fld_d = None if fld is None else torch.from_numpy(fld).to(dev)
        items_d = None if items is None else torch.from_numpy(items).to(dev)
        norm = torch.empty((b, 3, h, w), dtype=out_dtype, device=dev)
        u8 = ops.corrupt(img, prm, fld_d, items_d, norm_out=norm,
                         norm_params=ops_prep.normalize_params(mean, std), write_u8=keep_u8)
        return (norm, u8) if keep_u8 else norm

    # ------------------------------------------------------------------ reference signatures
    def apply_weather_effect(self, image: np.ndarray, weather_type: str,
                             intensity: Optional[float] = None) -> np.ndarray:
        """data/preprocessing.py:61-92: HWC image in, uint8 HWC out; 'clean' returns the same object."""
        if weather_type == "clean":
            return image
        if weather_type not in _KINDS:
            raise ValueError(f"Unknown weather type: {weather_type}")
        # the reference converts whatever dtype it is given with astype(float32)/255; the kernel's
        # input is the uint8 frame itself, which is what the dataset hands over (loader.py:264-267)
        if image.dtype != np.uint8:
            raise TypeError("libawx corrupts uint8 frames; got dtype %s" % image.dtype)
        h, w = image.shape[:2]
        d = self.draw(weather_type, h, w, intensity)
        return self.corrupt_batch(image[np.newaxis], [d])[0].cpu().numpy()

    @staticmethod
    def _unit_to_u8(image: np.ndarray) -> np.ndarray:
        u8 = np.rint(np.asarray(image, dtype=np.float64) * 255.0).clip(0, 255).astype(np.uint8)
        if not np.array_equal(u8.astype(np.float32) / 255.0, np.asarray(image, dtype=np.float32)):
            raise NotImplementedError(
                "the private _apply_* entry points accept the reference's own intermediate (uint8/255 as "
                "float32); arbitrary float images are outside what the uint8 kernels consume")
        return u8

    def _apply_one(self, kind: str, image: np.ndarray, intensity: Optional[float]) -> np.ndarray:
        u8 = self._unit_to_u8(image)
        d = self.draw(kind, u8.shape[0], u8.shape[1], intensity)
        return self.corrupt_batch(u8[np.newaxis], [d])[0].cpu().numpy()

    def _apply_fog(self, image: np.ndarray, intensity: Optional[float] = None) -> np.ndarray:
        return self._apply_one("fog", image, intensity)

    def _apply_rain(self, image: np.ndarray, intensity: Optional[float] = None) -> np.ndarray:
        return self._apply_one("rain", image, intensity)

    def _apply_snow(self, image: np.ndarray, intensity: Optional[float] = None) -> np.ndarray:
        return self._apply_one("snow", image, intensity)

    def _apply_night(self, image: np.ndarray, intensity: Optional[float] = None) -> np.ndarray:
        return self._apply_one("night", image, intensity)

    def _generate_synthetic_depth(self, height: int, width: int) -> np.ndarray:
        """:227-248: draws the N(0,10) field on the host, filters on the device; fp64 [H,W]."""
        noise = np.random.normal(0, 10, (height, width))
        return self.synthetic_depth(noise)[0].cpu().numpy()

    def get_fog_density_map(self, image: np.ndarray, depth: Optional[np.ndarray] = None) -> np.ndarray:
        """:250-288: fog density in [0,1] for the fog-density-aware loss.  `image`: HWC float in [0,1]
        (the reference's contract; uint8 frames are taken as they are); `depth`: [H,W] or None (then the
        synthetic depth is drawn from the global NumPy RNG as the reference does)."""
        h, w = image.shape[:2]
        if depth is None:
            depth = self._generate_synthetic_depth(h, w)
        img = np.ascontiguousarray(image)
        if img.dtype not in (np.uint8, np.float32, np.float64):
            img = img.astype(np.float64)
        out = ops_prep.fog_density_map(torch.from_numpy(img)[None], torch.from_numpy(np.ascontiguousarray(depth))[None])
        return out[0].cpu().numpy()


class DepthEstimationPreprocessor:
    """Depth estimation preprocessor (reference: data/preprocessing.py:291-411); the per-pixel work of
    ``estimate_depth`` runs in libawx.so (awx_estimate_depth)."""

    def __init__(self) -> None:
        self.depth_model = None
        logger.info("Initialized DepthEstimationPreprocessor")

    def estimate_depth(self, image: np.ndarray) -> np.ndarray:
        """:304-326: uint8 RGB HWC -> fp64 [H,W] in [0,1]."""
        return self._geometric_depth_estimation(image)

    def _geometric_depth_estimation(self, image: np.ndarray) -> np.ndarray:
        """:328-367: perspective ramp, sky / road bands, Laplacian texture cue, Gaussian sigma=2."""
        if image.dtype != np.uint8 or image.ndim != 3 or image.shape[2] != 3:
            raise TypeError("libawx estimates depth from uint8 RGB HWC frames; got %s %s" % (image.dtype, image.shape))
        out = ops_prep.estimate_depth(torch.from_numpy(np.ascontiguousarray(image))[None], scipy_gaussian_weights(2.0))
        return out[0].cpu().numpy()

    def estimate_depth_batch(self, images) -> torch.Tensor:
        """Batched, device-resident extension: uint8 [B,H,W,3] -> fp64 [B,H,W] (stays on the GPU)."""
        img = torch.from_numpy(np.ascontiguousarray(images)) if isinstance(images, np.ndarray) else images
        return ops_prep.estimate_depth(img, scipy_gaussian_weights(2.0))

    def depth_to_disparity(self, depth: np.ndarray, baseline: float = 0.54) -> np.ndarray:
        """:369-384 (a handful of host scalars per call in the reference's users; kept as NumPy)."""
        return baseline / np.maximum(depth, 1e-6)

    def preprocess_depth_for_training(self, depth: np.ndarray, target_size) -> torch.Tensor:
        """:386-411: resize to (height, width) if needed, min-max normalise in the array's dtype, fp32 tensor.
        Per-sample dataset-side work outside the hot path (the dataset's ``__getitem__``), kept on the host
        with the reference's own calls so that the class is complete."""
        if depth.shape != tuple(target_size):
            import cv2
            depth = cv2.resize(depth, (target_size[1], target_size[0]))
        lo = np.min(depth)
        return torch.from_numpy((depth - lo) / (np.max(depth) - lo + 1e-8)).float()

