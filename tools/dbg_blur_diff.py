"""Where do the strip and the tile blur kernels differ, and which one agrees with cv2? (dev tool)"""
import os, sys, numpy as np, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import WeatherDegradationTransforms
from oracle import weather as ow

for h, w, b in ((1024, 2048, 4), (200, 512, 6)):
    t = WeatherDegradationTransforms(seed=3)
    rng = np.random.RandomState(b)
    imgs = rng.randint(0, 256, (b, h, w, 3)).astype(np.uint8)
    kinds = ["rain", "snow", "snow", "clean", "rain", "snow"][:b]
    draws = [t.draw(k, h, w) for k in kinds]
    draws[1].blur_k, draws[2].blur_k = 7, 3
    os.environ.pop("AWX_BLUR_KERNEL", None)
    a = t.corrupt_batch(imgs, draws).cpu().numpy()
    os.environ["AWX_BLUR_KERNEL"] = "tile"
    c = t.corrupt_batch(imgs, draws).cpu().numpy()
    os.environ.pop("AWX_BLUR_KERNEL", None)
    idx = np.argwhere(a != c)
    print(h, w, b, "differences:", len(idx))
    for i, y, x, ch in idx[:10]:
        d = draws[i]
        want = ow.rain_apply(imgs[i], d.intensity, d.items) if d.kind == "rain" else ow.snow_apply(imgs[i], d.intensity, d.items[:, :3], d.blur_k)
        print(" image", i, d.kind, getattr(d, "blur_k", None), "y", y, "x", x, "c", ch, "strip", a[i, y, x, ch], "tile", c[i, y, x, ch], "cv2", want[y, x, ch],
              "| strip==cv2 everywhere:", bool(np.array_equal(a[i], want)), " tile mismatches:", int((c[i] != want).sum()))
