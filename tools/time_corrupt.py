"""CUDA-event timing of awx_corrupt per weather kind on config-2 shaped frames (dev tool)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops
from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import WeatherDegradationTransforms

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
fdt = np.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else np.float64
h, w = 1024, 2048
t = WeatherDegradationTransforms(seed=3)
rng = np.random.RandomState(0)
imgs = torch.from_numpy(rng.randint(0, 255, (B, h, w, 3)).astype(np.uint8)).cuda()
out = torch.empty_like(imgs)
ws = ops.corrupt_workspace(B, h, w)
fsz = 8 if fdt == np.float64 else 4
for kind, bpp in (("fog", 6 + fsz), ("night", 6 + 3 * fsz), ("rain", 6.125), ("snow3", 6.125), ("snow7", 6.125)):
    base = [t.draw(kind[:4] if kind.startswith("snow") else kind, h, w) for _ in range(2)]
    for d in base:
        if kind.startswith("snow"):
            d.blur_k = int(kind[4])
        if d.kind == "fog":
            d.depth = np.maximum(d.depth_noise + 50.0, 1.0)
    draws = [base[i % 2] for i in range(B)]
    prm, fld, items = t.pack(draws, h, w, fdt)
    fld_d = None if fld is None else torch.from_numpy(fld).cuda()
    items_d = None if items is None else torch.from_numpy(items).cuda()
    fn = lambda: ops.corrupt(imgs, prm, fld_d, items_d, out=out, workspace=ws)
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    n = 5
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    px = B * h * w
    print(f"{kind:6s} {ms:8.3f} ms  {px/ms/1e3:9.1f} Mpx/s  {px*bpp/ms/1e6:8.1f} GB/s", flush=True)

# fused Normalize + CHW epilogue vs two passes
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
norm = torch.empty((B, 3, h, w), dtype=torch.float32, device="cuda")
npar = ops_prep.normalize_params(ops_prep.IMAGENET_MEAN, ops_prep.IMAGENET_STD)
for kind in ("night", "rain"):
    base = [t.draw(kind, h, w) for _ in range(2)]
    draws = [base[i % 2] for i in range(B)]
    prm, fld, items = t.pack(draws, h, w, fdt)
    fld_d = None if fld is None else torch.from_numpy(fld).cuda()
    items_d = None if items is None else torch.from_numpy(items).cuda()
    for name, fn in (("two passes (u8, then normalize)", lambda: ops_prep.normalize_chw(ops.corrupt(imgs, prm, fld_d, items_d, out=out, workspace=ws), out=norm)),
                     ("fused epilogue, u8 + fp32", lambda: ops.corrupt(imgs, prm, fld_d, items_d, out=out, workspace=ws, norm_out=norm, norm_params=npar)),
                     ("fused epilogue, fp32 only", lambda: ops.corrupt(imgs, prm, fld_d, items_d, workspace=ws, norm_out=norm, norm_params=npar, write_u8=False))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        print(f"{kind:6s} {name:34s} {e0.elapsed_time(e1)/5:8.3f} ms", flush=True)
