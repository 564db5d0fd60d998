"""CUDA-event timing of the 8f-row kernels at Cityscapes frame size (dev tool)."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_prep
from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import scipy_gaussian_weights

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
h, w = 1024, 2048
gen = torch.Generator(device="cuda").manual_seed(0)
imgs = torch.randint(0, 255, (B, h, w, 3), device="cuda", dtype=torch.uint8, generator=gen)
px = B * h * w


def timeit(name, fn, bpp, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:34s} {ms:8.3f} ms  {px/ms/1e3:9.1f} Mpx/s  {px*bpp/ms/1e6:8.1f} GB/s (algorithmic {bpp} B/px)", flush=True)


out32 = torch.empty((B, 3, h, w), dtype=torch.float32, device="cuda")
out16 = torch.empty((B, 3, h, w), dtype=torch.bfloat16, device="cuda")
timeit("normalize_chw fp32", lambda: ops_prep.normalize_chw(imgs, out=out32), 15)
timeit("normalize_chw bf16", lambda: ops_prep.normalize_chw(imgs, out_dtype=torch.bfloat16, out=out16), 9)
so = torch.empty_like(imgs)
timeit("style_transfer night", lambda: ops_prep.style_transfer(imgs, "night", out=so), 6)
wts = scipy_gaussian_weights(2.0)
timeit("estimate_depth (5 kernels)", lambda: ops_prep.estimate_depth(imgs, wts), 3 + 3 + 8 + 16 + 16)
depth = torch.rand(B, h, w, device="cuda", dtype=torch.float64) * 100 + 1
timeit("fog_density_map (9 kernels + 1 sync)", lambda: ops_prep.fog_density_map(imgs, depth), 3 + 4 + 4 + 4 + 8 + 4 + 8 + 8)
Bt = min(B, 8)
logits = torch.randn(Bt, 19, h, w, device="cuda", generator=gen)
labels = torch.randint(0, 19, (Bt, h, w), device="cuda", generator=gen)
temps = torch.linspace(0.1, 10.0, 100)
px = Bt * h * w
timeit("temperature_nll 100 T (MUFU bound)", lambda: ops_prep.temperature_nll(logits, labels, temps), 84, n=2)
