"""CUDA-event timing of awx_fogloss at BASELINE config 4 (B=8, 1024x2048, fwd+bwd in one launch). Dev tool."""
import sys, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_loss

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
c, h, w = 19, 1024, 2048
dev = torch.device("cuda")
gen = torch.Generator(device=dev).manual_seed(0)
logits = torch.randn(B, c, h, w, device=dev, generator=gen)
lab64 = torch.randint(0, c, (B, h, w), device=dev, generator=gen)
lab8 = lab64.to(torch.uint8)
fd = torch.rand(B, h, w, device=dev, generator=gen)
dp = torch.rand(B, h, w, device=dev, generator=gen) * 50
dt = torch.rand(B, h, w, device=dev, generator=gen) * 50
px = B * h * w


def timeit(name, fn, bpp, n=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"{name:44s} {ms:8.3f} ms  {px/ms/1e3:9.1f} Mpx/s  {px*bpp/ms/1e6:8.1f} GB/s", flush=True)


timeit("fwd+bwd int64 labels, fog, depth (176 B/px)", lambda: ops_loss.fogloss_raw(logits, lab64, fd, dp, dt, 2.0, False, True), 176)
timeit("fwd+bwd uint8 labels, fog, depth (169 B/px)", lambda: ops_loss.fogloss_raw(logits, lab8, fd, dp, dt, 2.0, False, True), 169)
timeit("fwd only int64 labels, fog (88 B/px)", lambda: ops_loss.fogloss_raw(logits, lab64, fd, None, None, 2.0, False, False), 88)
timeit("focal fwd+bwd uint8, fog, depth (169 B/px)", lambda: ops_loss.fogloss_raw(logits, lab8, fd, dp, dt, 2.0, True, True), 169)
