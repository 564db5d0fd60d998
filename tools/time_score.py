"""Quick CUDA-event timing of awx_score on config-2/3 shaped inputs (dev tool, not the bench)."""
import sys, torch, time
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
c, h, w = 19, 1024, 2048
la = torch.randn(B, c, h, w, device=dev)
lb = torch.randn(B, c, h, w, device=dev)
tgt = torch.randint(0, c, (B, h, w), device=dev).to(torch.uint8)
px = B * h * w

def timeit(name, fn, bytes_per_px, n=20):
    """median (and min) of n individually timed launches"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(n)]
    for e0, e1 in ev:
        e0.record()
        fn()
        e1.record()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in ev)
    ms, mn = t[len(t) // 2], t[0]
    print(f"{name:34s} {ms:8.3f} ms (min {mn:6.3f})  {px/ms/1e3:9.1f} Mpx/s  {px*bytes_per_px/ms/1e6:8.1f} GB/s", flush=True)

bins1 = ops.new_bins(c, 15, 0)
bins2 = ops.new_bins(c, 15, 4096)
timeit("single (77 B/px)", lambda: ops.score(la, None, tgt, bins=bins1), 77)
timeit("ens weighted T=1 (153 B/px)", lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, auroc_bins=4096, bins=bins2), 153)
timeit("ens weighted T=1.7 (153 B/px)", lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, temperature=1.7, auroc_bins=4096, bins=bins2), 153)
timeit("ens mean noT (153 B/px)", lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_MEAN, auroc_bins=4096, bins=bins2), 153)
timeit("ens max_confidence T=1.7 (153 B/px)", lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_MAXCONF, temperature=1.7, auroc_bins=4096, bins=bins2), 153)
cp = torch.empty_like(la)
timeit("torch copy (152 B/px r+w)", lambda: cp.copy_(la), 152)
