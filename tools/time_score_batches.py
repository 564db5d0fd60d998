"""Fixed versus per-frame cost of one awx_score launch (ensemble, weighted, T = 1.7, uint8 labels, bins only): B = 1 .. 16
frames of 1024x2048, 30 launches back to back behind a spin kernel (dev tool)."""
import sys, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

dev = torch.device("cuda")
c, h, w, bmax = 19, 1024, 2048, 16
la = torch.randn(bmax, c, h, w, device=dev)
lb = torch.randn(bmax, c, h, w, device=dev)
tgt = torch.randint(0, c, (bmax, h, w), device=dev).to(torch.uint8)
rows = []
for nb_auroc, name in ((4096, "ensemble"), (0, "single member")):
    for b in (1, 2, 4, 8, 16):
        bins = ops.new_bins(c, 15, nb_auroc)
        if nb_auroc:
            fn = lambda: ops.score(la[:b], lb[:b], tgt[:b], strategy=_lib.FUSE_WEIGHTED, w0=0.354, w1=0.646, temperature=1.7, auroc_bins=4096, bins=bins)
        else:
            fn = lambda: ops.score(la[:b], None, tgt[:b], bins=bins)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        torch.cuda._sleep(100_000_000)
        e0.record()
        for _ in range(30):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        rows.append((name, b, ms))
        print(f"{name:14s} B = {b:2d}: {ms * 1e3:8.1f} us per launch, {ms * 1e3 / b:7.1f} us per frame", flush=True)
