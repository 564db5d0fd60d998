"""A/B of builds of libawx.so on the bench's score launch, in ONE process (dev tool).
usage: ab_score.py [--batch 64] [--rounds 3] [--launches 20] name=path ...
Every round times `launches` back-to-back ensemble launches (weighted, T = 1.7, uint8 labels, bins only: the kernel the
roofline is quoted on) per build, CUDA events around the whole run (sustained clocks, as inside the bench step)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

args = sys.argv[1:]
def opt(name, default):
    if name in args:
        i = args.index(name); v = args[i + 1]; del args[i:i + 2]; return int(v)
    return default
B, rounds, n = opt("--batch", 64), opt("--rounds", 3), opt("--launches", 20)
builds = [a.split("=", 1) for a in args]
dev = torch.device("cuda")
c, h, w = 19, 1024, 2048
gen = torch.Generator(device=dev).manual_seed(42)
la = torch.randn(B, c, h, w, device=dev, generator=gen)
lb = torch.randn(B, c, h, w, device=dev, generator=gen)
tgt = torch.randint(0, c, (B, h, w), device=dev, generator=gen, dtype=torch.uint8)
px = B * h * w
ref = None
for r in range(rounds):
    for name, path in builds:
        _lib._lib = None
        os.environ["AWX_LIB"] = path
        _lib.load()
        bins = ops.new_bins(c, 15, 4096)
        fn = lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=0.354, w1=0.646, temperature=1.7, auroc_bins=4096, bins=bins)
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        bins.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        digest = int(bins.sum().item())
        ref = digest if ref is None else ref
        print(f"round {r} {name:12s} {ms:7.3f} ms/launch  {px * 153 / ms / 1e6:7.1f} GB/s  bins {'same' if digest == ref else 'DIFFER'}", flush=True)
