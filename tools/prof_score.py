"""Few launches of awx_score (config-shaped planes) for ncu captures (dev tool).
usage: prof_score.py B n mode [T]   (mode: ens | single)"""
import sys, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = sys.argv[3] if len(sys.argv) > 3 else "ens"
T = float(sys.argv[4]) if len(sys.argv) > 4 else 1.7
dev = torch.device("cuda")
c, h, w = 19, 1024, 2048
la = torch.randn(B, c, h, w, device=dev)
lb = torch.randn(B, c, h, w, device=dev)
tgt = torch.randint(0, c, (B, h, w), device=dev).to(torch.uint8)
wts = torch.softmax(torch.tensor([0.3, 0.9]), 0)
if mode == "ens":
    bins = ops.new_bins(c, 15, 4096)
    for _ in range(n):
        ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=float(wts[0]), w1=float(wts[1]), temperature=T,
                  auroc_bins=4096, bins=bins)
else:
    bins = ops.new_bins(c, 15, 0)
    for _ in range(n):
        ops.score(la, None, tgt, bins=bins)
torch.cuda.synchronize()
print("ok", int(bins.sum()))
