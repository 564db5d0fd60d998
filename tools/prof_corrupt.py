"""A few launches of awx_corrupt for one weather kind on config-2 shaped frames, for ncu captures (dev tool).
usage: prof_corrupt.py kind B n   (kind: fog | night | rain | snow3 | snow7)"""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops
from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import WeatherDegradationTransforms

kind = sys.argv[1] if len(sys.argv) > 1 else "fog"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
n = int(sys.argv[3]) if len(sys.argv) > 3 else 3
h, w = 1024, 2048
t = WeatherDegradationTransforms(seed=3)
rng = np.random.RandomState(0)
imgs = torch.from_numpy(rng.randint(0, 255, (B, h, w, 3)).astype(np.uint8)).cuda()
out = torch.empty_like(imgs)
ws = ops.corrupt_workspace(B, h, w)
base = [t.draw(kind[:4] if kind.startswith("snow") else kind, h, w) for _ in range(2)]
for d in base:
    if kind.startswith("snow"):
        d.blur_k = int(kind[4])
    if d.kind == "fog":
        d.depth = np.maximum(d.depth_noise + 50.0, 1.0)
prm, fld, items = t.pack([base[i % 2] for i in range(B)], h, w, np.float64)
fld_d = None if fld is None else torch.from_numpy(fld).cuda()
items_d = None if items is None else torch.from_numpy(items).cuda()
for _ in range(n):
    ops.corrupt(imgs, prm, fld_d, items_d, out=out, workspace=ws)
torch.cuda.synchronize()
print("ok", int(out.sum()))
