"""Lead of DESIGN.md section 10: the four corruptions of a bench step on four streams instead of one (dev experiment).
Times awx_corrupt for fog, rain, snow (3- and 7-tap frames alternating) and night on B frames of 1024x2048, first one
after the other on one stream, then side by side on four streams, CUDA events around each group of four."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops
from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import WeatherDegradationTransforms

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
h, w = 1024, 2048
t = WeatherDegradationTransforms(seed=3)
rng = np.random.RandomState(0)
imgs = torch.from_numpy(rng.randint(0, 255, (B, h, w, 3)).astype(np.uint8)).cuda()
jobs = []
for kind in ("fog", "rain", "snow", "night"):
    base = [t.draw(kind, h, w) for _ in range(2)]
    for i, d in enumerate(base):
        if kind == "snow":
            d.blur_k = (3, 7)[i]
        if kind == "fog":
            d.depth = np.maximum(d.depth_noise + 50.0, 1.0)
    prm, fld, items = t.pack([base[i % 2] for i in range(B)], h, w, np.float64)
    jobs.append((kind, prm, None if fld is None else torch.from_numpy(fld).cuda(),
                 None if items is None else torch.from_numpy(items).cuda(), torch.empty_like(imgs), ops.corrupt_workspace(B, h, w)))
streams = [torch.cuda.Stream() for _ in jobs]
main = torch.cuda.current_stream()


def serial():
    for kind, prm, fld, items, out, ws in jobs:
        ops.corrupt(imgs, prm, fld, items, out=out, workspace=ws)


def side_by_side():
    for s, (kind, prm, fld, items, out, ws) in zip(streams, jobs):
        s.wait_stream(main)
        with torch.cuda.stream(s):
            ops.corrupt(imgs, prm, fld, items, out=out, workspace=ws)
    for s in streams:
        main.wait_stream(s)


def timeit(fn, n=8):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    torch.cuda._sleep(100_000_000)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


serial()
ref = [j[4].clone() for j in jobs]
side_by_side()
torch.cuda.synchronize()
same = all(torch.equal(a, j[4]) for a, j in zip(ref, jobs))
for r in range(2):
    print(f"round {r}: one stream {timeit(serial):.3f} ms, four streams {timeit(side_by_side):.3f} ms per group of 4 (B = {B}); outputs identical: {same}", flush=True)
