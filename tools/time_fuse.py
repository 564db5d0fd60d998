"""CUDA-event timing of the forward fusion (EnsembleModel.forward's fused logits) per strategy (dev tool)."""
import sys, statistics as st, torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda")
c, h, w = 19, 1024, 2048
la = torch.randn(B, c, h, w, device=dev)
lb = torch.randn(B, c, h, w, device=dev)
px = B * h * w
for name, kw in (("weighted T=1.7", dict(strategy=_lib.FUSE_WEIGHTED, w0=0.35, w1=0.65, temperature=1.7)),
                 ("mean no T", dict(strategy=_lib.FUSE_MEAN, temperature=None)),
                 ("max_confidence T=1", dict(strategy=_lib.FUSE_MAXCONF, temperature=1.0))):
    fns = {"awx_fuse_forward": lambda: ops.fuse_forward(la, lb, kw["strategy"], kw.get("w0", 0.5), kw.get("w1", 0.5), kw["temperature"]),
           "awx_score + fused map": lambda: ops.score(la, lb, want_fused=True, **kw)["fused"]}
    for label, fn in fns.items():
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = st.median(ts)
        print(f"{label:22s} {name:20s} {ms:7.3f} ms  {px*228/ms/1e6:8.1f} GB/s (228 B/px)", flush=True)
out = torch.empty_like(la)
ts = []
for _ in range(10):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); torch.add(la, lb, out=out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
print(f"torch.add(la, lb)          {st.median(ts):7.3f} ms  {px*228/st.median(ts)/1e6:8.1f} GB/s")
g = torch.randn_like(la)
for name, code, temp in (("weighted T=1.7", _lib.FUSE_WEIGHTED, 1.7), ("mean no T", _lib.FUSE_MEAN, None), ("max_confidence T=1", _lib.FUSE_MAXCONF, 1.0)):
    fn = lambda: ops.fuse_backward(g, la, lb, code, 0.35, 0.65, temp)
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = st.median(ts)
    print(f"awx_fuse_backward      {name:20s} {ms:7.3f} ms  {px*380/ms/1e6:8.1f} GB/s (380 B/px)", flush=True)
