"""Board power / SM clock while one kernel runs back to back for a few seconds (dev tool).
usage: power_probe.py [seconds]   -> one line per workload: median launch ms, power W, SM MHz, throttle reasons"""
import subprocess, sys, threading, time, statistics as st
import torch
sys.path.insert(0, ".")
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib

secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
dev = torch.device("cuda")
B, c, h, w = 64, 19, 1024, 2048
la = torch.randn(B, c, h, w, device=dev)
lb = torch.randn(B, c, h, w, device=dev)
tgt = torch.randint(0, c, (B, h, w), device=dev).to(torch.uint8)
wts = torch.softmax(torch.tensor([0.3, 0.9]), 0)
bins_e, bins_s = ops.new_bins(c, 15, 4096), ops.new_bins(c, 15, 0)
dst = torch.empty_like(la)
work = {
    "idle": lambda: time.sleep(0.004),
    "torch copy 10 GB": lambda: dst.copy_(la),
    "score single": lambda: ops.score(la, None, tgt, bins=bins_s),
    "score ens weighted T=1.7": lambda: ops.score(la, lb, tgt, strategy=_lib.FUSE_WEIGHTED, w0=float(wts[0]), w1=float(wts[1]),
                                                  temperature=1.7, auroc_bins=4096, bins=bins_e),
}
Q = "power.draw,clocks.sm,clocks.mem,temperature.gpu,clocks_throttle_reasons.active"


def sample(stop, out):
    while not stop.is_set():
        r = subprocess.run(["nvidia-smi", "--query-gpu=" + Q, "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True).stdout.strip().split(", ")
        if len(r) >= 5:
            out.append(r)
        time.sleep(0.05)


for name, fn in work.items():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    stop, samples = threading.Event(), []
    th = threading.Thread(target=sample, args=(stop, samples))
    th.start()
    t_end, times = time.time() + secs, []
    while time.time() < t_end:
        evs = [torch.cuda.Event(True) for _ in range(11)]
        evs[0].record()
        for i in range(10):
            fn()
            evs[i + 1].record()
        torch.cuda.synchronize()
        times += [evs[i].elapsed_time(evs[i + 1]) for i in range(10)]
    stop.set()
    th.join()
    half = samples[len(samples) // 2:]  # steady state
    pw = st.median(float(s[0]) for s in half)
    sm = st.median(float(s[1]) for s in half)
    mem = st.median(float(s[2]) for s in half)
    tmp = st.median(float(s[3]) for s in half)
    reasons = sorted({s[4] for s in half})
    print(f"{name:28s} {st.median(times):7.3f} ms (first 10: {st.median(times[:10]):6.3f})  {pw:6.0f} W  SM {sm:5.0f} MHz  mem {mem:5.0f} MHz  "
          f"{tmp:3.0f} C  reasons {reasons}  ({len(samples)} samples)", flush=True)
    time.sleep(2.0)
