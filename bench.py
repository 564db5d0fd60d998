#!/usr/bin/env python
"""Headline benchmark: fused corrupt -> ensemble fuse -> score (confusion / ECE / AUROC bins).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload (BASELINE.json configs[1], with the two-member fusion of the metric's name): per step and
per rank, for each of the 5 conditions (clean, fog, rain, snow, night) corrupt a batch of 64 synthetic
Cityscapes-shaped frames (1024x2048 uint8) and score two fp32 logit tensors [64,19,1024,2048] against
uint8 labels into that condition's integer bins; ranks shard frames (weak scaling) and one NCCL
all_reduce of the packed int64 bins closes the step.  One JSON line is printed by rank 0.

  value     device-resident inputs, CUDA-event time of exactly K steps, max over ranks
  e2e       same step through the public API with HOST (pinned) buffers: H2D of every input and
            D2H of the bins inside the timed region
  roofline  awx_score (the dominant kernel): 153 algorithmic B/px x pixels per launch / mean launch
            time (CUDA events on the launching stream, inside the timed region) vs the measured copy peak
  cpu_baseline  the REFERENCE's own hot-path code (oracle/_ref: its three modules byte-compiled by
            oracle/build_ref.py; the oracle port only if that is missing) on 2 full frames per condition,
            timed on this box's host cores (rank 0, N=1 only)
--impl reference times that CPU implementation alone, on full-size frames.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONDITIONS = ("clean", "fog", "rain", "snow", "night")
NUM_CLASSES = 19
SCORE_BYTES_PER_PX = 153.0          # 2 x 19 x 4 (logits) + 1 (uint8 label); SURVEY.md 8d
CORRUPT_BYTES_PER_PX = {"clean": 0.0, "fog": 14.0, "rain": 6.125, "snow": 6.0, "night": 30.0}  # fp64 fields
METRIC = "Mpixel/s fused corrupt->ensemble->mIoU/ECE eval"
RAW_WEIGHTS = (0.3, 0.9)
TEMPERATURE = 1.7
AUROC_BINS = 4096
POOL = 64                            # distinct host-drawn parameter sets per condition (a pool of 64 per rank)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--height", type=int, default=1024)
    ap.add_argument("--width", type=int, default=2048)
    ap.add_argument("--sweep-frames", type=int, default=0,
                    help="BASELINE configs[4]: evaluation sweep over this many frames in total (e.g. 10000), sharded "
                         "over the ranks, one all_reduce at the end; prints its own JSON line and exits")
    ap.add_argument("--pool", type=int, default=POOL, help="host-drawn corruption parameter sets per condition")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2..4] side measurements")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi polled every 50 ms from BEFORE the warm-up (its start-up takes longer than a short timed region);
    every line is stamped on arrival and `stop` keeps the samples that fall inside the marked region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.t_begin = self.t_end = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.monotonic(), line.strip()))

    def begin_region(self):
        self.t_begin = time.monotonic()

    def end_region(self):
        self.t_end = time.monotonic()

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t0 = self.t_begin if self.t_begin is not None else 0.0
        t1 = self.t_end if self.t_end is not None else float("inf")
        # a sample reports the state at some point of the 50 ms before it arrives
        inside = [ln for t, ln in self.lines if t0 <= t <= t1 + 0.06]
        where = "timed region"
        if not inside and self.lines:   # region shorter than one polling interval: the sample nearest to it
            mid = 0.5 * (t0 + min(t1, self.lines[-1][0]))
            inside = [min(self.lines, key=lambda tl: abs(tl[0] - mid))[1]]
            where = "nearest sample (region shorter than the 50 ms polling interval)"
        sm, mx, power, reasons = [], None, [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in inside:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
                power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "power_w_max": max(power) if power else None, "sampled": where, "reasons": sorted(reasons)}


# ------------------------------------------------------------------- CPU arm (the reference itself)
def cpu_sample_inputs(h, w, n_frames=1, seed=0):
    """`n_frames` synthetic frames the way the reference's synthetic dataset makes them (loader.py:206,231) plus two
    members' logits each (torch.randn, conftest.py:188,200)."""
    import numpy as np
    import torch
    frames = []
    for i in range(n_frames):
        rng = np.random.RandomState(seed + i)
        img = rng.randint(0, 255, (h, w, 3)).astype(np.uint8)
        lab = torch.from_numpy(rng.randint(0, NUM_CLASSES, (1, h, w)).astype(np.uint8))
        gen = torch.Generator().manual_seed(42 + seed + i)
        frames.append((img, lab, torch.randn(1, NUM_CLASSES, h, w, generator=gen),
                       torch.randn(1, NUM_CLASSES, h, w, generator=gen)))
    return frames


def cpu_kind():
    """'reference': oracle/_ref holds the reference's own byte-compiled hot-path modules (oracle/build_ref.py, built
    where /root/reference exists and shipped with the snapshot); 'port': only the oracle restatement is available."""
    from oracle import reference as oref
    return "reference" if oref.available() else "port"


def cpu_step(frames) -> float:
    """Every frame through every condition on the host: corrupt + fuse + the reference's
    compute_comprehensive_metrics set (IoU, accuracy, ECE, disagreement AUROC).  Returns pixels."""
    import numpy as np
    import torch
    from oracle import reference as oref
    if oref.available():     # the reference's own code (apply_weather_effect / EnsembleModel.forward / RobustnessMetrics)
        return oref.reference_step(frames, CONDITIONS, RAW_WEIGHTS, TEMPERATURE, NUM_CLASSES)
    from oracle import weather as ow, fusion as of_, metrics as om
    raw_w = torch.tensor(RAW_WEIGHTS)
    temp = torch.tensor([TEMPERATURE])
    np.random.seed(42)
    px = 0.0
    for kind in CONDITIONS:
        for img, lab, la, lb in frames:
            ow.apply(img, kind)
            fused = of_.fuse_logits(la, lb, "weighted_average", raw_w, temp)
            om.iou(fused, lab, NUM_CLASSES)
            om.pixel_accuracy(fused, lab)
            om.ece(fused, lab)
            om.disagreement_auroc([la, lb], lab)
            px += float(img.shape[0] * img.shape[1])
    return px


def cpu_sample_text(n_frames, h, w):
    return (f"{n_frames} full {h}x{w} frame(s) per condition x {len(CONDITIONS)} conditions per step: "
            "apply_weather_effect + EnsembleModel fusion + compute_comprehensive_metrics (IoU / accuracy / ECE / AUROC)")


def run_reference_arm(args, rank):
    """The reference's CPU implementation of the path on this box's host cores, FULL-SIZE frames: 2 per condition
    per step when the whole --steps / --warmup run then fits ~5 minutes, else 1."""
    import torch
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    kind = cpu_kind()
    budget_s = 300.0
    one = cpu_sample_inputs(args.height, args.width, 1)
    cpu_step(cpu_sample_inputs(max(32, args.height // 16), args.width, 1))   # imports, thread pools
    t0 = time.perf_counter()
    cpu_step(one)
    t_frame = time.perf_counter() - t0
    total_steps = max(args.steps + args.warmup, 1)
    n_frames = 2 if 2 * t_frame * total_steps <= budget_s else 1
    frames = cpu_sample_inputs(args.height, args.width, n_frames)
    for _ in range(args.warmup):
        cpu_step(frames)
    t0 = time.perf_counter()
    px = 0.0
    for _ in range(args.steps):
        px += cpu_step(frames)
    dt = time.perf_counter() - t0
    value = px / dt / 1e6
    # `config` is the GPU arm's (the driver pairs the two lines on it); what THIS arm ran per step -- full-size frames,
    # fewer of them -- is stated in `sample` / `cpu_baseline.sample`
    cfg = workload_config(args)
    sample = cpu_sample_text(n_frames, args.height, args.width)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
        "config": cfg, "sample": sample, "frames_per_condition_per_step": n_frames,
        "cpu_baseline": {"value": value, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    return {"workload": "configs[1]: all 5 conditions (clean/fog/rain/snow/night) corrupt + 2-member ensemble "
                        "fuse + score, batch %d at %dx%d per GPU" % (args.batch, args.height, args.width),
            "batch_per_gpu": args.batch, "height": args.height, "width": args.width, "num_classes": NUM_CLASSES,
            "members": 2, "labels": "uint8", "fields": "fp64 depth / noise", "strategy": "weighted_average",
            "raw_weights": list(RAW_WEIGHTS), "temperature": TEMPERATURE, "ece_bins": 15, "auroc_bins": AUROC_BINS,
            "l2": "inputs (>=20 GB per step) far exceed the 126 MB L2; no flush needed",
            "parallelism": "frames sharded by rank, one all_reduce(SUM) of int64 bins per step"}


# -------------------------------------------------------------------------------- GPU arm
class Workload:
    """Device-resident inputs of one rank + the step function."""

    def __init__(self, args, rank, device):
        import numpy as np
        import torch
        from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops, _lib
        from adverse_weather_semantic_segmentation_robustness_benchmark_b200.data.preprocessing import (
            WeatherDegradationTransforms)
        from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation.streaming import (
            StreamingEvaluator)
        self.torch, self.ops, self.lib = torch, ops, _lib.load()
        self.args = args
        b, h, w = args.batch, args.height, args.width
        self.b, self.h, self.w = b, h, w
        # the same synthetic frames on every rank (seeded by content, not by rank): weak scaling does not care, and
        # the sharded sweep / probe results are then comparable across world sizes bit for bit
        gen = torch.Generator(device=device).manual_seed(42)
        self.la = torch.randn(b, NUM_CLASSES, h, w, device=device, generator=gen)
        self.lb = torch.randn(b, NUM_CLASSES, h, w, device=device, generator=gen)
        # images / labels the way the reference's synthetic dataset makes them (loader.py:206,231)
        self.images = torch.randint(0, 255, (b, h, w, 3), device=device, generator=gen, dtype=torch.uint8)
        self.labels = torch.randint(0, NUM_CLASSES, (b, h, w), device=device, generator=gen, dtype=torch.uint8)
        self.out = torch.empty_like(self.images)
        self.workspace = ops.corrupt_workspace(b, h, w)
        # host draws with the reference's RNG and order: a pool of `--pool` parameter sets per condition, cycled over
        # the batch; fog / night fields are uploaded slot by slot (a slot's host arrays are dropped at once)
        t = WeatherDegradationTransforms(seed=42)
        self.params, self.fields, self.items = {}, {}, {}
        pool = max(1, min(args.pool, b))
        for kind in CONDITIONS[1:]:
            draws = []
            fld = None
            if kind == "fog":
                fld = torch.empty((b, h, w), dtype=torch.float64, device=device)
            elif kind == "night":
                fld = torch.empty((b, h, w, 3), dtype=torch.float64, device=device)
            for i in range(pool):
                d = t.draw(kind, h, w)
                if kind == "fog":
                    fld[i] = t.synthetic_depth(d.depth_noise[None])[0]     # device fp64 [H,W]
                    d.depth_noise = None
                elif kind == "night":
                    fld[i] = torch.from_numpy(d.noise).to(device)
                    d.noise = None
                draws.append(d)
            if fld is not None:
                for i in range(pool, b):
                    fld[i] = fld[i % pool]
                fld = fld.reshape(-1)
            full = [draws[i % pool] for i in range(b)]
            prm, _, items = t.pack(full, h, w, gather_fields=False)  # fields are assembled on the device
            self.params[kind] = prm
            self.fields[kind] = fld
            self.items[kind] = None if items is None else torch.from_numpy(items).to(device)
        self.ev = StreamingEvaluator(NUM_CLASSES, CONDITIONS, 15, AUROC_BINS, "weighted_average", RAW_WEIGHTS,
                                     TEMPERATURE, ensemble=True)
        self.score_events = []
        self.record_score_events = False

    def step(self, images=None, labels=None, la=None, lb=None, fields=None):
        torch, ops = self.torch, self.ops
        images = self.images if images is None else images
        labels = self.labels if labels is None else labels
        la = self.la if la is None else la
        lb = self.lb if lb is None else lb
        fields = self.fields if fields is None else fields
        self.ev.reset()  # a step is one sweep: score -> all_reduce -> (finalise); the merged bins are the step's result
        for kind in CONDITIONS:
            if kind == "clean":
                # 'clean' aliases its input in the reference (preprocessing.py:78-79): a bare score launch, which is
                # also the one the roofline times (the other four are the same launch behind their corruption kernels,
                # inside ONE C call each -- awx_corrupt_score -- where no event can be placed between the two)
                if self.record_score_events:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    self.ev.update(kind, la, lb, labels)
                    e1.record()
                    self.score_events.append((e0, e1))
                else:
                    self.ev.update(kind, la, lb, labels)
            else:
                self.ev.update_corrupted(kind, images, self.params[kind], fields[kind], self.items[kind], self.out,
                                         self.workspace, la, lb, labels)
        self.ev.all_reduce()


def bind_near_gpu(local_rank, world):
    """Multi-rank runs: pin this process to the CPUs NVML reports as closest to its GPU, so that the pinned
    staging buffers of the e2e leg are allocated on that socket (first touch) and every rank's H2D copies stay
    off the inter-socket link.  Single-rank runs keep all cores (the cpu_baseline leg uses them)."""
    if world <= 1:
        return None
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[local_rank]) if visible and visible.split(",")[local_rank].isdigit() else local_rank
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def run_gpu_arm(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    near_cpus = bind_near_gpu(local_rank, world)
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import _lib
    lib = _lib.load()
    wl = Workload(args, rank, device)
    px_step = float(len(CONDITIONS) * args.batch * args.height * args.width)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident throughput
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(max(args.warmup, 3)):
        wl.step()
    wl.ev.reset()
    barrier()
    sampler.begin_region()
    wl.record_score_events = True
    launches0 = lib.awx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wl.step()
    e1.record()
    barrier()
    sampler.end_region()
    launches = lib.awx_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    wl.record_score_events = False
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = world * px_step / (ms_step * 1e-3) / 1e6
    score_ms = float(np.mean([a.elapsed_time(b) for a, b in wl.score_events]))
    results = wl.ev.finalize()

    # ---- end to end through the public API with host buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, wl, device, world, barrier, max_over_ranks, px_step)
        if near_cpus is not None:
            e2e["cpu_affinity"] = "rank pinned to the %d CPUs NVML reports closest to its GPU" % near_cpus

    # ---- CPU baseline (rank 0, N=1 only): the reference's own code on 2 full frames per condition
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        torch.set_num_threads(os.cpu_count() or 1)
        cpu_step(cpu_sample_inputs(max(32, args.height // 16), args.width, 1))    # imports, thread pools
        frames = cpu_sample_inputs(args.height, args.width, 2)
        t0 = time.perf_counter()
        px = cpu_step(frames)
        dt = time.perf_counter() - t0
        cpu = {"value": px / dt / 1e6, "unit": "Mpixel/s", "cores": torch.get_num_threads(), "kind": cpu_kind(),
               "sample": cpu_sample_text(2, args.height, args.width) + f", {dt:.1f} s"}

    # parity gates of the timed steps' own bins, before the side measurements reuse the evaluator
    parity = parity_object(wl, results, world) if rank == 0 else None
    configs = None
    if not args.no_configs:
        configs = configs_object(wl, args, rank, world, barrier, max_over_ranks)
    shard_ok, shard_digest = sharding_probe(wl, args, rank, world, device)

    if rank == 0:
        peak, peak_src = measured_peak()
        px_launch = float(args.batch * args.height * args.width)
        achieved = SCORE_BYTES_PER_PX * px_launch / (score_ms * 1e-3) / 1e9
        # dram__bytes_read + write of one launch of this kernel at this size, from the newest committed `ncu --set full`
        # capture (round 2's kernel: profiles/r2d_score_g3_summary.json; ncu cannot run inside a timed benchmark)
        traffic, traffic_src = None, None
        for name in ("r2d_score_g3_summary.json", "score_kernel_summary.json"):
            try:
                with open(os.path.join(ROOT, "profiles", name)) as fh:
                    summ = json.load(fh)
                if summ.get("pixels_per_launch") == px_launch:
                    traffic, traffic_src = summ.get("dram_bytes_per_launch"), "profiles/" + name
                    break
            except Exception:
                pass
        step_bytes = sum((SCORE_BYTES_PER_PX + CORRUPT_BYTES_PER_PX[k]) for k in CONDITIONS) * px_launch
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "fp32", "data": "synthetic", "config": workload_config(args),
            "roofline": {"bound": "hbm", "kernel": "score_v2_kernel<weighted,u8 labels,bins only> (awx_score)", "timed_launch": "the clean condition's launch of every timed step (the four others run the same kernel behind their corruption kernels inside awx_corrupt_score)", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": SCORE_BYTES_PER_PX * px_launch,
                         "ms_per_launch": score_ms,
                         "whole_step_GBps": step_bytes / (ms_step * 1e-3) / 1e9},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "parity": parity,
            "configs": configs,
        }
        line["parity"]["sharded_bins_equal_single"] = shard_ok
        line["parity"]["sharding_probe"] = ("16 frames of %dx%d, frame i -> rank i mod %d, all_reduce vs all 16 on rank 0; "
                                            "bins sha256 %s" % (args.height, args.width, world, shard_digest))
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sweep_blocks(total_frames, batch, rank, world):
    """BASELINE configs[4] sharding: the sweep is cut into blocks of `batch` consecutive frames; block k goes to
    rank k mod world and is evaluated under condition k mod 5.  Every block is the rank's (rank-independent) pool of
    frames, so the merged bins depend on the total number of frames only -- not on the number of ranks."""
    n_blocks = (total_frames + batch - 1) // batch
    return [(k, min(batch, total_frames - k * batch)) for k in range(n_blocks) if k % world == rank]


def run_one_sweep(wl, blocks):
    wl.ev.reset()
    for k, n in blocks:
        kind = CONDITIONS[k % len(CONDITIONS)]
        if kind == "clean":
            wl.ev.update(kind, wl.la[:n], wl.lb[:n], wl.labels[:n])
        else:
            wl.ev.update_corrupted(kind, wl.images[:n], wl.params[kind][:n], wl.fields[kind], wl.items[kind], wl.out[:n],
                                   wl.workspace, wl.la[:n], wl.lb[:n], wl.labels[:n])
    wl.ev.all_reduce()       # ONE collective closes the sweep


def bins_digest(ev):
    import hashlib
    return hashlib.sha256(ev.canonical_bins().cpu().numpy().tobytes()).hexdigest()[:16]


def measure_sweep(wl, args, frames, rank, world, barrier, max_over_ranks):
    """One warm-up sweep, one timed sweep (CUDA events, max over ranks).  Strong scaling: `frames` is the total."""
    import torch
    blocks = sweep_blocks(frames, args.batch, rank, world)
    lib = wl.lib
    run_one_sweep(wl, blocks)
    barrier()
    launches0 = lib.awx_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_one_sweep(wl, blocks)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    res = wl.ev.finalize()
    px = float(frames) * args.height * args.width
    step_bytes = sum((SCORE_BYTES_PER_PX + CORRUPT_BYTES_PER_PX[CONDITIONS[k % 5]]) * n * args.height * args.width
                     for k, n in sweep_blocks(frames, args.batch, 0, 1))
    peak, _ = measured_peak()
    return {"frames": frames, "ms": ms, "frames_per_s": frames / (ms * 1e-3), "Mpixel_per_s": px / (ms * 1e-3) / 1e6,
            "GBps_all_gpus": step_bytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak_per_gpu": step_bytes / (ms * 1e-3) / 1e9 / peak / world,
            "gpu_launches_this_rank": int(lib.awx_launch_count() - launches0), "blocks_this_rank": len(blocks),
            "bins_sha256_16": bins_digest(wl.ev),
            "results": {k: (float(v) if v == v else None) for k, v in res.items()}}


def run_sweep(args, rank, world, local_rank):
    """BASELINE configs[4] as its own run: `--sweep-frames N` frames in total, sharded by block over the ranks, one
    all_reduce of the packed bins at the end.  `bins_sha256_16` / `results` must not depend on the number of ranks."""
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    wl = Workload(args, rank, device)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    run_one_sweep(wl, sweep_blocks(min(args.sweep_frames, 4 * args.batch * world), args.batch, rank, world))   # warm-up
    barrier()
    sampler.begin_region()
    m = measure_sweep(wl, args, args.sweep_frames, rank, world, barrier, max_over_ranks)
    sampler.end_region()
    clocks = sampler.stop() if rank == 0 else None
    if rank == 0:
        cfg = workload_config(args)
        cfg["workload"] = ("configs[4]: %d-frame synthetic Cityscapes eval sweep in blocks of %d frames, block k -> rank k mod %d "
                           "under condition k mod 5, one all_reduce of the bins" % (args.sweep_frames, args.batch, world))
        print(json.dumps({
            "metric": METRIC, "value": m["Mpixel_per_s"], "unit": "Mpixel/s", "n_gpus": world, "steps": 1, "warmup": 1,
            "ms_per_step": m["ms"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "fp32",
            "data": "synthetic", "config": cfg, "frames": args.sweep_frames, "frames_per_s": m["frames_per_s"],
            "gpu_launches": m["gpu_launches_this_rank"], "clocks": clocks, "results": m["results"],
            "bins_sha256_16": m["bins_sha256_16"]}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def time_launches(fn, n=20):
    """Device time (ms) per call of `n` back-to-back calls on the current stream, after 3 warm-ups.  The calls are
    enqueued behind a ~50 ms spin kernel, so the host (Python + ctypes, tens of microseconds per call) runs ahead and
    the GPU executes them without gaps: for launches of 0.1-0.5 ms, events around each single call would time the
    host, not the device."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    try:
        torch.cuda._sleep(100_000_000)
    except Exception:      # private helper: without it the first calls may still see a host-side gap
        pass
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def configs_object(wl, args, rank, world, barrier, max_over_ranks):
    """BASELINE configs[2..4] beside the headline (configs[1]): CUDA-event timings on this rank's GPU with the same
    library, so that the loss kernel's and the sweep's numbers are driver-visible.
      c2_fusion_frame  one 19x1024x2048 frame per member: fuse (weighted, /T) + softmax + JS map + ECE + AUROC, the
                       fused logits and the JS map materialised (153 B/px read + 80 written)
      c3_loss_fwd_bwd  fog-density-aware loss forward + backward, batch 8, depth prediction + target, int64 labels
                       (176 B/px)
      c4_sweep         the 10 000-frame sweep sharded over the ranks (total frames fixed: strong scaling)"""
    import torch
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import ops_loss
    peak, _ = measured_peak()
    h, w = args.height, args.width
    ev = wl.ev
    out = {}
    bins2 = wl.ops.new_bins(NUM_CLASSES, 15, AUROC_BINS)
    px2 = float(h * w)
    ms = time_launches(lambda: wl.ops.score(wl.la[:1], wl.lb[:1], wl.labels[:1], strategy=ev.strategy, w0=ev.w0, w1=ev.w1,
                                            temperature=ev.temperature, auroc_bins=AUROC_BINS, bins=bins2,
                                            want_fused=True, want_js=True))
    out["c2_fusion_frame"] = {"ms": ms, "bytes_per_px": 233.0, "GBps": 233.0 * px2 / (ms * 1e-3) / 1e9,
                              "frac_of_measured_peak": 233.0 * px2 / (ms * 1e-3) / 1e9 / peak,
                              "note": "one 2 Mpx frame per launch (148 CTAs x 29 tiles: ring fill, histogram flush and tail are not amortised); fused logits + JS map written; 20 launches back to back"}
    nb = min(8, args.batch)
    gen = torch.Generator(device=wl.la.device).manual_seed(7)
    lab64 = wl.labels[:nb].long()
    fd = torch.rand(nb, h, w, device=wl.la.device, generator=gen)
    dpred = torch.rand(nb, 1, h, w, device=wl.la.device, generator=gen) * 50
    dtgt = torch.rand(nb, h, w, device=wl.la.device, generator=gen) * 50
    px3 = float(nb * h * w)
    ms = time_launches(lambda: ops_loss.fogloss_raw(wl.la[:nb], lab64, fd, dpred, dtgt, 2.0, False, True))
    out["c3_loss_fwd_bwd"] = {"ms": ms, "batch": nb, "bytes_per_px": 176.0, "GBps": 176.0 * px3 / (ms * 1e-3) / 1e9,
                              "frac_of_measured_peak": 176.0 * px3 / (ms * 1e-3) / 1e9 / peak,
                              "note": "includes the wrapper's output allocations; kernel: fogloss_ring_kernel"}
    del fd, dpred, dtgt, lab64
    m = measure_sweep(wl, args, 10000, rank, world, barrier, max_over_ranks)
    m.pop("results")
    out["c4_sweep"] = m
    # the corruption kernels on their own (awx_corrupt on the step's batch and parameter pool; rain / snow include the
    # overlay rasterisation and the mask memset): algorithmic bytes per pixel as in SURVEY 8d
    cor = {}
    pxb = float(args.batch * h * w)
    for kind in CONDITIONS[1:]:
        ms = time_launches(lambda: wl.ops.corrupt(wl.images, wl.params[kind], wl.fields[kind], wl.items[kind], out=wl.out,
                                                  workspace=wl.workspace))
        gbps = CORRUPT_BYTES_PER_PX[kind] * pxb / (ms * 1e-3) / 1e9
        cor[kind] = {"ms": ms, "bytes_per_px": CORRUPT_BYTES_PER_PX[kind], "GBps": gbps, "frac_of_measured_peak": gbps / peak}
    cor["kernels"] = {"fog": "fog_kernel (fp32 screen + fp64 exact path)", "rain": "blur_strip_kernel<1, rain>",
                      "snow": "blur_strip_kernel<1 | 3, snow> (the pool's 3- and 7-tap frames in two launches)",
                      "night": "pointwise_kernel<double>"}
    out["corruption_batch"] = cor
    return out


def sharding_probe(wl, args, rank, world, device):
    """On-hardware proof that sharded bins do not depend on the number of ranks: a fixed probe of 16 full-size frames
    (content seeded by the frame's index, condition = index mod 5) is scored with frame i on rank i mod world and
    merged by the all_reduce; rank 0 also scores all 16 alone.  The two buffers must be identical."""
    import torch
    from adverse_weather_semantic_segmentation_robustness_benchmark_b200.evaluation.streaming import StreamingEvaluator
    h, w = args.height, args.width
    ev = wl.ev
    make = lambda: StreamingEvaluator(NUM_CLASSES, CONDITIONS, 15, AUROC_BINS, "weighted_average", RAW_WEIGHTS,
                                      TEMPERATURE, ensemble=True, bins_device=device)
    shard, single = make(), make()

    def frame(i):
        gen = torch.Generator(device=device).manual_seed(7000 + i)
        la = torch.randn(1, NUM_CLASSES, h, w, device=device, generator=gen)
        lb = torch.randn(1, NUM_CLASSES, h, w, device=device, generator=gen)
        lab = torch.randint(0, NUM_CLASSES, (1, h, w), device=device, generator=gen, dtype=torch.uint8)
        lab[0, : 1 + i % 3] = 255
        return la, lb, lab

    for i in range(16):
        if i % world == rank or rank == 0:
            la, lb, lab = frame(i)
            if i % world == rank:
                shard.update(CONDITIONS[i % 5], la, lb, lab)
            if rank == 0:
                single.update(CONDITIONS[i % 5], la, lb, lab)
    shard.all_reduce()
    return bool(torch.equal(shard.canonical_bins(), single.canonical_bins())), bins_digest(single)


def parity_object(wl, results, world):
    """Parity gates reported with the benchmark: the step's own counters, a small probe of the
    hot path against the CPU oracle (N = 1 only: the oracle is the checker here, never the thing measured), and
    the versions of the libraries whose arithmetic the oracle calls."""
    per = wl.ev.per_condition()
    out = {"overall_miou": results.get("overall_miou"), "step_pixels": int(sum(v["pixels"] for v in per.values()))}
    for key in ("ece_ambiguous_pixels", "epred_ambiguous_pixels", "marg_ambiguous_pixels"):
        out[key] = int(sum(v.get(key, 0) for v in per.values()))
    if world == 1:
        import torch
        import __graft_entry__ as entry
        try:
            # ONE FULL FRAME of the step's own inputs through the kernel the step launches (bins only), against the
            # oracle on the host: true mismatch of every integer output next to the reported ambiguous pixels
            ev = wl.ev
            got = wl.ops.read_bins(wl.ops.score(wl.la[:1], wl.lb[:1], wl.labels[:1], strategy=ev.strategy, w0=ev.w0, w1=ev.w1,
                                                temperature=ev.temperature, auroc_bins=ev.auroc_bins)["bins"],
                                   NUM_CLASSES, 15, ev.auroc_bins)
            full = entry.bins_vs_oracle(got, wl.la[:1].cpu(), wl.lb[:1].cpu(), wl.labels[:1].cpu(),
                                        torch.tensor(RAW_WEIGHTS), TEMPERATURE, NUM_CLASSES)
            out["full_frame"] = full
            out["ece_true_mismatch"] = full["ece_true_mismatch"]
            out["ens_wrong_mismatch"] = full["ens_wrong_mismatch"]
        except Exception as exc:  # a failed gate must be visible in the line, not kill the measurement
            out["full_frame"] = {"failed": repr(exc)[:300]}
        try:
            out["oracle_probe"] = entry.parity_probe()
        except Exception as exc:
            out["oracle_probe"] = {"failed": repr(exc)[:300]}
    try:
        import cv2, numpy, scipy, sklearn, torch
        out["versions"] = {"numpy": numpy.__version__, "cv2": cv2.__version__, "scipy": scipy.__version__,
                           "torch": torch.__version__, "sklearn": sklearn.__version__}
    except Exception:
        pass
    return out


def run_e2e(args, wl, device, world, barrier, max_over_ranks, px_step):
    """Same step, but every input starts in (pinned) host memory and the bins are read back."""
    import torch

    def pin(t):
        host = torch.empty(t.shape, dtype=t.dtype, device="cpu", pin_memory=True)
        host.copy_(t)
        return host

    # every rank stages its whole step's inputs in host memory (25 GB at the default size): refuse
    # rather than push the box into swap / the OOM killer when the ranks together would not fit
    need = sum(t.numel() * t.element_size() for t in (wl.images, wl.labels, wl.la, wl.lb))
    need += sum(v.numel() * v.element_size() for v in wl.fields.values() if v is not None)
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = None
    if avail is not None and need * world * 1.25 > avail:
        return {"value": None, "unit": "Mpixel/s", "skipped": "host staging buffers (%.1f GB x %d ranks) exceed the "
                "available host memory (%.1f GB)" % (need / 1e9, world, avail / 1e9)}

    try:
        h_images, h_labels = pin(wl.images), pin(wl.labels)
        h_la, h_lb = pin(wl.la), pin(wl.lb)
        h_fields = {k: (None if v is None else pin(v)) for k, v in wl.fields.items()}
        pinned = True
    except RuntimeError:
        h_images, h_labels, h_la, h_lb = wl.images.cpu(), wl.labels.cpu(), wl.la.cpu(), wl.lb.cpu()
        h_fields = {k: (None if v is None else v.cpu()) for k, v in wl.fields.items()}
        pinned = False
    h2d = sum(t.numel() * t.element_size() for t in (h_images, h_labels, h_la, h_lb))
    h2d += sum(v.numel() * v.element_size() for v in h_fields.values() if v is not None)
    h2d += sum(p.nbytes for p in wl.params.values())
    h2d += sum(v.numel() * v.element_size() for v in wl.items.values() if v is not None)
    d2h = wl.ev.bins.numel() * 8
    # device staging buffers are reused (the copies below overwrite them every step)
    d_images, d_labels, d_la, d_lb = wl.images, wl.labels, wl.la, wl.lb
    d_fields = wl.fields

    # The batch goes over in CHUNKS on a copy stream while the default stream corrupts + scores the chunks that
    # have landed (all five conditions per chunk): H2D and compute overlap, the step is bound by PCIe alone.
    n_chunks = 8 if args.batch % 8 == 0 and args.batch >= 8 else 1
    cb = args.batch // n_chunks
    hw = args.height * args.width
    copy_stream = torch.cuda.Stream(device=device)
    ready = [torch.cuda.Event() for _ in range(n_chunks)]
    done = [torch.cuda.Event() for _ in range(n_chunks)]
    per_frame = {"fog": hw, "night": hw * 3}
    main = torch.cuda.current_stream(device)

    def e2e_step():
        wl.ev.reset()
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            for k in range(n_chunks):
                sl = slice(k * cb, (k + 1) * cb)
                copy_stream.wait_event(done[k])        # the previous step's kernels are finished with this slice
                d_images[sl].copy_(h_images[sl], non_blocking=True)
                d_labels[sl].copy_(h_labels[sl], non_blocking=True)
                d_la[sl].copy_(h_la[sl], non_blocking=True)
                d_lb[sl].copy_(h_lb[sl], non_blocking=True)
                for kind, v in h_fields.items():
                    if v is not None:
                        fs = slice(k * cb * per_frame[kind], (k + 1) * cb * per_frame[kind])
                        d_fields[kind][fs].copy_(v[fs], non_blocking=True)
                ready[k].record(copy_stream)
        for k in range(n_chunks):
            sl = slice(k * cb, (k + 1) * cb)
            main.wait_event(ready[k])
            for kind in CONDITIONS:
                if kind == "clean":
                    wl.ev.update(kind, d_la[sl], d_lb[sl], d_labels[sl])
                else:
                    wl.ev.update_corrupted(kind, d_images[sl], wl.params[kind][sl], d_fields[kind], wl.items[kind],
                                           wl.out[sl], wl.workspace, d_la[sl], d_lb[sl], d_labels[sl])
            done[k].record(main)
        wl.ev.all_reduce()
        return wl.ev.bins.cpu()      # D2H read of the step's result

    steps = max(1, min(args.steps, 3))
    e2e_step()
    wl.ev.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / steps)

    # the same copies with no kernel behind them: if the step takes what its copies alone take, the leg is bound by
    # the host -> device path (PCIe link at N = 1, the host's memory / IO fabric when several ranks copy at once)
    def copies_only():
        copy_stream.wait_stream(main)
        with torch.cuda.stream(copy_stream):
            for k in range(n_chunks):
                sl = slice(k * cb, (k + 1) * cb)
                d_images[sl].copy_(h_images[sl], non_blocking=True)
                d_labels[sl].copy_(h_labels[sl], non_blocking=True)
                d_la[sl].copy_(h_la[sl], non_blocking=True)
                d_lb[sl].copy_(h_lb[sl], non_blocking=True)
                for kind, v in h_fields.items():
                    if v is not None:
                        fs = slice(k * cb * per_frame[kind], (k + 1) * cb * per_frame[kind])
                        d_fields[kind][fs].copy_(v[fs], non_blocking=True)
        main.wait_stream(copy_stream)

    barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    copies_only()
    c1.record()
    barrier()
    copy_ms = max_over_ranks(c0.elapsed_time(c1))
    # the pipelined, chunked step must produce exactly the (merged) bins of one plain device-resident step
    got = wl.ev.canonical_bins()
    wl.ev.reset()
    wl.step()
    torch.cuda.synchronize()
    bins_match = bool(torch.equal(got, wl.ev.canonical_bins()))
    return {"value": world * px_step / (ms * 1e-3) / 1e6, "bins_match_device_resident_step": bins_match, "unit": "Mpixel/s", "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "ms_per_step": ms, "steps": steps, "pinned": pinned,
            "h2d_GBps_per_gpu": h2d / (ms * 1e-3) / 1e9, "h2d_GBps_all_gpus": world * h2d / (ms * 1e-3) / 1e9,
            "copies_only_ms": copy_ms, "copies_only_GBps_per_gpu": h2d / (copy_ms * 1e-3) / 1e9,
            "bound": ("host->device copies (the step takes %.2fx what its copies alone take; tools/h2d_probe.py, "
                      "profiles/r2_h2d_probe_*.json)" % (ms / copy_ms)) if ms < 1.25 * copy_ms else "compute",
            "api": "StreamingEvaluator.update_corrupted (awx_corrupt_score via the C ABI; 'clean': update / awx_score)",
            "pipeline": "%d chunks of %d frames: H2D on a copy stream overlapped with corrupt + score" % (n_chunks, cb)}


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # convenience: re-launch under torchrun when asked for several GPUs from a plain shell
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29531", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.sweep_frames > 0:
        run_sweep(args, rank, world, local_rank)
        return
    run_gpu_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
