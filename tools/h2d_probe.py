"""Copy-only probe of the end-to-end leg's limiter: pinned host -> device bandwidth per GPU and in aggregate.

    python tools/h2d_probe.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29541 \
        tools/h2d_probe.py [--gib 8] [--chunk-mib 1280]         # N ranks copying at the same time

Every rank stages `--gib` GiB in host memory and copies it to its GPU in chunks of `--chunk-mib` MiB (bench.py's
e2e leg moves 1.27 GB per logits chunk), all ranks starting together after a barrier; device-timed, max over ranks.
Variants, to separate the PCIe link from the host side:
  pinned          torch pin_memory (cudaHostAlloc default), one copy stream           <- what bench.py's e2e leg uses
  pinned_2streams the same buffer, even / odd chunks on two streams
  write_combined  cudaHostAlloc(cudaHostAllocWriteCombined): no CPU cache snooping on the way out
  near_cpu        pinned, allocated after binding the process to the CPUs NVML reports nearest the GPU
  pageable        plain malloc'd memory (the driver stages through its own pinned bounce buffers)
One JSON line per rank-0 run: GB/s per GPU (slowest rank) and aggregate for every variant.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import sys

import torch
import torch.distributed as dist


def alloc_write_combined(nbytes):
    """cudaHostAlloc(..., cudaHostAllocWriteCombined) wrapped as a uint8 tensor (freed by the caller)."""
    rt = ctypes.CDLL("libcudart.so")
    ptr = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(ptr), ctypes.c_size_t(nbytes), ctypes.c_uint(0x04))
    if rc != 0:
        raise RuntimeError(f"cudaHostAlloc(write combined) failed: {rc}")
    buf = (ctypes.c_uint8 * nbytes).from_address(ptr.value)
    t = torch.frombuffer(buf, dtype=torch.uint8)
    return t, (lambda: rt.cudaFreeHost(ptr))


def bind_near_gpu(local_rank):
    try:
        import pynvml
        pynvml.nvmlInit()
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank))
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gib", type=float, default=8.0)
    ap.add_argument("--chunk-mib", type=int, default=1280)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nbytes = int(args.gib * (1 << 30))
    chunk = args.chunk_mib << 20
    dst = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(src, two_streams=False):
        best = None
        for _ in range(args.repeats + 1):          # first pass = warm-up (page faults, driver staging buffers)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            s1.wait_stream(torch.cuda.current_stream())
            s2.wait_stream(torch.cuda.current_stream())
            for i, off in enumerate(range(0, nbytes, chunk)):
                st = s2 if (two_streams and i % 2) else s1
                with torch.cuda.stream(st):
                    dst[off:off + chunk].copy_(src[off:off + chunk], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s1)
            torch.cuda.current_stream().wait_stream(s2)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            if world > 1:
                t = torch.tensor([ms], device=dev, dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t.item())
            best = ms if best is None else min(best, ms)
        return nbytes / (best * 1e-3) / 1e9

    out = {}
    pinned = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    pinned.fill_(1)
    out["pinned"] = run(pinned)
    out["pinned_2streams"] = run(pinned, two_streams=True)
    del pinned
    try:
        wc, free = alloc_write_combined(nbytes)
        out["write_combined"] = run(wc)
        del wc
        free()
    except Exception as exc:
        out["write_combined"] = repr(exc)[:120]
    pageable = torch.empty(nbytes, dtype=torch.uint8)
    pageable.fill_(1)
    out["pageable"] = run(pageable)
    del pageable
    near = bind_near_gpu(local)
    if near is not None:
        p2 = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        p2.fill_(1)
        out["near_cpu"] = run(p2)
        out["near_cpu_count"] = near
        del p2
    if rank == 0:
        line = {"probe": "pinned host -> device copies, all ranks at once", "n_gpus": world, "gib_per_rank": args.gib,
                "chunk_mib": args.chunk_mib, "host_cpus": os.cpu_count(),
                "GBps_per_gpu_slowest_rank": {k: (round(v, 2) if isinstance(v, float) else v) for k, v in out.items()},
                "GBps_aggregate": {k: round(v * world, 1) for k, v in out.items() if isinstance(v, float)}}
        try:
            import psutil
            line["host_mem_gb"] = round(psutil.virtual_memory().total / 1e9, 1)
        except Exception:
            pass
        try:
            import subprocess
            line["numa_nodes"] = subprocess.run("ls -d /sys/devices/system/node/node* | wc -l", shell=True, capture_output=True,
                                                text=True).stdout.strip()
            line["gpu_topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout[-1500:]
        except Exception:
            pass
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    sys.exit(main())
