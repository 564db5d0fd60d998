"""Build libawx.so (hand-written sm_100a CUDA + C ABI) in-tree with nvcc.

    python -m adverse_weather_semantic_segmentation_robustness_benchmark_b200.build [--force] [--verbose]

The library lands next to this file so that it travels with the repo snapshot to the GPU box;
there is no JIT cache and no pip install.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB_PATH = os.path.join(PKG_DIR, "libawx.so")
STAMP = os.path.join(PKG_DIR, "libawx.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--fmad=true",          # exact paths use __f*_rn intrinsics explicitly
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
    "-shared",
    "--threads", "0",       # one ptxas/cicc job per source file
]


def _nvcc() -> str:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(cand):
        raise RuntimeError("nvcc not found; libawx.so cannot be built")
    return cand


def sources() -> list:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest() -> str:
    h = hashlib.sha256()
    files = sources() + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    files += sorted(os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE))
    for f in files:
        h.update(os.path.relpath(f, ROOT).encode())   # repo-relative: the stamp is valid in any checkout dir
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current() -> bool:
    if not (os.path.exists(LIB_PATH) and os.path.exists(STAMP)):
        return False
    with open(STAMP) as fh:
        return fh.read().strip() == _digest()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile every csrc/*.cu into libawx.so for sm_100a; returns the library path."""
    if not force and is_current():
        return LIB_PATH
    cmd = [_nvcc()] + NVCC_FLAGS + ["-I", INCLUDE, "-I", CSRC]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    cmd += sources() + ["-o", LIB_PATH]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed building libawx.so (see stderr)")
    with open(STAMP, "w") as fh:
        fh.write(_digest())
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
