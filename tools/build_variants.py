"""Build experimental variants of libawx.so with extra -D flags (dev tool).
usage: build_variants.py name1="-DA=1 -DB=2" name2="..."   ->  build/variants/libawx_<name>.so
Select one at run time with AWX_LIB=build/variants/libawx_<name>.so."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import build as b

out_dir = os.path.join(b.ROOT, "build", "variants")
os.makedirs(out_dir, exist_ok=True)
procs = []
for arg in sys.argv[1:]:
    name, flags = arg.split("=", 1)
    out = os.path.join(out_dir, f"libawx_{name}.so")
    cmd = [b._nvcc()] + b.NVCC_FLAGS + flags.split() + ["-I", b.INCLUDE, "-I", b.CSRC] + b.sources() + ["-o", out]
    procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
for name, pr in procs:
    log, _ = pr.communicate()
    print(name, "rc", pr.returncode, [l for l in log.splitlines() if "error" in l][:5])
