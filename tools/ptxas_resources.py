"""Regenerate profiles/r1_ptxas_resources.md: `-Xptxas -v` of every translation unit of libawx.so (dev tool)."""
import os, re, subprocess, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from adverse_weather_semantic_segmentation_robustness_benchmark_b200 import build as b

rows = {}
procs = []
tmp = tempfile.mkdtemp()
flags = [f for f in b.NVCC_FLAGS if f not in ("-shared",)]
for src in b.sources():
    obj = os.path.join(tmp, os.path.basename(src) + ".o")
    cmd = [b._nvcc()] + flags + ["-Xptxas", "-v", "-c", "-I", b.INCLUDE, "-I", b.CSRC, src, "-o", obj]
    procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
for pr in procs:
    log, _ = pr.communicate()
    name = None
    for line in log.splitlines():
        m = re.search(r"Compiling entry function '(\w+)'", line)
        if m:
            name = m.group(1)
            rows[name] = {"regs": 0, "stack": None, "spill": 0, "smem": 0}
            continue
        if name is None:
            continue
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores", line)
        if m and rows[name]["stack"] is None:  # the entry function's own frame (device functions it calls follow)
            rows[name]["stack"], rows[name]["spill"] = int(m.group(1)), int(m.group(2))
        m = re.search(r"Used (\d+) registers", line)
        if m:
            rows[name]["regs"] = int(m.group(1))
            s = re.search(r"(\d+) bytes smem", line)
            rows[name]["smem"] = int(s.group(1)) if s else 0
names = subprocess.run(["c++filt"], input="\n".join(rows), capture_output=True, text=True).stdout.splitlines()


def short(n):
    n = n.replace("void ", "").replace("awx::(anonymous namespace)::", "").replace("(anonymous namespace)::", "")
    n = n.replace("awx::", "").replace("(bool)1", "true").replace("(bool)0", "false").replace("(int)", "")
    return re.sub(r"\(.*", "", n)


out = ["# ptxas resource usage of every kernel in libawx.so (sm_100a, round 2)", "",
       "`nvcc -gencode arch=compute_100a,code=sm_100a -O3 -Xptxas -v` (`tools/ptxas_resources.py`); static shared memory only "
       "(the score / blur / temperature kernels add dynamic shared memory at launch).", "",
       "| kernel | registers | stack B | spill stores B | static smem B |", "|---|---|---|---|---|"]
for n, (k, r) in sorted(zip(map(short, names), rows.items())):
    out.append(f"| `{n}` | {r['regs']} | {r['stack']} | {r['spill']} | {r['smem']} |")
open(os.path.join(b.ROOT, "profiles", os.environ.get("AWX_PTXAS_OUT", "r2_ptxas_resources.md")), "w").write("\n".join(out) + "\n")
print(len(rows), "kernels")
