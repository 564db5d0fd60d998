"""Static SASS opcode histogram of the kernels of libawx.so whose (mangled) name contains a substring (dev tool).
usage: sass_hist.py <substring> [--dump DIR]   e.g.  sass_hist.py blur_strip_kernel"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "adverse_weather_semantic_segmentation_robustness_benchmark_b200", "libawx.so")
sub = sys.argv[1]
dump = sys.argv[sys.argv.index("--dump") + 1] if "--dump" in sys.argv else None
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
name, funcs = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        funcs[name] = []
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(.*?)\s*;?\s*/\*", line)
    if m and name:
        funcs[name].append(m.group(1))
for name, ins in funcs.items():
    if sub not in name:
        continue
    short = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()[:110]
    hist = collections.Counter()
    for i in ins:
        parts = i.split()
        op = parts[1] if parts[0].startswith("@") else parts[0]
        hist[op.split(".")[0]] += 1
    print(f"== {short}\n   {len(ins)} instructions: " + ", ".join(f"{k} {v}" for k, v in hist.most_common(22)))
    if dump:
        os.makedirs(dump, exist_ok=True)
        with open(os.path.join(dump, re.sub(r"\W+", "_", short)[:80] + ".sass"), "w") as fh:
            fh.write("\n".join(ins))
