"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (dev tool).
usage: launch_summary.py launches.csv out.json "<description of the profiled command>" """
import collections, csv, json, sys

src, dst, desc = sys.argv[1], sys.argv[2], sys.argv[3]
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr = rows[0]
iK, iM, iU, iV = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value")
scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
agg = collections.OrderedDict()
for r in rows[1:]:
    if r[iM] != "gpu__time_duration.sum":
        continue
    ms = float(r[iV].replace(",", "")) * scale[r[iU]]
    k = r[iK][:120]
    n, t = agg.get(k, (0, 0.0))
    agg[k] = (n + 1, t + ms)
total = sum(t for _, t in agg.values())
out = {"source": f"{src} ({desc})", "total_kernel_ms": round(total, 3),
       "kernels": [{"kernel": k, "launches": n, "total_ms": round(t, 3), "avg_ms": round(t / n, 4), "share": round(t / total, 4)}
                   for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
json.dump(out, open(dst, "w"), indent=1)
for e in out["kernels"][:8]:
    print(e["share"], e["avg_ms"], e["launches"], e["kernel"][:90])
