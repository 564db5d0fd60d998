"""Summarise an .ncu-rep (raw page + per-opcode executed-instruction histogram). Dev tool.
usage: ncu_summary.py report.ncu-rep pixels_per_launch [out.json]"""
import csv, collections, io, json, re, subprocess, sys

rep, px = sys.argv[1], float(sys.argv[2])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
get = lambda k: next((vals[i] for i, h in enumerate(hdr) if h == k), None)
keys = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__cycles_elapsed.avg.per_second", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio"]
out = {}
for k in keys:
    v = get(k)
    if v is not None:
        out[k] = (v, units[hdr.index(k)])
stalls = {}
for i, h in enumerate(hdr):
    m = re.match(r"smsp__average_warps_issue_stalled_(\w+)_per_issue_active.ratio", h)
    if m:
        stalls[m.group(1)] = float(vals[i])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
sh = srows[1]
iS, iE = sh.index("Source"), sh.index("Instructions Executed")
ops, tot = collections.Counter(), 0
for r in srows[2:]:
    if len(r) <= iE:
        continue
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_.]+)", r[iS].strip())
    op = m.group(2).split(".")[0] if m else r[iS][:10]
    n = int(r[iE]); ops[op] += n; tot += n
dur_ms = float(out["gpu__time_duration.sum"][0]) * (1e-3 if out["gpu__time_duration.sum"][1] in ("us", "usecond") else 1.0)
if out["gpu__time_duration.sum"][1] in ("ns", "nsecond"):
    dur_ms = float(out["gpu__time_duration.sum"][0]) * 1e-6
def gb(k):
    v, u = out[k]; v = float(v)
    return v * {"Gbyte": 1.0, "Mbyte": 1e-3, "Kbyte": 1e-6, "byte": 1e-9, "Tbyte": 1e3}.get(u, 1.0)
dram = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
summary = {"report": rep, "pixels_per_launch": px, "duration_ms": dur_ms, "dram_bytes_per_launch": dram * 1e9,
           "dram_bytes_per_pixel": dram * 1e9 / px, "dram_GBps": dram / (dur_ms * 1e-3),
           "thread_instr_per_pixel": tot * 32 / px, "metrics": {k: v[0] for k, v in out.items()},
           "stall_warps_per_issue": dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8]),
           "opcodes_thread_instr_per_pixel": {k: round(v * 32 / px, 1) for k, v in ops.most_common(28)}}
print(json.dumps(summary, indent=1))
if len(sys.argv) > 3:
    json.dump(summary, open(sys.argv[3], "w"), indent=1)
