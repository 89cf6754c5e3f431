"""Per-kernel table from an ncu metrics CSV taken over whole steps of bench.py:

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,\
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none --csv --log-file X.csv ...

    python tools/kernel_table.py X.csv <steps captured> [min share %] [json out]

Columns: launches / step, device time / step and its share, DRAM bytes read + written / step,
achieved DRAM GB/s (bytes / time), time-weighted tensor-pipe % and DRAM-throughput %.  ncu replays
every launch alone with cold caches: compare shares and per-kernel rates, not the absolute total."""
import collections
import csv
import json
import re
import sys

path, steps = sys.argv[1], float(sys.argv[2])
min_share = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
lines = [l for l in open(path) if not l.startswith("==")]
per_launch = collections.OrderedDict()
for row in csv.DictReader(lines):
    key = (row["ID"], row["Kernel Name"])
    v = float(row["Metric Value"].replace(",", "") or 0)
    unit, name = row["Metric Unit"], row["Metric Name"]
    if name == "gpu__time_duration.sum":
        v = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0) * v
    elif name.startswith("dram__bytes"):
        v = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0) * v
    per_launch.setdefault(key, {})[name] = v
agg = collections.OrderedDict()
for (_, kname), m in per_launch.items():
    k = re.sub(r"\(.*", "", kname).replace("void ", "").replace("b2n::", "")
    a = agg.setdefault(k, collections.Counter())
    t = m.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["us"] += t
    a["rd"] += m.get("dram__bytes_read.sum", 0.0)
    a["wr"] += m.get("dram__bytes_write.sum", 0.0)
    a["tensor_w"] += t * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 0.0)
    a["dram_w"] += t * m.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 0.0)
tot = sum(a["us"] for a in agg.values())
print("| kernel | launches/step | us/step | share | DRAM read GB/step | DRAM write GB/step | achieved GB/s | tensor pipe % | DRAM throughput % |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|")
out = {}
for k, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    share = 100 * a["us"] / tot
    rec = {"launches_per_step": a["n"] / steps, "us_per_step": a["us"] / steps, "share_pct": share,
           "dram_read_bytes": a["rd"] / steps, "dram_write_bytes": a["wr"] / steps,
           "achieved_gbs": (a["rd"] + a["wr"]) / (a["us"] * 1e-6) / 1e9 if a["us"] else 0.0,
           "tensor_pipe_pct": a["tensor_w"] / a["us"] if a["us"] else 0.0,
           "dram_throughput_pct": a["dram_w"] / a["us"] if a["us"] else 0.0}
    out[k] = rec
    if share >= min_share:
        print("| `%s` | %.1f | %.0f | %.1f %% | %.2f | %.2f | %.0f | %.1f | %.1f |" % (
            k[:72], rec["launches_per_step"], rec["us_per_step"], share, rec["dram_read_bytes"] / 1e9,
            rec["dram_write_bytes"] / 1e9, rec["achieved_gbs"], rec["tensor_pipe_pct"], rec["dram_throughput_pct"]))
print("| **all kernels** | %.0f | %.0f | 100 %% | %.2f | %.2f | | | |" % (
    sum(a["n"] for a in agg.values()) / steps, tot / steps, sum(a["rd"] for a in agg.values()) / steps / 1e9,
    sum(a["wr"] for a in agg.values()) / steps / 1e9))
if len(sys.argv) > 4:
    json.dump(out, open(sys.argv[4], "w"), indent=1)
