"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
lines = [l for l in open(path) if not l.startswith("==")]
agg, tot = collections.OrderedDict(), 0.0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(row["Metric Value"].replace(",", ""))
    v = v / 1000 if row["Metric Unit"] == "ns" else (v * 1000 if row["Metric Unit"] == "ms" else v)
    k = re.sub(r"\(.*", "", row["Kernel Name"])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    tot += v
print("| kernel | launches | total us | us/step | share |\n|---|---:|---:|---:|---:|")
for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1])[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]:
    print("| `%s` | %d | %.0f | %.0f | %.1f %% |" % (k[:88], n, t, t / steps, 100 * t / tot))
print("| **all kernels** | %d | %.0f | %.0f | 100 %% |" % (sum(a[0] for a in agg.values()), tot, tot / steps))
