"""Top SASS instructions by stall samples from `ncu --page source --csv --print-source sass`."""
import csv
import subprocess
import sys

rep, which = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                     capture_output=True, text=True).stdout.splitlines()
# split into per-kernel sections
secs, cur = [], None
for ln in out:
    if ln.startswith('"Kernel Name"'):
        cur = []
        secs.append(cur)
    elif cur is not None:
        cur.append(ln)
rows = list(csv.reader(secs[which]))
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[1:]
tot = sum(int(r[ix["# Samples"]]) for r in data)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("total samples", tot)
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:int(sys.argv[3]) if len(sys.argv) > 3 else 30]
for i in sorted(top):
    r = data[i]
    st = sorted(((int(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
    print("%5d %6d %5.1f%% exec=%-9s %-70s %s" % (i, int(r[ix["# Samples"]]), 100.0 * int(r[ix["# Samples"]]) / tot,
                                          r[ix["Instructions Executed"]], r[ix["Source"]].strip()[:70], st))

# histogram of samples over the instruction stream (buckets of 250 SASS instructions)
B = 250
print("bucket  samples  executed(warp-instr)")
for b in range(0, len(data), B):
    sm = sum(int(r[ix["# Samples"]]) for r in data[b:b + B])
    ex = sum(int(r[ix["Instructions Executed"]]) for r in data[b:b + B])
    print("%5d-%5d %7d %5.1f%%  %12d" % (b, b + B, sm, 100.0 * sm / tot, ex))
