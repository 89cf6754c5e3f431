"""Side-by-side per-launch conv timings of two B2N_PROF_DUMP files (tools/ab_run.sh)."""
import json
import sys

a, b = (json.load(open(f)) for f in sys.argv[1:3])
for name in a:
    ea, eb = a[name], b.get(name, [])
    n = len(ea) // 2
    if len(eb) != len(ea):
        print(name, "launch counts differ", len(ea), len(eb))
        continue
    ta = tb = 0.0
    for i in range(n):
        x = (ea[i][0] + ea[n + i][0]) / 2
        y = (eb[i][0] + eb[n + i][0]) / 2
        ta += x
        tb += y
        if abs(x - y) > 0.04 * x:
            print("  %-18s %3d %8.1f -> %8.1f us  (%+.0f %%)  %7.2f GFLOP" % (name, i, x, y, 100 * (y - x) / x, ea[i][1] / 1e9))
    print("%s total %.2f -> %.2f ms" % (name, ta / 1e3, tb / 1e3))
