"""Pretty-print the per-launch conv timings bench.py writes with B2N_PROF_DUMP=<file>."""
import json
import sys

d = json.load(open(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/prof_dump.json"))
for name, ev in d.items():
    n = len(ev) // 2
    print(name, n, "launches/step")
    tot = 0
    for i, (us, w) in enumerate(ev[:n]):
        t = (us + ev[n + i][0]) / 2
        tot += t
        print("  %3d %8.1f us  %8.2f GFLOP  %7.1f TFLOP/s" % (i, t, w / 1e9, w / t / 1e6))
    print("  total ms %.2f" % (tot / 1e3))
