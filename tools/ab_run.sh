#!/bin/bash
# A/B of kernel-selection knobs on one box: per-launch conv timings (B2N_PROF_DUMP) and step times.
# usage: tools/ab_run.sh <tag>=<ENV=VAL,...> ...   (results under gpurun_out/ab_<tag>.{json,log})
mkdir -p gpurun_out
for spec in "$@"; do
  tag="${spec%%=*}"; envs="${spec#*=}"
  envs="${envs//,/ }"
  [ "$envs" = "$spec" ] && envs=""
  env $envs B2N_PROF_DUMP=gpurun_out/ab_${tag}.json python bench.py --steps 8 --warmup 3 \
      --no-cpu-baseline --no-library-baseline --no-secondary --no-peaks > gpurun_out/ab_${tag}.log 2>&1
  python - "$tag" <<'PY'
import json, sys
tag = sys.argv[1]
try:
    line = [l for l in open("gpurun_out/ab_%s.log" % tag) if l.startswith("{")][-1]
    d = json.loads(line)
    print(tag, "ms_per_step %.3f" % d["ms_per_step"], "e2e %.3f" % d["e2e"]["ms_per_step"], d["clocks"])
except Exception as e:
    print(tag, "FAILED", e)
    print(open("gpurun_out/ab_%s.log" % tag).read()[-1500:])
PY
done
