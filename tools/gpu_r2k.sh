#!/bin/bash
# r2_k: grid of the border kernel (threads per tile) vs build time, inst10m and tess1m
mkdir -p gpurun_out
for w in inst10m tess1m; do
for k in 4 8 16 32 64 130; do
  RTCORE_BORDER_JOBS_PER_TILE=$k timeout 300 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 6 > gpurun_out/border_$w_$k.json 2> gpurun_out/border_$w_$k.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/border_$w_$k.json").read().strip().splitlines()[-1])
print("$w per_tile=$k build=%.0f Mtri/s total=%.4f refit_ms=%.4f crc=%s" % (d["build"]["value"], d["build"]["ms"], d["build"]["phases_ms"]["refit_ms"], d["crc32"]["rgba"]))
PY
done; done
