#!/bin/bash
# Fixed per-frame cost of the trace: the same scene at 1/1, 1/4, 1/8, 1/16 of the 4K pixel count on one GPU
mkdir -p gpurun_out
for wh in "3840 2160" "1920 1080" "1360 768" "960 540"; do
  set -- $wh
  timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --build-reps 1 --width $1 --height $2 > gpurun_out/res_$1.json 2> gpurun_out/res_$1.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/res_$1.json").read().strip().splitlines()[-1])
print("$1x$2 rays=%d Mrays/s=%.1f ms=%.3f kernel_ms=%.3f ns/ray=%.3f" % (d["rays_per_step"], d["value"], d["ms_per_step"], d["trace_kernel_ms"], d["ms_per_step"]*1e6/d["rays_per_step"]))
PY
done
