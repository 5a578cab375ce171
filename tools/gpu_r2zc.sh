#!/bin/bash
# r2_zc: prefetch.global.L1 of a leaf's first record when a lane reaches the leaf (trace experiment)
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 3 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f kernel_ms=%.3f build_Mtri/s=%.0f crc=%s" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d["build"]["value"], d.get("crc32",{}).get("rgba")))
PY
BENCH_ARGS="--no-issue-counters --build-reps 3" bash tools/gpu_variants.sh 2>&1 | sed 's/build_Mtri.*crc=/crc=/; s/, .primary_hits.*//'
