#!/bin/bash
# r2_h: 256-bit node loads re-measured on the current kernel
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --build-reps 3 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f kernel_ms=%.3f crc=%s" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d.get("crc32")))
PY
bash tools/gpu_variants.sh
