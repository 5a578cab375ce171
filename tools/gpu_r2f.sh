#!/bin/bash
# r2_f: parity tests on the new default (region-major ray numbering + per-region fetch counters), then the variant table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2f.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r2f.log
tail -5 gpurun_out/pytest_gpu_r2f.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --build-reps 3 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f kernel_ms=%.3f crc=%s" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d.get("crc32")))
PY
bash tools/gpu_variants.sh
