#!/bin/bash
# r2_x: regrouped tile climb + border jobs that carry their segment and direction (new default), tile / regroup variants, trace cache hints
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2x.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2x.log; tail -3 gpurun_out/pytest_gpu_r2x.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 5 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f build_Mtri/s=%.0f refit_ms=%.3f sort_ms=%.3f crc=%s" % (d["value"], d["ms_per_step"], d["build"]["value"], d["build"]["phases_ms"]["refit_ms"], d["build"]["phases_ms"]["sort_ms"], d.get("crc32")))
PY
BENCH_ARGS="--no-issue-counters --build-reps 5" bash tools/gpu_variants.sh
for w in tess1m soup10m; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 5 > gpurun_out/bench_r2x_$w.json 2> gpurun_out/bench_r2x_$w.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2x_$w.json").read().strip().splitlines()[-1])
print("$w", "Mrays/s=%.1f build_Mtri/s=%.0f phases=%s parity=%s" % (d["value"], d["build"]["value"], d["build"]["phases_ms"], d.get("parity",{}).get("primary") if d.get("parity") else None))
PY
done
