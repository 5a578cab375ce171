#!/bin/bash
# Compare tuning variants of librtcore (built by tools/build_variants.py) on the bench workload.
mkdir -p gpurun_out
for lib in build-up-phase_b200/build/librtcore_*.so; do
  name=$(basename $lib .so)
  RTCORE_LIB=$PWD/$lib timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --build-reps 3 ${BENCH_ARGS} > gpurun_out/var_$name.json 2> gpurun_out/var_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/var_$name.json").read().strip().splitlines()[-1])
    print("$name", "Mrays/s=%.1f ms=%.3f build_Mtri/s=%.0f setup_ms=%.3f refit_ms=%.3f sort_ms=%.3f crc=%s" % (d["value"], d["ms_per_step"], d["build"]["value"], d["build"]["phases_ms"]["setup_ms"], d["build"]["phases_ms"]["refit_ms"], d["build"]["phases_ms"]["sort_ms"], d.get("crc32")))
except Exception as e:
    print("$name", "FAILED", e, open("gpurun_out/var_$name.err").read()[-500:])
PY
done
