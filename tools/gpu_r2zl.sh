#!/bin/bash
# r2_zl: sorted-record emission in the LIGHT per-BLAS kernel (two CTAs per SM) vs the tile kernel's own gather
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 5 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f build_Mtri/s=%.0f refit_ms=%.3f sort_ms=%.3f crc=%s" % (d["value"], d["build"]["value"], d["build"]["phases_ms"]["refit_ms"], d["build"]["phases_ms"]["sort_ms"], d.get("crc32",{}).get("rgba")))
PY
BENCH_ARGS="--no-issue-counters --build-reps 5" bash tools/gpu_variants.sh 2>&1 | sed 's/, .primary_hits.*//'
RTCORE_LIB=$PWD/build-up-phase_b200/build/librtcore_emit3.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
