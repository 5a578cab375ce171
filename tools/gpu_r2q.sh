#!/bin/bash
# r2_q: border climbs inlined into the tile kernel (RT_TREE_INLINE_BORDER=1) vs the two-kernel tree pass: parity of the build, then build time
mkdir -p gpurun_out
RTCORE_LIB=$PWD/build-up-phase_b200/build/librtcore_inline_border.so timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest_gpu_r2q_inline.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/pytest_gpu_r2q_inline.log
for w in inst10m tess1m soup10m; do
for lib in base inline_border inline_border_t256; do
  L=""; [ $lib != base ] && L=$PWD/build-up-phase_b200/build/librtcore_$lib.so
  RTCORE_LIB=$L timeout 300 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 6 > gpurun_out/ib_${w}_$lib.json 2> gpurun_out/ib_${w}_$lib.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ib_${w}_$lib.json").read().strip().splitlines()[-1])
    print("$w $lib build=%.0f Mtri/s total=%.4f refit_ms=%.4f tlas_ms=%.4f crc=%s" % (d["build"]["value"], d["build"]["ms"], d["build"]["phases_ms"]["refit_ms"], d["build"]["tlas_ms"], d["crc32"]["rgba"]))
except Exception as e: print("$w $lib FAILED", e, open("gpurun_out/ib_${w}_$lib.err").read()[-400:])
PY
done; done
