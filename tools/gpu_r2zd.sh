#!/bin/bash
# r2_zd: the global onesweep sort on random keys (soup10m) and on a surface (tess1m): MATCH.ANY vs 8 ballots, CTAs per SM, look-back window
mkdir -p gpurun_out
for w in soup10m tess1m; do
for lib in "" build-up-phase_b200/build/librtcore_ballot.so build-up-phase_b200/build/librtcore_sort12_6.so build-up-phase_b200/build/librtcore_lb16.so; do
  name=$(basename "${lib:-base}" .so)
  RTCORE_LIB=${lib:+$PWD/$lib} timeout 600 python bench.py --workload $w --width 1920 --height 1080 --steps 3 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 7 > gpurun_out/zd_${w}_$name.json 2> gpurun_out/zd_${w}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/zd_${w}_$name.json").read().strip().splitlines()[-1])
    print("$w $name", "build_Mtri/s=%.0f phases=%s" % (d["build"]["value"], {k: round(v,4) for k,v in d["build"]["phases_ms"].items()}))
except Exception as e: print("$w $name FAILED", e)
PY
done; done
