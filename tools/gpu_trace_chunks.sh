#!/bin/bash
# Frame traced in K row chunks alternating over two compute streams (tail filling). Device-timed value and e2e per K.
mkdir -p gpurun_out
for k in 1 2 3 4 6 8; do
  RTCORE_TRACE_CHUNKS=$k timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --build-reps 1 "$@" > gpurun_out/trace_chunks_$k.json 2> gpurun_out/trace_chunks_$k.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/trace_chunks_$k.json").read().strip().splitlines()[-1])
    print("chunks=$k value=%.1f ms=%.3f kernel_ms=%.3f e2e=%.1f e2e_ms=%.3f crc=%s launches=%d" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["crc32"]["rgba"], d["gpu_launches"]))
except Exception as e:
    print("chunks=$k FAILED", e, open("gpurun_out/trace_chunks_$k.err").read()[-600:])
PY
done
