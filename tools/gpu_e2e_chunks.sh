#!/bin/bash
# e2e (host-buffer rt_trace) as a function of the number of row chunks whose D2H copy overlaps the next chunk's trace
mkdir -p gpurun_out
for k in ${CHUNK_LIST:-1 2 3 4 6 8}; do
  RTCORE_E2E_CHUNKS=$k timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/e2e_chunks_$k.json 2> gpurun_out/e2e_chunks_$k.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_chunks_$k.json").read().strip().splitlines()[-1])
print("chunks=$k value=%.1f e2e=%.1f e2e_ms=%.3f crc=%s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["crc32"]["rgba"]))
PY
done
