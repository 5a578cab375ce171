#!/bin/bash
# r2_m: pipelined host output (two frames in flight): GPU tests, default bench at N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2m.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r2m.log
tail -6 gpurun_out/pytest_gpu_r2m.log
timeout 900 python bench.py > gpurun_out/bench_r2m.json 2> gpurun_out/bench_r2m.err; tail -3 gpurun_out/bench_r2m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2m.json").read().strip().splitlines()[-1])
print("default", "Mrays/s=%.1f e2e=%s kernel_ms=%.3f build=%.0f roofline=%s/%.3f parity=%s crc=%s" % (d["value"], d["e2e"], d["trace_kernel_ms"], d["build"]["value"], d["roofline"]["bound"], d["roofline"]["frac"], d["parity"]["primary"], d["crc32"]))
PY
