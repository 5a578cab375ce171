#!/bin/bash
# r2_j: border-kernel grid + small-segment rank sort: parity tests, smoke, default bench line (what the driver runs), TLAS update time
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2j.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r2j.log
tail -4 gpurun_out/pytest_gpu_r2j.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2j.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_r2j.log
timeout 300 python tools/gpu_tlas_time.py > gpurun_out/tlas_time_r2j.txt 2>&1; cat gpurun_out/tlas_time_r2j.txt
timeout 900 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2j.json").read().strip().splitlines()[-1])
print("default", "Mrays/s=%.1f e2e=%.1f kernel_ms=%.3f build=%.0f phases=%s tlas_ms=%.3f roofline=%s/%.3f parity=%s cpu=%s" % (d["value"], d["e2e"]["value"], d["trace_kernel_ms"], d["build"]["value"], d["build"]["phases_ms"], d["build"]["tlas_ms"], d["roofline"]["bound"], d["roofline"]["frac"], d["parity"]["primary"], d["cpu_baseline"]["value"]))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_r2j.json 2> gpurun_out/bench_ref_r2j.err; tail -c 600 gpurun_out/bench_ref_r2j.json
