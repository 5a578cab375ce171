#!/bin/bash
# r2_head: the GPU tiers the driver runs at round end, on the final commit: pytest -m gpu, smoke(), default bench.py
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2_head.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2_head.log; tail -3 gpurun_out/pytest_gpu_r2_head.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_r2_head.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_r2_head.log
timeout 900 python bench.py > gpurun_out/bench_r2_head.json 2> gpurun_out/bench_r2_head.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2_head.json").read().strip().splitlines()[-1])
print("head", "Mrays/s=%.1f e2e=%.1f sync=%.1f build=%.1f Mtri/s (%.4f ms) tlas_ms=%.4f roofline=%s/%.3f issue=%.3f parity=%s crc=%s" % (d["value"], d["e2e"]["value"], d["e2e"].get("sync_value",0), d["build"]["value"], d["build"]["ms"], d["build"]["tlas_ms"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"]["sm_issue"]["frac"], d["parity"]["primary"], d["crc32"]["rgba"]))
PY
