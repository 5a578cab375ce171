#!/bin/bash
# r2_w: new trace default (min/max slab test, 9 CTAs per SM, primary origin from the parameter block): GPU tests, variants, default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2w.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2w.log; tail -3 gpurun_out/pytest_gpu_r2w.log
bash tools/gpu_r2h.sh
timeout 900 python bench.py > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2w.json").read().strip().splitlines()[-1])
print("default", "Mrays/s=%.1f e2e=%.1f sync=%.1f kernel_ms=%.3f build=%.0f roofline=%s/%.3f issue=%s parity=%s" % (d["value"], d["e2e"]["value"], d["e2e"]["sync_value"], d["trace_kernel_ms"], d["build"]["value"], d["roofline"]["bound"], d["roofline"]["frac"], d["roofline"].get("sm_issue",{}).get("frac"), d["parity"]["primary"]))
PY
