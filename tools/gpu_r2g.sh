#!/bin/bash
# r2_g: shared-memory TLAS (BIG kernel variant): parity tests, variant table, full ncu capture of both trace stages
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2g.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r2g.log
tail -5 gpurun_out/pytest_gpu_r2g.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --build-reps 3 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f kernel_ms=%.3f crc=%s" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d.get("crc32")))
PY
bash tools/gpu_variants.sh
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -o gpurun_out/prof_trace_r2g -f python tools/frame_once.py > gpurun_out/ncu_trace_r2g.log 2>&1
tail -3 gpurun_out/ncu_trace_r2g.log
