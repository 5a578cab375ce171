#!/bin/bash
# GPU call: gpu tests + default bench + full ncu capture of the build kernels of the second build. Usage: gpu_call_e.sh TAG
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step")}, d["e2e"]["value"], d["build"]["value"], d["build"]["phases_ms"], d.get("parity"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_tree_border|k_onesweep_pass|k_tri_setup|k_tri_morton|k_sort_hist' -s 9 -c 9 \
    -o gpurun_out/prof_build_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 2 > gpurun_out/ncu_build_$TAG.log 2>&1
tail -3 gpurun_out/ncu_build_$TAG.log
