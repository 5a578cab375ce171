#!/bin/bash
# r2_zk: light segment kernel: indices fetched two iterations ahead, Morton re-reads in batches of four
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2zk.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2zk.log; tail -3 gpurun_out/pytest_gpu_r2zk.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2zk_build.csv python tools/frame_once.py > gpurun_out/launches_r2zk_build.log 2>&1
grep -E "k_refit_tris|k_tree_border<2|k_seg2_setup_sort|k_seg_setup_sort" gpurun_out/launches_r2zk_build.csv | awk -F'","' '{print substr($5,1,50), $(NF-2), $NF}' | head -12
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-issue-counters --build-reps 5 > gpurun_out/var_base.json 2> gpurun_out/var_base.err
python - <<PY
import json
d=json.loads(open("gpurun_out/var_base.json").read().strip().splitlines()[-1])
print("base", "Mrays/s=%.1f ms=%.3f build_Mtri/s=%.0f refit_ms=%.3f sort_ms=%.3f tlas_ms=%.4f crc=%s" % (d["value"], d["ms_per_step"], d["build"]["value"], d["build"]["phases_ms"]["refit_ms"], d["build"]["phases_ms"]["sort_ms"], d["build"]["tlas_ms"], d.get("crc32",{}).get("rgba")))
PY
BENCH_ARGS="--no-issue-counters --build-reps 5" bash tools/gpu_variants.sh 2>&1 | sed 's/crc=.*//'
