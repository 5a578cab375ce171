#!/bin/bash
# r2_i: refit-only update + compaction: GPU parity tests, refit-vs-rebuild quality table, inst10m traced compacted vs not
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2i.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_gpu_r2i.log
tail -15 gpurun_out/pytest_gpu_r2i.log
timeout 600 python tools/gpu_refit_quality.py > gpurun_out/refit_quality_r2i.jsonl 2> gpurun_out/refit_quality_r2i.err; echo "rc=$?"; cat gpurun_out/refit_quality_r2i.jsonl; tail -3 gpurun_out/refit_quality_r2i.err
for tag in plain compact; do
  extra=""; [ $tag = compact ] && extra="--compact"
  timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --build-reps 3 $extra > gpurun_out/bench_r2i_$tag.json 2> gpurun_out/bench_r2i_$tag.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2i_$tag.json").read().strip().splitlines()[-1])
print("$tag", "Mrays/s=%.1f kernel_ms=%.3f build=%.0f roofline=%s/%.3f compaction=%s crc=%s" % (d["value"], d["trace_kernel_ms"], d["build"]["value"], d["roofline"]["bound"], d["roofline"]["frac"], d.get("compaction"), d["crc32"]["rgba"]))
PY
done
