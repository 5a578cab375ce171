#!/bin/bash
# r2_final2: verification of the final code (light per-BLAS kernel, per-size stand-alone segment sort) on one B200: GPU tests, smoke, default bench + reference arm, launch list, full captures of the
mkdir -p gpurun_out
TAG=r2_final2
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; echo "ref rc=$?"
timeout 300 python bench.py --workload sample > gpurun_out/bench_${TAG}_cfg2.json 2> gpurun_out/bench_${TAG}_cfg2.err; echo "cfg2 rc=$?"
timeout 300 python bench.py --workload sample --width 1200 --height 800 > gpurun_out/bench_${TAG}_cfg1.json 2> gpurun_out/bench_${TAG}_cfg1.err; echo "cfg1 rc=$?"
timeout 600 python bench.py --workload tess1m > gpurun_out/bench_${TAG}_tess1m.json 2> gpurun_out/bench_${TAG}_tess1m.err; echo "tess1m rc=$?"
python - <<PY
import json
for f in ("bench_$TAG","bench_${TAG}_cfg2","bench_${TAG}_cfg1","bench_${TAG}_tess1m"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().splitlines()[-1])
        print(f, "Mrays/s=%.1f e2e=%.1f sync=%.1f ms/step=%.4f build=%.1f Mtri/s (%.4f ms) tlas_ms=%.4f roofline=%s/%.3f parity=%s launches=%s" % (d["value"], d["e2e"]["value"], d["e2e"].get("sync_value",0), d["ms_per_step"], d["build"]["value"], d["build"]["ms"], d["build"]["tlas_ms"], d["roofline"]["bound"], d["roofline"]["frac"], d.get("parity",{}).get("primary"), d["gpu_launches"]))
    except Exception as e: print(f, "ERR", e)
print(open("gpurun_out/bench_ref_$TAG.json").read()[-300:])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-issue-counters > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_tree_border|k_seg2_setup_sort' -c 3 -o gpurun_out/prof_build_$TAG -f python tools/frame_once.py > gpurun_out/ncu_build_$TAG.log 2>&1
tail -1 gpurun_out/ncu_build_$TAG.log
timeout 300 python tools/gpu_tlas_time.py 2>&1 | tail -3
