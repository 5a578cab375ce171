#!/bin/bash
# GPU round script: tests, bench (N=1), ncu launch list of the same bench command, one full ncu capture
# of the dominant kernel. Run under gpurun from the repo root; outputs land in gpurun_out/.
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err; cat gpurun_out/bench_$TAG.json
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
# launch list of the same command (per-launch times are cold-cache + serialised: compare SHARES)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
# one full capture of the dominant kernel (skip the stats-variant launch and warm-ups)
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 2 -c 1 -o gpurun_out/prof_trace_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out
