#!/bin/bash
# GPU call: tests, default bench + reference arm, launch list, one full ncu capture (with source) of both trace stages,
# then the tuning variants. Usage: gpu_call_a.sh TAG
mkdir -p gpurun_out
TAG=${1:-r01g}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err; cut -c1-400 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2>> gpurun_out/bench_$TAG.err
cut -c1-300 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -o gpurun_out/prof_trace_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -2 gpurun_out/ncu_full_$TAG.log
bash tools/gpu_variants.sh
ls -la gpurun_out | head -40
