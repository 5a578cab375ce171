#!/bin/bash
# GPU call: gpu tests + default bench + tuning variants
mkdir -p gpurun_out
TAG=${1:-r01d}
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1; nproc >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_$TAG.err; cut -c1-600 gpurun_out/bench_$TAG.json
bash tools/gpu_variants.sh
