#!/bin/bash
# GPU call: gpu tests + tuning variants (no full bench)
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
bash tools/gpu_variants.sh
