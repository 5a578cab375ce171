#!/bin/bash
# GPU call: gpu tests (default lib, and again with the split traverse|shade path), then the tuning variants.
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
RTCORE_SPLIT_SHADE=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_split_$TAG.log 2>&1; echo "pytest(split) rc=$?" >> gpurun_out/pytest_gpu_split_$TAG.log
tail -4 gpurun_out/pytest_gpu_split_$TAG.log
bash tools/gpu_variants.sh
