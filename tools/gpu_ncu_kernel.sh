#!/bin/bash
# One full ncu capture (with source) of kernels matching a regex during a short bench run.
# Usage: gpu_ncu_kernel.sh TAG REGEX [skip] [count] [extra bench args...]
mkdir -p gpurun_out
TAG=$1; RE=$2; SKIP=${3:-1}; CNT=${4:-1}; shift 4
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c $CNT -o gpurun_out/prof_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 2 "$@" > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
