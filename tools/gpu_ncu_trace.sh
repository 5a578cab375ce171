#!/bin/bash
# One full ncu capture (with source) of both trace stages. Usage: gpu_ncu_trace.sh TAG [lib]
mkdir -p gpurun_out
TAG=${1:-x}
[ -n "$2" ] && export RTCORE_LIB=$PWD/$2
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -o gpurun_out/prof_trace_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
tail -3 gpurun_out/ncu_full_$TAG.log
ls -la gpurun_out/*.ncu-rep
