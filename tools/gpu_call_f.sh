#!/bin/bash
# GPU call: full ncu captures (with source) of the trace stages and of one whole build. Usage: gpu_call_f.sh TAG
mkdir -p gpurun_out
TAG=${1:-x}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -s 4 -c 2 -o gpurun_out/prof_trace_$TAG -f \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/ncu_trace_$TAG.log 2>&1
tail -2 gpurun_out/ncu_trace_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_tree_border|k_seg_setup_sort|k_onesweep_pass|k_tri_setup|k_tri_morton|k_sort_hist' -s 3 -c 3 \
    -o gpurun_out/prof_build_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 2 > gpurun_out/ncu_build_$TAG.log 2>&1
tail -2 gpurun_out/ncu_build_$TAG.log
ls -la gpurun_out/*.ncu-rep
