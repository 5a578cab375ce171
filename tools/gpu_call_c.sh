#!/bin/bash
# GPU call: gpu tests, default bench, then one full ncu capture of the BUILD kernels (second build of the run).
mkdir -p gpurun_out
TAG=${1:-x}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_$TAG.log
tail -4 gpurun_out/pytest_gpu_$TAG.log
timeout 900 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; echo "bench rc=$?"
head -c 600 gpurun_out/bench_$TAG.json; echo
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_karras|k_onesweep_pass|k_tri_setup|k_tri_morton|k_sort_hist|k_seg' -s 10 -c 10 \
    -o gpurun_out/prof_build_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --build-reps 2 > gpurun_out/ncu_build_$TAG.log 2>&1
tail -3 gpurun_out/ncu_build_$TAG.log
ls -la gpurun_out/*.ncu-rep
