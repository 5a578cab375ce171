#!/bin/bash
# r2_zj: source-level capture of the light per-BLAS kernel
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_seg2_setup_sort' -c 1 -o gpurun_out/prof_seg2_r2zj -f python tools/frame_once.py > gpurun_out/ncu_seg2_r2zj.log 2>&1
tail -1 gpurun_out/ncu_seg2_r2zj.log
