#!/bin/bash
# r2_y: per-kernel times of the build (ncu launch list) with the regrouped tile climb, source-level capture of the two tree kernels, more regroup / tile variants
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2y_build.csv python tools/frame_once.py > gpurun_out/launches_r2y_build.log 2>&1
grep -E "k_refit_tris|k_tree_border|k_seg_setup_sort" gpurun_out/launches_r2y_build.csv | awk -F'","' '{print $5, $NF}' | head -12
for v in rg0 rg2; do
  [ -f build-up-phase_b200/build/librtcore_$v.so ] || continue
  RTCORE_LIB=$PWD/build-up-phase_b200/build/librtcore_$v.so timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r2y_build_$v.csv python tools/frame_once.py > gpurun_out/launches_r2y_build_$v.log 2>&1
  echo "== $v"; grep -E "k_refit_tris|k_tree_border|k_seg_setup_sort" gpurun_out/launches_r2y_build_$v.csv | awk -F'","' '{print $5, $NF}' | head -12
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_tree_border' -c 2 -o gpurun_out/prof_tree_r2y -f python tools/frame_once.py > gpurun_out/ncu_tree_r2y.log 2>&1
tail -2 gpurun_out/ncu_tree_r2y.log
BENCH_ARGS="--no-issue-counters --build-reps 5" bash tools/gpu_variants.sh
