#!/bin/bash
# compute-sanitizer (memcheck, then racecheck + synccheck on the shared-memory kernels) over the tests that exercise every kernel on
# small scenes; the reports land in gpurun_out/sanitize_*.log
mkdir -p gpurun_out
SEL='sample_scene or single_triangle or edge_cases or anyhit or refit or compaction or tlas_update or sbt_stride or instanced_batched or lbvh_build_matches'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitize_memcheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "$SEL" > gpurun_out/sanitize_memcheck_pytest.log 2>&1; echo "memcheck rc=$?"
tail -3 gpurun_out/sanitize_memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitize_memcheck.log | head -5
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitize_racecheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "sample_scene or compaction or tlas_update or instanced_batched" > gpurun_out/sanitize_racecheck_pytest.log 2>&1; echo "racecheck rc=$?"
tail -3 gpurun_out/sanitize_racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | head -5
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 --log-file gpurun_out/sanitize_synccheck.log python -m pytest tests/test_gpu_parity.py -x -q -k "sample_scene or compaction or instanced_batched" > gpurun_out/sanitize_synccheck_pytest.log 2>&1; echo "synccheck rc=$?"
tail -3 gpurun_out/sanitize_synccheck_pytest.log; grep -E "ERROR SUMMARY" gpurun_out/sanitize_synccheck.log | head -3
