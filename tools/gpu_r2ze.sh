#!/bin/bash
# r2_ze: per-warp choice between MATCH.ANY and 8 ballots in the onesweep ranking (threshold = distinct digits of the warp's first item)
mkdir -p gpurun_out
for lib in "" bd0 bd8 bd24 bd33; do
  L=${lib:+$PWD/build-up-phase_b200/build/librtcore_$lib.so}
  for args in "--workload soup10m" "--workload tess1m" "--workload inst10m --flags 0x400"; do
    RTCORE_LIB=$L timeout 300 python tools/gpu_build_paths.py $args 2>gpurun_out/ze.err | sed "s/^/${lib:-default16} /" | tee -a gpurun_out/sort_rank_r2ze.jsonl | cut -c1-260
  done
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lbvh or build or golden or tess or batch" 2>&1 | tail -2
