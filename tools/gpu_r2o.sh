#!/bin/bash
# r2_o: final-code evidence on one GPU: GPU tests, default bench line, ncu launch list of the bench command, full captures of both trace
# stages and of one whole build, SAH table, tess1m line
mkdir -p gpurun_out
TAG=r2o
bash tools/gpu_call_d.sh $TAG launches
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_trace -c 2 -o gpurun_out/prof_trace_$TAG -f python tools/frame_once.py > gpurun_out/ncu_trace_$TAG.log 2>&1
tail -2 gpurun_out/ncu_trace_$TAG.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_refit_tris|k_tree_border|k_seg_setup_sort' -s 3 -c 3 -o gpurun_out/prof_build_$TAG -f python tools/frame_once.py > gpurun_out/ncu_build_$TAG.log 2>&1
tail -2 gpurun_out/ncu_build_$TAG.log
timeout 600 python tools/gpu_sah.py inst10m tess1m soup10m > gpurun_out/sah_$TAG.jsonl 2> gpurun_out/sah_$TAG.err; cat gpurun_out/sah_$TAG.jsonl; tail -2 gpurun_out/sah_$TAG.err
timeout 600 python bench.py --workload tess1m > gpurun_out/bench_${TAG}_tess1m.json 2> gpurun_out/bench_${TAG}_tess1m.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2o_tess1m.json").read().strip().splitlines()[-1])
print("tess1m", d["value"], d["e2e"]["value"], d["build"]["value"], d["roofline"]["bound"], d["roofline"]["frac"], d["parity"]["primary"])
PY
