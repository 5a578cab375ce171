#!/bin/bash
# N-GPU bench in the three gather modes. Usage: gpu_gather_modes.sh TAG N [steps]
mkdir -p gpurun_out
TAG=${1:-x}; N=${2:-2}; STEPS=${3:-20}
for mode in ${MODES:-p2p p2p:nccl gather}; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29540+N)) bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu-baseline --build-reps 1 --gather ${mode%%:*} --barrier $( [ "${mode##*:}" = "nccl" ] && echo nccl || echo flags ) \
      > gpurun_out/gather_${TAG}_${mode}_n$N.json 2> gpurun_out/gather_${TAG}_${mode}_n$N.err
  echo "$mode N=$N rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/gather_${TAG}_${mode}_n$N.json").read().strip().splitlines()[-1])
    print("$mode N=$N Mrays/s=%.1f ms=%.3f kernel_ms=%.3f e2e=%.1f crc=%s launches=%s" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d["e2e"]["value"], d["crc32"], d["gpu_launches"]))
except Exception as e:
    print("$mode FAILED", e); print(open("gpurun_out/gather_${TAG}_${mode}_n$N.err").read()[-1500:])
PY
done
