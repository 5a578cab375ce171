#!/bin/bash
# Scaling run on an N-GPU box: bench at 1, 2, 4, ..., N GPUs (driver's launch line). Usage: gpu_scale.sh TAG N
mkdir -p gpurun_out
TAG=${1:-x}; NMAX=${2:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$TAG.txt
n=1
while [ $n -le $NMAX ]; do
  if [ $n -eq 1 ]; then
    timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --build-reps 1 > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+n)) bench.py --gpus $n --steps 20 --warmup 3 --no-cpu-baseline --build-reps 1 \
      > gpurun_out/scale_${TAG}_n$n.json 2> gpurun_out/scale_${TAG}_n$n.err
  fi
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/scale_${TAG}_n$n.json").read().strip().splitlines()[-1])
    print("N=$n Mrays/s=%.1f ms=%.3f kernel_ms=%.3f e2e=%.1f build=%.0f" % (d["value"], d["ms_per_step"], d["trace_kernel_ms"], d["e2e"]["value"], d["build"]["value"]), {k:v for k,v in d.items() if k in ("gather_ms","unpack_ms","phases")})
except Exception as e:
    print("N=$n FAILED", e); print(open("gpurun_out/scale_${TAG}_n$n.err").read()[-800:])
PY
  n=$((n*2))
done
