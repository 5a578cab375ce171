#!/bin/bash
# Multi-GPU call: bench at N GPUs under torchrun (driver's launch line). Usage: gpu_round_multi.sh TAG N [extra bench args]
mkdir -p gpurun_out
TAG=${1:-x}; N=${2:-2}; shift; shift
nvidia-smi --query-gpu=index,name --format=csv,noheader > gpurun_out/gpus_$TAG.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
   > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "bench N=$N rc=$?"
tail -5 gpurun_out/bench_${TAG}_n$N.err; cut -c1-900 gpurun_out/bench_${TAG}_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 1 --warmup 0 --cpu-seconds 5 \
   > gpurun_out/bench_ref_${TAG}_n$N.json 2>> gpurun_out/bench_${TAG}_n$N.err; echo "ref N=$N rc=$?"; cut -c1-300 gpurun_out/bench_ref_${TAG}_n$N.json
