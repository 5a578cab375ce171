#!/bin/bash
# Extra workloads (cfg3 tess1m, cfg5 soup) on an N-GPU box. Usage: gpu_round_workloads.sh TAG N
mkdir -p gpurun_out
TAG=${1:-x}; N=${2:-2}
run1() { # name args...
  name=$1; shift
  timeout 1500 python bench.py "$@" > gpurun_out/bench_${TAG}_$name.json 2> gpurun_out/bench_${TAG}_$name.err; echo "$name rc=$?"
  tail -2 gpurun_out/bench_${TAG}_$name.err; cut -c1-250 gpurun_out/bench_${TAG}_$name.json; echo
}
runN() { # name n args...
  name=$1; n=$2; shift; shift
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $n "$@" \
     > gpurun_out/bench_${TAG}_${name}_n$n.json 2> gpurun_out/bench_${TAG}_${name}_n$n.err; echo "$name N=$n rc=$?"
  tail -2 gpurun_out/bench_${TAG}_${name}_n$n.err; cut -c1-250 gpurun_out/bench_${TAG}_${name}_n$n.json; echo
}
run1 tess1m --workload tess1m --steps 20 --warmup 3
runN soup10m $N --workload soup10m --steps 10 --warmup 3
run1 soup10m_n1 --workload soup10m --steps 10 --warmup 3 --no-cpu-baseline
runN soup100m $N --workload soup100m --steps 5 --warmup 3 --build-reps 1
