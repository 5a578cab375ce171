#!/bin/bash
# r2_l: N GPUs of one box: inst10m (default bench line) and soup100m (cfg5) with parity + CPU arm; usage: gpu_r2l.sh N
N=${1:-8}; TAG=${2:-r2l}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | sort | uniq -c > gpurun_out/gpus_${TAG}_n$N.txt
if [ "$N" = 1 ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541"; fi
timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_n$N.err
timeout 1500 $TR bench.py --gpus $N --workload soup100m --steps 3 > gpurun_out/bench_${TAG}_soup100m_n$N.json 2> gpurun_out/bench_${TAG}_soup100m_n$N.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_soup100m_n$N.err
for f in gpurun_out/bench_${TAG}*n$N.err; do echo "== $f"; tail -n 3 $f; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], 'Mrays/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel ms', d['trace_kernel_ms_per_rank']['min'], d['trace_kernel_ms_per_rank']['max'], 'step', d['ms_per_step'], 'build', round(d['build']['value']), d['build'].get('variants'), d.get('parity',{}).get('primary'), d['crc32'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
