#!/bin/bash
# r2_final2 at N GPUs: inst10m default + soup10m (split / replicated / replicated_batched builds) + the group tests; usage: gpu_r2_final2_n2.sh N
N=${1:-2}; TAG=r2_final2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_n$N.err
timeout 900 $TR bench.py --gpus $N --workload soup10m --steps 5 > gpurun_out/bench_${TAG}_soup10m_n$N.json 2> gpurun_out/bench_${TAG}_soup10m_n$N.err; echo "rc=$?" >> gpurun_out/bench_${TAG}_soup10m_n$N.err
for f in gpurun_out/bench_${TAG}*n$N.err; do echo "== $f"; tail -n 3 $f; done
python - <<PY
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_${TAG}*_n$N.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']), 'kernel ms', d['trace_kernel_ms_per_rank']['min'], d['trace_kernel_ms_per_rank']['max'], 'step', d['ms_per_step'], 'build', round(d['build']['value']), {k: round(v['ms'],3) for k,v in (d['build'].get('variants') or {}).items()}, d.get('parity',{}).get('primary'), d['crc32'].get('rgba'), d['crc32'].get('assembled_equals_single_gpu_frame'))
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 python -m pytest tests/test_group_gpu.py -m gpu -x -q 2>&1 | tail -2
