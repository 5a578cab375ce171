#!/bin/bash
# cfg5 (soup100m: 100 M triangles, 7680x4320, primary + bounce) at the listed GPU counts; usage: gpu_soup100m.sh TAG N [N ...]
TAG=$1; shift
mkdir -p gpurun_out
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 1500 python bench.py --workload soup100m --steps 3 --warmup 3 --build-reps 1 > gpurun_out/bench_${TAG}_soup100m_n1.json 2> gpurun_out/bench_${TAG}_soup100m_n1.err
  else
    timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $N --workload soup100m --steps 3 --warmup 3 --build-reps 1 \
        > gpurun_out/bench_${TAG}_soup100m_n$N.json 2> gpurun_out/bench_${TAG}_soup100m_n$N.err
  fi
  echo "N=$N rc=$?"; tail -n 3 gpurun_out/bench_${TAG}_soup100m_n$N.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_*soup100m*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'N', d['n_gpus'], round(d['value']), 'e2e', round(d['e2e']['value']), 'step ms', round(d['ms_per_step'],2), 'build Mtri/s', round(d['build']['value']), json.dumps(d['build'].get('variants')), d.get('parity',{}).get('primary'), d.get('parity',{}).get('secondary'), d['crc32'], d.get('cpu_baseline'))
    except Exception as e: print(f, 'ERR', e)
PY
