#!/bin/bash
# round-2 first bench contact: N = 1 default (pipelined frames, live ncu counters), without pipelining, cfg3, a small soup
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "rc=$?" >> gpurun_out/bench_r2b.err
timeout 300 python bench.py --no-pipeline --no-cpu-baseline --no-issue-counters > gpurun_out/bench_r2b_nopipe.json 2> gpurun_out/bench_r2b_nopipe.err
timeout 600 python bench.py --workload tess1m > gpurun_out/bench_r2b_tess1m.json 2> gpurun_out/bench_r2b_tess1m.err
timeout 600 python bench.py --workload soup10m --steps 5 > gpurun_out/bench_r2b_soup10m.json 2> gpurun_out/bench_r2b_soup10m.err
tail -3 gpurun_out/*.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_r2b*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, round(d['value']), 'e2e', round(d['e2e']['value']), 'build', round(d['build']['value']), 'roof', d['roofline'].get('bound'), round(d['roofline']['frac'],3), d.get('parity',{}).get('primary'), d['crc32'])
    except Exception as e: print(f, 'ERR', e)
PY
