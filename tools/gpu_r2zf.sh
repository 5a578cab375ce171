#!/bin/bash
# r2_zf: full GPU suite at HEAD + soup10m / tess1m lines with the batched cfg5 build variant
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r2zf.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2zf.log; tail -3 gpurun_out/pytest_gpu_r2zf.log
for w in soup10m tess1m; do
  timeout 900 python bench.py --workload $w --steps 5 --warmup 3 --build-reps 5 > gpurun_out/bench_r2zf_$w.json 2> gpurun_out/bench_r2zf_$w.err; echo "$w rc=$?"
  python - <<PY
import json
d=json.loads(open("gpurun_out/bench_r2zf_$w.json").read().strip().splitlines()[-1])
print("$w", "Mrays/s=%.1f e2e=%.1f build_Mtri/s=%.0f variants=%s phases=%s parity=%s cpu=%s" % (d["value"], d["e2e"]["value"], d["build"]["value"], {k: round(v["ms"],3) for k,v in (d["build"].get("variants") or {}).items()}, {k: round(v,4) for k,v in d["build"]["phases_ms"].items()}, d.get("parity",{}).get("primary"), d.get("cpu_baseline",{}).get("value")))
PY
done
