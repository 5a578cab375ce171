#!/usr/bin/env python
"""Build phases (CUDA events inside the library) of a bench workload through a chosen build path: --flags 0x400 forces the global onesweep
sort for batches of small BLASes (RT_BUILD_NO_SEGMENTED_SORT), 0x800 the unfused setup. Prints the median of --reps builds as one JSON line."""
import argparse, json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from build_up_phase_b200 import rtcore  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="inst10m")
ap.add_argument("--flags", type=lambda s: int(s, 0), default=0)
ap.add_argument("--reps", type=int, default=7)
a = ap.parse_args()
scene = bench.make_workload(a.workload, 64, 64)
with rtcore.Context(0) as ctx:
    rows = []
    for _ in range(a.reps):
        blases = ctx.build_blas_batch(scene.blases, flags=a.flags) if len(scene.blases) > 1 else [ctx.build_blas(scene.blases[0], flags=a.flags)]
        rows.append(ctx.build_timing())
        for b in blases:
            b.free()
    keys = [k for k, v in rows[0].items() if isinstance(v, (int, float))]
    med = {k: float(np.median([r[k] for r in rows])) for k in keys}
    print(json.dumps({"workload": a.workload, "flags": a.flags, "lib": os.environ.get("RTCORE_LIB", "default"), **{k: round(v, 4) for k, v in med.items()}}))
