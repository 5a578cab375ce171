#!/usr/bin/env python
"""One frame of a bench workload through the C ABI, nothing else: what bench.py runs under
`ncu --metrics smsp__inst_executed.sum,...` to COUNT the warp-instructions of the trace kernels (no timing is taken from this run).
Also usable as the target of a full `ncu --set full` capture of the trace kernels."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="inst10m")
    ap.add_argument("--width", type=int, default=0)
    ap.add_argument("--height", type=int, default=0)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--build-flags", type=int, default=0)
    a = ap.parse_args()
    import torch
    import bench
    from build_up_phase_b200 import rtcore
    scene = bench.make_workload(a.workload, a.width, a.height)
    dev = torch.device("cuda", 0)
    ctx = rtcore.Context(0)
    blases = ctx.build_blas_batch(scene.blases, flags=a.build_flags) if len(scene.blases) > 1 else [ctx.build_blas(scene.blases[0], flags=a.build_flags)]
    tlas = ctx.build_tlas(scene.instances, blases)
    ctx.set_hit_records(scene.hit_records)
    ctx.set_miss_color(scene.miss_color)
    cam = ctx.camera(scene.camera_pos, scene.yfov_deg)
    frame = torch.zeros((scene.height, scene.width, 4), dtype=torch.uint8, device=dev)
    for _ in range(a.frames):
        ctx.trace_device(tlas, cam, scene.width, scene.height, scene.bounces, frame)
    torch.cuda.synchronize()
    print("frame_once", a.workload, int(frame.sum().item()))


if __name__ == "__main__":
    main()
