#!/usr/bin/env python
"""RT_BUILD_MODE_REFIT vs a full rebuild on an animated 1M-triangle height field (the tess1m grid whose bumps drift away from the
pose the tree was sorted for): update time (CUDA events) and the trace rate of the resulting tree, per frame.
Prints one JSON line per frame; tools/gpu_r2i.sh stores them under gpurun_out/."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import rtcore, scenes  # noqa: E402

W, H = 3840, 2160
base = scenes.heightfield(1000, 500, -8.0, 8.0, -4.0, 4.0, 1.0, 1)


def frame(f):
    v = base.vertices.copy()
    # the surface swells and a travelling wave crosses it: vertices move by up to ~0.6 (40 triangle widths) at f = 8
    v[:, 2] = (v[:, 2] * np.float32(1.0 + 0.1 * f) + np.float32(0.08 * f) * np.sin(v[:, 0] * np.float32(1.5) + np.float32(0.7 * f))).astype(np.float32)
    v[:, 0] = (v[:, 0] + np.float32(0.03 * f) * np.cos(v[:, 1] * np.float32(2.0))).astype(np.float32)
    return scenes.Geometry(np.ascontiguousarray(v), base.indices, None)


def mrays(ctx, tlas, cam, fb, reps=5):
    best = 1e9
    st = None
    for r in range(reps + 1):
        ctx.trace_device(tlas, cam, W, H, 1, fb, stats=(r == 0))
        if r == 0:
            st = ctx.trace_stats()
        else:
            best = min(best, ctx.trace_ms())
    rays = st["rays_primary"] + st["rays_secondary"]
    return rays / (best * 1e-3) / 1e6, st


with rtcore.Context(0) as ctx:
    inst = [scenes.Instance(scenes.IDENTITY_3X4.copy(), 3, 0xFF, 0, 1, 0)]
    ctx.set_hit_records(scenes.SAMPLE_HIT_RECORDS[1:2].copy())
    cam = ctx.camera((0.0, 0.0, 10.0), 60.0)
    fb = torch.zeros((H, W, 4), dtype=torch.uint8, device="cuda")
    refit_blas = ctx.build_blas([frame(0)], flags=rtcore.RT_BUILD_ALLOW_UPDATE)
    full_blas = ctx.build_blas([frame(0)])
    for f in (0, 1, 2, 4, 8):
        geo = frame(f)
        t_refit, t_full = [], []
        for _ in range(3):
            ctx.update_blas(refit_blas, [geo], flags=rtcore.RT_BUILD_ALLOW_UPDATE | rtcore.RT_BUILD_MODE_REFIT)
            t_refit.append(ctx.build_timing()["total_ms"])
            ctx.update_blas(full_blas, [geo])
            t_full.append(ctx.build_timing()["total_ms"])
        out = {"frame": f, "refit_ms": min(t_refit), "rebuild_ms": min(t_full)}
        for name, b in (("refit", refit_blas), ("rebuild", full_blas)):
            tlas = ctx.build_tlas(inst, [b])
            m, st = mrays(ctx, tlas, cam, fb)
            out[f"{name}_mrays"] = m
            out[f"{name}_nodes_per_ray"] = st["nodes_visited"] / (st["rays_primary"] + st["rays_secondary"])
            out[f"{name}_tris_per_ray"] = st["triangles_tested"] / (st["rays_primary"] + st["rays_secondary"])
            out[f"{name}_crc"] = int(fb.to(torch.int64).sum().item())
            tlas.free()
        assert out["refit_crc"] == out["rebuild_crc"], "a refitted tree and a rebuilt tree must render the same frame"
        print(json.dumps(out), flush=True)
