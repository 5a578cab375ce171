#!/usr/bin/env python
"""Times rt_update_tlas (the per-frame path of animated instances) for the inst10m instance set: segmented single-kernel sort
vs the global onesweep sort (RT_BUILD_NO_SEGMENTED_SORT). Prints min / median CUDA-event milliseconds of 30 updates each."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from build_up_phase_b200 import rtcore, scenes  # noqa: E402

scene = scenes.instanced_scene(n_side=32, quads=4, width=64, height=64, bounces=0)
with rtcore.Context(0) as ctx:
    sh = rtcore.SceneHandles(ctx, scene)
    arr = ctx.instance_array(scene.instances, sh.blases)
    for name, flags in (("segmented", 0x4), ("onesweep", 0x4 | 0x400)):
        ts = []
        for _ in range(30):
            ctx._check(ctx.L.rt_update_tlas(ctx.h, sh.tlas.handle, C.addressof(arr), len(scene.instances), flags))
            ts.append(ctx.build_timing()["total_ms"])
        print(f"{name}: n={len(scene.instances)} min={min(ts):.4f} ms median={float(np.median(ts)):.4f} ms")
    sh.free()
