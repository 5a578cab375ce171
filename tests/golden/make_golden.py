#!/usr/bin/env python
"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle (brute-force mode: every triangle of
every instance is tested for every ray, no BVH involved). The reference ships no golden images (parity
unpinned), so these are OUR pins: they freeze the oracle's answers for the reference's own scene at its own resolution
and for one fuzz scene with a bounce, so that neither side can drift unnoticed. Run from the repo root:
    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)

from build_up_phase_b200 import scenes  # noqa: E402
import oracle_binding as ob  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def golden_scenes():
    return {
        "sample_1200x800": scenes.sample_scene(1200, 800),                     # BASELINE configs[0]: the sample's default resolution
        "triangle_640x400": scenes.single_triangle_scene(640, 400),            # the literal single-triangle BLAS/TLAS
        "fuzz_seed7_320x200": scenes.random_scene(n_blas=3, tris_per_blas=400, n_instances=9, seed=7, width=320, height=200,
                                                  bounces=1, shared_edges=True),
    }


def main():
    for name, scene in golden_scenes().items():
        o = ob.OracleScene(scene)
        rgba, prim, sec, st = o.trace(mode=ob.MODE_BRUTE)
        o.close()
        np.savez_compressed(os.path.join(HERE, name + ".npz"), rgba=rgba, primary=prim, secondary=sec,
                            stats=np.array([st[k] for k in ("rays_primary", "rays_secondary", "primary_hits", "secondary_hits", "near_edge_hits")], dtype=np.uint64))
        print(name, rgba.shape, st["primary_hits"], st["secondary_hits"], os.path.getsize(os.path.join(HERE, name + ".npz")), "bytes")


if __name__ == "__main__":
    main()
