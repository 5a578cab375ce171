#!/usr/bin/env python
"""Tree quality as a measured quantity: the SAH cost of the LBVH the product builds (C_traversal = C_intersection = 1):
    cost = sum over reachable internal nodes of A(node) / A(root)  +  sum over leaves of A(leaf) / A(root) * triangles(leaf)
per BLAS, from the exported 64-byte nodes (level-by-level walk from the root, so collapsed Karras slots are never read), next to
the per-ray traversal counters of a frame. One JSON line per workload."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench  # noqa: E402
from build_up_phase_b200 import rtcore  # noqa: E402


def half_area(lo, hi):
    d = np.maximum(hi.astype(np.float64) - lo.astype(np.float64), 0.0)
    return 2.0 * (d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0])


def sah(nodes_u32, root, root_lo, root_hi):
    if not (0 <= root < 0x7FFFFFF0):
        return 1.0, 0, 1, 0           # a single leaf
    f = nodes_u32.view(np.float32)
    r = nodes_u32.view(np.int32)
    a_root = half_area(np.asarray([root_lo], dtype=np.float32), np.asarray([root_hi], dtype=np.float32))[0]
    cost_nodes, cost_leaves, n_nodes, n_leaves, depth = 1.0, 0.0, 1, 0, 0
    frontier = np.asarray([root], dtype=np.int64)
    while frontier.size:
        depth += 1
        nxt = []
        for h in (0, 1):
            lo = f[frontier, 8 * h:8 * h + 3]
            hi = f[frontier, 8 * h + 3:8 * h + 6]
            ref = r[frontier, 8 * h + 6]
            a = half_area(lo, hi) / a_root
            internal = (ref >= 0) & (ref < 0x7FFFFFF0)
            leaf = ref < 0
            cost_nodes += a[internal].sum()
            cnt = ((~ref[leaf]) & 7) + 1
            cost_leaves += (a[leaf] * cnt).sum()
            n_nodes += int(internal.sum()); n_leaves += int(leaf.sum())
            nxt.append(ref[internal].astype(np.int64))
        frontier = np.concatenate(nxt)
    return cost_nodes + cost_leaves, n_nodes, n_leaves, depth


def main():
    workloads = sys.argv[1:] or ["inst10m", "tess1m", "soup10m"]
    with rtcore.Context(0) as ctx:
        for w in workloads:
            scene = bench.make_workload(w, 0, 0)
            blases = ctx.build_blas_batch(scene.blases) if len(scene.blases) > 1 else [ctx.build_blas(scene.blases[0])]
            costs, nn, nl, dep = [], 0, 0, 0
            sample = blases if len(blases) <= 64 else blases[::16]
            for b in sample:
                info = b.info()
                nodes, _ = b.export()
                c, n, l, d = sah(nodes, info.root_ref, list(info.bounds_lo), list(info.bounds_hi))
                costs.append(c); nn += n; nl += l; dep = max(dep, d)
            tris = sum(b.info().triangle_count for b in sample)
            print(json.dumps({"workload": w, "blases": len(blases), "blases_sampled": len(sample), "triangles_sampled": tris,
                              "sah_cost_mean": float(np.mean(costs)), "sah_cost_min": float(np.min(costs)), "sah_cost_max": float(np.max(costs)),
                              "internal_nodes_per_triangle": nn / tris, "leaves_per_triangle": nl / tris, "max_depth": dep}), flush=True)
            for b in blases:
                b.free()


if __name__ == "__main__":
    main()
