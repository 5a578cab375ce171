"""Self-made pins of the oracle (SURVEY §8(c) "additional self-made pins"): brute force vs its own
LBVH traversal on random scenes, watertightness on a tessellated mesh, LBVH invariants, hash/Morton
known values."""
import numpy as np
import pytest

from build_up_phase_b200 import scenes
from parity import MISS, walk_compare_bvh


@pytest.mark.parametrize("seed", [1, 2])
def test_brute_force_equals_bvh_random(oracle, seed):
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=8, seed=seed, width=200, height=120, bounces=1,
                                shared_edges=(seed % 2 == 1))
    o = oracle.OracleScene(scene)
    a = o.trace(mode=oracle.MODE_BRUTE)
    b = o.trace(mode=oracle.MODE_BVH)
    assert a[1].tobytes() == b[1].tobytes(), "primary hits differ between brute force and BVH traversal"
    assert a[2].tobytes() == b[2].tobytes(), "secondary hits differ"
    assert np.array_equal(a[0], b[0])
    assert a[3]["primary_hits"] > 500 and a[3]["secondary_hits"] > 20
    assert b[3]["triangles_tested"] < a[3]["triangles_tested"] / 10


def test_watertight_tessellated_mesh(oracle):
    """A height field covering the whole view: every primary ray must hit (no leaks through the
    millions of shared edges/vertices), including rays aimed exactly at lattice vertices."""
    scene = scenes.tess_scene(nx=64, ny=64, width=513, height=513, bounces=0)
    g = scene.blases[0][0]
    g.vertices[:, 0] *= 4.0       # [-32,32] x [-16,16]: covers the 60-degree frustum at z ~ 0 from z = 10
    g.vertices[:, 1] *= 8.0
    o = oracle.OracleScene(scene)
    rgba, prim, _, st = o.trace(mode=oracle.MODE_BVH)
    assert st["primary_hits"] == 513 * 513
    b = o.trace(mode=oracle.MODE_BRUTE, rows=(0, None, 16))
    rows = slice(0, None, 16)
    assert prim[rows].tobytes() == b[1][rows].tobytes()


def test_lbvh_invariants(oracle):
    scene = scenes.tess_scene(nx=50, ny=40, width=8, height=8, bounces=0)
    o = oracle.OracleScene(scene)
    info, nodes, tris, keys, prims = o.blas_export(0)
    n = info.triangle_count
    assert n == 4000
    assert np.all(keys[1:] >= keys[:-1])
    assert np.array_equal(np.sort(prims), np.arange(n, dtype=np.uint32))
    # stability: equal keys keep primitive order
    same = keys[1:] == keys[:-1]
    assert np.all(prims[1:][same] > prims[:-1][same])
    assert walk_compare_bvh(nodes, info.root_ref, nodes, info.root_ref) > n // 8
    # every triangle in exactly one leaf
    covered = np.zeros(n, dtype=np.int32)
    stack = [info.root_ref]
    while stack:
        r = stack.pop()
        for half in (0, 1):
            ref = int(np.int32(nodes[r, 8 * half + 6]))
            if ref < 0:
                x = ~ref
                covered[(x >> 3):(x >> 3) + (x & 7) + 1] += 1
            else:
                stack.append(ref)
    assert np.all(covered == 1)


def test_hash_and_morton_known_values(oracle):
    L = oracle.lib()
    # PCG hash: scalar C == numpy generator used by scenes.py
    vals = np.array([0, 1, 2, 12345, 0xFFFFFFFF], dtype=np.uint32)
    assert [L.orc_pcg_hash(int(v)) for v in vals] == [int(x) for x in scenes.pcg_hash(vals)]
    import ctypes as C
    lo = (C.c_float * 3)(0, 0, 0)
    hi = (C.c_float * 3)(1, 1, 1)
    assert L.orc_morton30(0, 0, 0, lo, hi) == 0
    assert L.orc_morton30(1, 1, 1, lo, hi) == 0x3FFFFFFF              # clamped to cell 1023 on every axis
    assert L.orc_morton30(0.5, 0, 0, lo, hi) == 0b100 << 27           # x is the most significant interleaved bit
    assert L.orc_morton30(0, 0.5, 0, lo, hi) == 0b010 << 27
    assert L.orc_morton30(0, 0, 0.5, lo, hi) == 0b001 << 27


def test_scene_generators_sizes():
    assert scenes.tess_scene(nx=100, ny=50).triangle_count == 10000
    s = scenes.instanced_scene(n_side=2, quads=70)
    assert s.triangle_count == 4 * 9800 and len(s.instances) == 4
    assert scenes.soup_scene(1000).triangle_count == 1000
    # the full-size definitions (not generated here): 1000x500 quads -> 1,000,000; 1024 x 9,800 -> 10,035,200
    assert 1000 * 500 * 2 == 1_000_000 and 32 * 32 * 2 * 70 * 70 == 10_035_200
