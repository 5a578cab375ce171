"""CPU check of the product's 8-wide BVH logic (build-up-phase_b200/csrc/wide_bvh.cuh compiled as plain C++ by
tests/wide_host.cpp): collapse the oracle's binary LBVH, check the structural invariants, and check that for every
primary ray the triangle the brute-force oracle reports as closest hit is REACHED by the wide traversal (the
conservative quantised boxes and the group/bit bookkeeping never lose a hit)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from build_up_phase_b200 import scenes
from parity import check_wide_bvh

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def wide():
    src = os.path.join(HERE, "wide_host.cpp")
    hdr = os.path.join(HERE, "..", "build-up-phase_b200", "csrc", "wide_bvh.cuh")
    out = os.path.join(HERE, "_build", "libwide_host.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-ffp-contract=off", "-o", out, src])
    L = C.CDLL(out)
    vp, u32 = C.c_void_p, C.c_uint32
    L.wide_host_build.argtypes = [vp, C.c_int32, vp, vp, u32, vp, u32, vp, C.POINTER(u32), C.POINTER(u32)]
    L.wide_host_reach.argtypes = [vp, u32, vp, vp, vp, u32, C.c_float, C.c_float, vp, vp]
    return L


def collapse(L, onodes, oinfo, otris):
    n = oinfo.triangle_count
    cap = n + 2
    wn = np.zeros((cap, 20), dtype=np.uint32)
    perm = np.zeros(n, dtype=np.uint32)
    lo = np.array(list(oinfo.bounds_lo), dtype=np.float32)
    hi = np.array(list(oinfo.bounds_hi), dtype=np.float32)
    nn, depth = C.c_uint32(), C.c_uint32()
    rc = L.wide_host_build(onodes.ctypes.data, oinfo.root_ref, lo.ctypes.data, hi.ctypes.data, n, wn.ctypes.data, cap,
                           perm.ctypes.data, C.byref(nn), C.byref(depth))
    assert rc == 0, rc
    tris_wide = otris[perm].copy()
    tris_wide[:, 11] = perm
    return wn[:nn.value].copy(), tris_wide, perm, depth.value, lo, hi


@pytest.mark.parametrize("case", ["tess", "soup", "sample", "tri"])
def test_wide_collapse_and_reach(wide, oracle, case):
    W, H = 160, 96
    if case == "tess":
        scene = scenes.tess_scene(nx=60, ny=40, width=W, height=H, bounces=0)
    elif case == "soup":
        g = scenes.soup_part(6000, 0, 1, seed=3, edge=0.4, split="index")
        scene = scenes.Scene("soup", [[g]], [scenes.Instance(scenes.IDENTITY_3X4.copy(), 1, 0xFF, 0, 1, 0)], scenes.SAMPLE_HIT_RECORDS[:1].copy(),
                             width=W, height=H)
    elif case == "sample":
        s0 = scenes.sample_scene(W, H)
        scene = scenes.Scene("sample1", s0.blases, [scenes.Instance(scenes.IDENTITY_3X4.copy(), 1, 0xFF, 0, 1, 0)], s0.hit_records, width=W, height=H)
    else:
        scene = scenes.single_triangle_scene(W, H)
    o = oracle.OracleScene(scene)
    oinfo, onodes, otris, okeys, oprims = o.blas_export(0)
    wn, tris_wide, perm, depth, lo, hi = collapse(wide, onodes, oinfo, otris)
    n_nodes = check_wide_bvh(wn, tris_wide, 0, lo, hi)
    assert n_nodes == wn.shape[0]
    # brute-force closest hits of the primary rays (identity instance: object space == world space)
    rgba, prim, _, st = o.trace(mode=oracle.MODE_BRUTE)
    # rays exactly like the raygen shader (main.cpp:1033-1046)
    aspect_y = np.float32(np.tan(np.float32(np.float32(scene.yfov_deg) * np.float32(0.017453292519943295)) * np.float32(0.5)))
    aspect_x = np.float32(aspect_y * np.float32(W) / np.float32(H))
    xs = (np.arange(W, dtype=np.float32) + np.float32(0.5)) / np.float32(W) * np.float32(2) - np.float32(1)
    ys = (np.arange(H, dtype=np.float32) + np.float32(0.5)) / np.float32(H) * np.float32(2) - np.float32(1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    rays = np.zeros((H, W, 6), dtype=np.float32)
    rays[..., 0:3] = scene.camera_pos
    rays[..., 3] = X * aspect_x
    rays[..., 4] = -(Y * aspect_y)
    rays[..., 5] = -1.0
    # expected primitive in wide position: find the sorted position of (geometry, primitive), then its wide slot
    geo_prim_sorted = otris[:, 9].astype(np.int64) << 32 | otris[:, 10].astype(np.int64)
    lut = {int(k): i for i, k in enumerate(geo_prim_sorted)}
    inv = np.zeros(perm.shape[0], dtype=np.int64)
    inv[perm] = np.arange(perm.shape[0])
    hit = prim["instance_id"] != 0xFFFFFFFF
    expect = np.full((H, W), -1, dtype=np.int32)
    keys = (prim["geometry_index"].astype(np.int64) << 32 | prim["primitive_id"].astype(np.int64))
    for y, x in zip(*np.nonzero(hit)):
        expect[y, x] = inv[lut[int(keys[y, x])]]
    absmax = np.maximum(np.abs(lo), np.abs(hi)).astype(np.float32)
    tested = np.zeros(H * W, dtype=np.uint32)
    visited = np.zeros(H * W, dtype=np.uint32)
    rays = np.ascontiguousarray(rays.reshape(-1, 6))
    expect = np.ascontiguousarray(expect.reshape(-1))
    missing = wide.wide_host_reach(wn.ctypes.data, 0, absmax.ctypes.data, rays.ctypes.data, expect.ctypes.data, H * W,
                                   C.c_float(0.0), C.c_float(100.0), tested.ctypes.data, visited.ctypes.data)
    print(case, "wide nodes", n_nodes, "depth", depth, "hits", int(hit.sum()), "missing", missing,
          "avg nodes/ray %.1f tris/ray %.1f (no distance culling)" % (visited.mean(), tested.mean()))
    assert hit.sum() > 0
    assert missing == 0
