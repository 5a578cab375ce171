"""Pins the CPU oracle against the known answers of the reference sample scene.

The reference ships no tests or golden images ("parity unpinned"), so these are the analytically derived
facts any conformant Vulkan driver must show for vulkan-raytracing-basic/main.cpp's scene:
four 2x2 quads in the z=0 plane seen from (0,0,10) with a 60-degree vertical fov at 1200x800.
"""
import numpy as np
import pytest

from build_up_phase_b200 import scenes

MISS = 0xFFFFFFFF


@pytest.fixture(scope="module")
def traced(oracle):
    s = scenes.sample_scene()
    o = oracle.OracleScene(s)
    out = {m: o.trace(mode=m) for m in (oracle.MODE_BRUTE, oracle.MODE_BVH)}
    return s, o, out


def test_aspect_constants(oracle):
    ay = oracle.lib().orc_aspect_y(60.0)
    assert np.float32(ay) == np.float32(0.5773503)                      # tan(30 deg) in fp32
    assert np.float32(ay) * np.float32(1200) / np.float32(800) == np.float32(0.86602545)


def test_hit_counts_and_pixel_boxes(traced, oracle):
    _, _, out = traced
    rgba, prim, sec, st = out[oracle.MODE_BRUTE]
    hit = prim["instance_id"] != MISS
    assert hit.sum() == 77284 and (~hit).sum() == 882716
    assert st["primary_hits"] == 77284 and st["rays_primary"] == 960000
    boxes = {(0, 0): (392, 530, 192, 330), (0, 1): (669, 807, 192, 330),
             (1, 0): (392, 530, 469, 607), (1, 1): (669, 807, 469, 607)}
    for (inst, geo), (x0, x1, y0, y1) in boxes.items():
        m = hit & (prim["instance_id"] == inst) & (prim["geometry_index"] == geo)
        assert m.sum() == 139 * 139 == 19321
        ys, xs = np.nonzero(m)
        assert (xs.min(), xs.max(), ys.min(), ys.max()) == (x0, x1, y0, y1)
        # every hit is on the z = 0 plane 10 units along the un-normalised direction (z = -1)
        assert np.all(np.abs(prim["t"][m] - 10.0) <= 4e-6)
        assert np.all(prim["custom_index"][m] == 100)
        # both triangles of the quad are seen, split along the v1-v3 diagonal
        p0, p1 = (prim["primitive_id"][m] == 0).sum(), (prim["primitive_id"][m] == 1).sum()
        assert p0 + p1 == 19321 and 9500 < p0 < 9800


def test_colours(traced, oracle):
    _, _, out = traced
    rgba, prim, _, _ = out[oracle.MODE_BRUTE]
    assert tuple(rgba[0, 0]) == (0, 0, 51, 0)                           # miss (0,0,0.2), alpha 0
    expect = {(0, 0): (153, 26, 51, 0), (0, 1): (26, 204, 102, 0), (1, 0): (230, 178, 26, 0)}
    for (inst, geo), col in expect.items():
        m = (prim["instance_id"] == inst) & (prim["geometry_index"] == geo)
        assert np.all(rgba[m] == np.array(col, dtype=np.uint8))
    # record 3 except the barycentric special case of the closest-hit shader (prim 1, inst 1, geo 1, customIndex 100)
    m = (prim["instance_id"] == 1) & (prim["geometry_index"] == 1) & (prim["primitive_id"] == 0)
    assert np.all(rgba[m] == np.array((76, 153, 230, 0), dtype=np.uint8))
    h = prim[509, 766]
    assert (h["instance_id"], h["geometry_index"], h["primitive_id"]) == (1, 1, 1)
    assert abs(h["u"] - 0.4114) < 2e-4 and abs(h["v"] - 0.2984) < 2e-4
    w0 = np.float32(1.0) - h["u"] - h["v"]
    exp = [int(np.rint(np.float32(c) * np.float32(255))) for c in (w0, h["u"], h["v"])]
    assert list(rgba[509, 766][:3]) == exp and rgba[509, 766][3] == 0


def test_unorm8_ties(oracle):
    # four SBT components are exact .5 ties in fp32; RNE decides them
    import ctypes as C
    out = (C.c_uint8 * 4)()
    for c, e in ((0.1, 26), (0.9, 230), (0.7, 178), (0.3, 76), (0.2, 51), (1.5, 255), (-1.0, 0), (float("nan"), 0)):
        rgb = (C.c_float * 3)(c, c, c)
        oracle.lib().orc_unorm8(rgb, out)
        assert out[0] == e, (c, out[0], e)


def test_brute_equals_bvh(traced, oracle):
    _, _, out = traced
    a, b = out[oracle.MODE_BRUTE], out[oracle.MODE_BVH]
    assert np.array_equal(a[0], b[0])
    assert a[1].tobytes() == b[1].tobytes()


def test_1920x1080(oracle):
    s = scenes.sample_scene(1920, 1080)
    o = oracle.OracleScene(s)
    rgba, prim, _, st = o.trace(mode=oracle.MODE_BRUTE)
    hit = prim["instance_id"] != MISS
    # each quad spans 2 world units = 2/(2*10*tan30) * 1080 = 187.06 px in both directions
    for inst in (0, 1):
        for geo in (0, 1):
            m = hit & (prim["instance_id"] == inst) & (prim["geometry_index"] == geo)
            ys, xs = np.nonzero(m)
            assert xs.max() - xs.min() + 1 == 187 and ys.max() - ys.min() + 1 == 187
    assert st["primary_hits"] == 4 * 187 * 187


def test_single_triangle(oracle):
    s = scenes.single_triangle_scene(600, 400)
    o = oracle.OracleScene(s)
    rgba, prim, _, st = o.trace(mode=oracle.MODE_BRUTE)
    hit = prim["instance_id"] != MISS
    # triangle (-1,-1) (1,-1) (0,1): area 2 world units^2; one pixel = (2*10*tan30/400)^2
    px = 2 * 10 * np.tan(np.radians(30)) / 400
    assert abs(hit.sum() - 2.0 / px ** 2) < 100      # boundary pixels: ~perimeter/2
    assert np.all(prim["custom_index"][hit] == 7)
    assert np.all(rgba[hit] == np.array((153, 26, 51, 0), dtype=np.uint8))
