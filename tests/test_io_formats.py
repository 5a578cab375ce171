"""Host-side formats either side of the path (include/rtcore_io.h): the library's .obj loader
against an independent pure-Python statement of the same grammar (bit-exact arrays), and the PPM / sRGB export.
CPU-only: no kernel is involved."""
import os

import numpy as np
import pytest


def py_parse_obj(text: str):
    """Independent restatement of the documented grammar (include/rtcore_io.h): v / f / o / g, 1-based and negative
    indices, v/vt/vn forms, polygon fans, backslash continuation; double -> float32 rounding of coordinates."""
    verts, tris, groups = [], [], []
    lines, cur = [], ""
    for raw in text.split("\n"):
        raw = raw.rstrip(" \t\r\v\f")
        if raw.endswith("\\"):
            cur += raw[:-1] + " "
            continue
        lines.append(cur + raw)
        cur = ""
    if cur:
        lines.append(cur)

    def open_group(name):
        if groups and groups[-1][2] == 0:
            groups[-1][0] = name
        else:
            groups.append([name, len(tris), 0])

    for ln in lines:
        s = ln.strip(" \t\r\v\f")
        if not s or s.startswith("#"):
            continue
        parts = s.split()
        kw = parts[0]
        if kw == "v":
            verts.append([np.float32(float(c)) for c in parts[1:4]])
        elif kw == "f":
            ids = []
            for tok in parts[1:]:
                if tok.startswith("#"):
                    break
                v = int(tok.split("/")[0])
                ids.append(v - 1 if v > 0 else len(verts) + v)
            if not groups:
                open_group("")
            for i in range(1, len(ids) - 1):
                tris.append([ids[0], ids[i], ids[i + 1]])
                groups[-1][2] += 1
        elif kw in ("o", "g"):
            open_group(s[1:].strip(" \t\r\v\f"))
    if groups and groups[-1][2] == 0:
        groups.pop()
    return (np.array(verts, dtype=np.float32).reshape(-1, 3), np.array(tris, dtype=np.uint32).reshape(-1, 3),
            [(g[0], g[1], g[2]) for g in groups])


SAMPLE_OBJ = """# the reference sample's quad (main.cpp:676-695) as an .obj, plus format corner cases
o quad
v -1 -1 0
v 1 -1 0
v 1 1 0 1.0
v -1.0e0 +1 0
vt 0 0
vn 0 0 1
f 1 2 4
f 2/1 3/1 4/1
g fan polygon
v 0.1 0.2 0.30000001192092896
v 3.4028234e38 -1e-45 0.1
v 2.5 2.5 2.5
f -3//1 -2//1 -1//1 1/1/1 \\
  2
usemtl whatever
s off
g empty_then_replaced
o last   \r
f 5 6 7   # trailing comment
"""


def test_obj_parser_matches_python_statement(rt, tmp_path):
    m = rt.ObjMesh(text=SAMPLE_OBJ.encode())
    v, t, g = py_parse_obj(SAMPLE_OBJ)
    assert np.array_equal(m.vertices.view(np.uint32), v.view(np.uint32))
    assert np.array_equal(m.indices, t)
    assert m.groups() == g == [("quad", 0, 2), ("fan polygon", 2, 3), ("last", 5, 1)]
    assert t.tolist()[:2] == [[0, 1, 3], [1, 2, 3]]               # the sample's index buffer {0,1,3,1,2,3}
    # through a file, with a larger random mesh (polygons of 3..6 vertices, mixed reference forms)
    rng = np.random.default_rng(5)
    nv = 400
    lines = [f"v {x!r} {y!r} {z!r}" for x, y, z in rng.normal(size=(nv, 3)).astype(np.float64).tolist()]
    for k in range(300):
        if k % 100 == 0:
            lines.append(f"g part{k // 100}")
        n = int(rng.integers(3, 7))
        ids = rng.integers(1, nv + 1, size=n)
        form = k % 4
        toks = [str(i) if form == 0 else f"{i}/{i}" if form == 1 else f"{i}//{i}" if form == 2 else str(int(i) - nv - 1) for i in ids]
        lines.append("f " + " ".join(toks))
    text = "\n".join(lines) + "\n"
    path = tmp_path / "mesh.obj"
    path.write_text(text)
    m2 = rt.ObjMesh(path=str(path))
    v2, t2, g2 = py_parse_obj(text)
    assert np.array_equal(m2.vertices.view(np.uint32), v2.view(np.uint32))
    assert np.array_equal(m2.indices, t2) and m2.groups() == g2 and len(g2) == 3
    cg = m2.c_geometry(1)
    assert cg.triangle_count == g2[1][2] and cg.vertex_count == nv and cg.vertex_stride_bytes == 12


def test_obj_errors(rt, tmp_path):
    for bad, what in [("v 1 2\n", "three coordinates"), ("v 0 0 0\nf 1 2 3\n", "out of range"), ("v 0 0 0\nf 1 1\n", "at least three"),
                      ("v 0 0 0\nf 1 x 1\n", "bad face"), ("v 0 0 0\nf 0 1 1\n", "out of range")]:
        with pytest.raises(rt.RtError) as e:
            rt.ObjMesh(text=bad.encode())
        assert what in str(e.value) and "line" in str(e.value)
    with pytest.raises(rt.RtError):
        rt.ObjMesh(path=str(tmp_path / "missing.obj"))
    empty = rt.ObjMesh(text=b"# nothing\n")
    assert empty.vertices.shape == (0, 3) and empty.indices.shape == (0, 3) and empty.groups() == []


def test_ppm_and_srgb(rt, tmp_path):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, size=(7, 13, 4), dtype=np.uint8)
    p = tmp_path / "a.ppm"
    rt.write_ppm(str(p), img)
    raw = p.read_bytes()
    assert raw.startswith(b"P6\n13 7\n255\n")
    body = np.frombuffer(raw[len(b"P6\n13 7\n255\n"):], dtype=np.uint8).reshape(7, 13, 3)
    assert np.array_equal(body, img[:, :, :3])
    # sRGB encode (what storing into a *_SRGB swapchain does, main.cpp:50) + flip
    lut = rt.srgb8_table()
    i = np.arange(256) / 255.0
    ref = np.where(i <= 0.0031308, 12.92 * i, 1.055 * np.power(i, 1 / 2.4) - 0.055)
    assert np.array_equal(lut, np.floor(255.0 * ref + 0.5).astype(np.uint8))
    assert lut[0] == 0 and lut[255] == 255 and np.all(np.diff(lut.astype(int)) >= 0)
    rt.write_ppm(str(p), img, rt.IMAGE_SRGB_ENCODE | rt.IMAGE_FLIP_Y)
    body = np.frombuffer(p.read_bytes()[len(b"P6\n13 7\n255\n"):], dtype=np.uint8).reshape(7, 13, 3)
    assert np.array_equal(body, lut[img[::-1, :, :3]])


def test_png(rt, tmp_path):
    """rt_write_png against a minimal independent decoder (chunk walk, CRCs, zlib inflate, filter 0)."""
    import struct
    import zlib
    rng = np.random.default_rng(4)
    for (h, w) in ((5, 9), (300, 401)):          # the second one needs several stored deflate blocks (> 65535 bytes of scanlines)
        img = rng.integers(0, 256, size=(h, w, 4), dtype=np.uint8)
        p = tmp_path / "a.png"
        rt.write_png(str(p), img, rt.IMAGE_FLIP_Y)
        raw = p.read_bytes()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(raw):
            n, typ = struct.unpack(">I4s", raw[pos:pos + 8])
            data = raw[pos + 8:pos + 8 + n]
            assert struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])[0] == zlib.crc32(typ + data)
            chunks.append((typ, data))
            pos += 12 + n
        assert [c[0] for c in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (w, h, 8, 2, 0, 0, 0)
        lines = np.frombuffer(zlib.decompress(chunks[1][1]), dtype=np.uint8).reshape(h, 1 + 3 * w)
        assert np.all(lines[:, 0] == 0)
        assert np.array_equal(lines[:, 1:].reshape(h, w, 3), img[::-1, :, :3])
