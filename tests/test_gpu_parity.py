"""GPU parity tests proper: the CUDA path (through the C ABI, host buffers) against the CPU oracle.

Bar: hit/miss masks and (instance, geometry, primitive, customIndex) ids bit-exact, t/u/v bit-exact,
RGBA8 within +-1 LSB. Rays within 2^-20 (barycentric) of a shared edge are counted and reported.
"""
import numpy as np
import pytest

from build_up_phase_b200 import scenes
from parity import MISS, assert_parity, compare_hits, walk_compare_bvh, walk_compare_bvh_renumbered

pytestmark = pytest.mark.gpu


def _run(rt, ctx, oracle, scene, mode, rows=(0, None, 1), batch=True, stats=False):
    sh = rt.SceneHandles(ctx, scene, batch=batch)
    try:
        g = sh.trace(want_hits=True, stats=stats)
        gstats = ctx.trace_stats() if stats else None
    finally:
        sh.free()
    o = oracle.OracleScene(scene)
    r = o.trace(mode=mode, rows=rows)
    o.close()
    return g, r, gstats


@pytest.mark.parametrize("wh", [(1200, 800), (1920, 1080)])
def test_sample_scene(rt, ctx, oracle, wh):
    """BASELINE configs[0] and configs[1]: the samples' scene at 1200x800 and 1920x1080."""
    scene = scenes.sample_scene(*wh)
    g, r, gs = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE, stats=True)
    rp, _, rc = assert_parity(g, r, what=f"sample{wh}")
    rgba, prim, _ = g
    hit = prim["instance_id"] != MISS
    if wh == (1200, 800):    # the known answers of SURVEY §8(c), checked on the GPU output itself
        assert hit.sum() == 77284
        assert tuple(rgba[0, 0]) == (0, 0, 51, 0)
        m = hit & (prim["instance_id"] == 0) & (prim["geometry_index"] == 0)
        ys, xs = np.nonzero(m)
        assert (xs.min(), xs.max(), ys.min(), ys.max()) == (392, 530, 192, 330)
        assert np.all(rgba[m] == np.array((153, 26, 51, 0), dtype=np.uint8))
    assert gs["rays_primary"] == wh[0] * wh[1] and gs["primary_hits"] == hit.sum()
    if wh == (1200, 800):    # swapchain byte order on request (main.cpp:50): the same image with R and B exchanged
        sh = rt.SceneHandles(ctx, scene)
        bgra, _, _ = ctx.trace(sh.tlas, sh.cam, wh[0], wh[1], 0, bgra=True)
        sh.free()
        assert np.array_equal(bgra[..., [2, 1, 0, 3]], rgba)
    assert gs["near_edge_hits"] == r[3]["near_edge_hits"]
    print("sample", wh, rp, rc)


def test_single_triangle(rt, ctx, oracle):
    scene = scenes.single_triangle_scene(640, 400)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE)
    assert_parity(g, r, what="triangle")


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_scenes_vs_brute_force(rt, ctx, oracle, seed):
    """Fuzz: several BLASes x geometries x transformed instances, masks, one diffuse bounce; the oracle
    tests EVERY triangle of EVERY instance for every ray (no BVH on the reference side)."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=200 + 150 * seed, n_instances=9, seed=seed, width=320, height=200,
                                bounces=1, shared_edges=(seed % 2 == 1))
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE, batch=(seed % 2 == 0))
    rp, rs, rc = assert_parity(g, r, what=f"random{seed}")
    assert rp["hits"] > 1000 and rs["hits"] > 50
    print("random", seed, rp, rs, rc)


def test_tess_mesh_shared_edges(rt, ctx, oracle):
    """Tessellated height field (every interior edge shared by two triangles): brute force at low
    resolution, BVH oracle at higher resolution, one bounce."""
    scene = scenes.tess_scene(nx=60, ny=30, width=400, height=224, bounces=1)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE)
    rp, rs, rc = assert_parity(g, r, what="tess-brute")
    print("tess-brute", rp, rs)
    scene = scenes.tess_scene(nx=400, ny=200, width=1280, height=720, bounces=1)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BVH)
    rp, rs, rc = assert_parity(g, r, what="tess-bvh")
    assert rp["hits"] > 100000
    print("tess-bvh", rp, rs)


def test_instanced_batched_scene(rt, ctx, oracle):
    """cfg4 in miniature: 64 distinct BLASes built in ONE batched launch set, 64 rotated instances."""
    scene = scenes.instanced_scene(n_side=8, quads=20, width=960, height=540, bounces=1)
    g, r, gs = _run(rt, ctx, oracle, scene, oracle.MODE_BVH, stats=True)
    rp, rs, rc = assert_parity(g, r, what="inst-batched")
    # batched build == one-by-one builds
    sh = rt.SceneHandles(ctx, scene, batch=False)
    g2 = sh.trace(want_hits=True)
    sh.free()
    assert g[1].tobytes() == g2[1].tobytes() and np.array_equal(g[0], g2[0])
    assert gs["rays_secondary"] == rp["hits"]
    print("inst-batched", rp, rs, gs)


@pytest.mark.parametrize("flags", [0, 0x200], ids=["packed-sort", "pair-sort"])
@pytest.mark.parametrize("kind", ["tess", "dupkeys", "tiny", "seg", "dupseg"])
def test_lbvh_build_matches_oracle_bit_for_bit(rt, ctx, oracle, flags, kind):
    """Integer/byte work must be bit-exact: sorted Morton keys, primitive order, tree topology and
    every node box of the GPU LBVH equal the CPU restatement's (both sort record formats). The CPU side finds the
    radix tree top-down (Karras), the GPU bottom-up in shared-memory tiles: "dupkeys" exercises the index-augmented
    prefix rule on long runs of equal keys, "tiny" a tree smaller than one tile. "seg" / "dupseg" / "tiny" fit one CTA's
    shared memory and take the segmented sort (dupseg: exactly its capacity), the others the global onesweep sort."""
    if kind == "tess":
        scene = scenes.tess_scene(nx=120, ny=70, width=64, height=64, bounces=0)
    elif kind == "dupkeys":
        scene = scenes.duplicate_key_scene(20_000)
    elif kind == "seg":
        scene = scenes.tess_scene(nx=80, ny=70, width=64, height=64, bounces=0)
    elif kind == "dupseg":
        scene = scenes.duplicate_key_scene(11264)
    else:
        scene = scenes.tess_scene(nx=9, ny=7, width=64, height=64, bounces=0)
    blas = ctx.build_blas(scene.blases[0], flags=flags)
    keys, prims = ctx.last_sorted_keys()
    info = blas.info()
    nodes, tris = blas.export()
    blas.free()
    o = oracle.OracleScene(scene)
    oinfo, onodes, otris, okeys, oprims = o.blas_export(0)
    assert info.triangle_count == oinfo.triangle_count == scene.triangle_count
    if kind in ("dupkeys", "dupseg"):
        assert np.count_nonzero(keys[1:] == keys[:-1]) > info.triangle_count // 2
    if kind in ("seg", "tiny") and flags == 0:
        # the segmented and the global sort are interchangeable, bit for bit
        for other in (0x400, 0x800):       # global sort; segmented sort with separate setup / Morton kernels
            blas2 = ctx.build_blas(scene.blases[0], flags=other)
            keys2, prims2 = ctx.last_sorted_keys()
            nodes2, tris2 = blas2.export()
            blas2.free()
            assert np.array_equal(keys, keys2) and np.array_equal(prims, prims2) and np.array_equal(nodes, nodes2) and np.array_equal(tris, tris2), hex(other)
    assert np.array_equal(keys, okeys), "sorted Morton keys differ"
    assert np.array_equal(prims, oprims), "sorted primitive order differs (sort not stable?)"
    assert np.array_equal(tris[:, :11], otris[:, :11]), "sorted triangle records differ"
    assert info.root_ref == oinfo.root_ref and info.max_depth == oinfo.max_depth
    assert list(info.bounds_lo) == list(oinfo.bounds_lo) and list(info.bounds_hi) == list(oinfo.bounds_hi)
    n = walk_compare_bvh(nodes, info.root_ref, onodes, oinfo.root_ref)
    assert n >= info.triangle_count // 8
    print("lbvh nodes compared:", n, "depth", info.max_depth)


def test_build_invariants_large(rt, ctx):
    """Size-independent properties at 2M triangles (no oracle needed): keys sorted, values a permutation,
    every triangle in exactly one leaf, parent boxes contain children."""
    scene = scenes.tess_scene(nx=1000, ny=1000, width=64, height=64, bounces=0)
    blas = ctx.build_blas(scene.blases[0])
    t = ctx.build_timing()
    keys, prims = ctx.last_sorted_keys()
    info = blas.info()
    nodes, tris = blas.export()
    blas.free()
    n = info.triangle_count
    assert n == 2_000_000
    assert np.all(keys[1:] >= keys[:-1])
    assert np.array_equal(np.sort(prims), np.arange(n, dtype=np.uint32))
    # walk the tree: vectorised level-by-level
    f = nodes.view(np.float32)
    covered = np.zeros(n, dtype=np.int32)
    frontier = np.array([info.root_ref], dtype=np.int64)
    box_lo = np.array([info.bounds_lo], dtype=np.float32)
    box_hi = np.array([info.bounds_hi], dtype=np.float32)
    depth = 0
    while frontier.size:
        nd = nodes[frontier]
        fl = f[frontier]
        nxt, nlo, nhi = [], [], []
        for half in (0, 1):
            lo = fl[:, 8 * half:8 * half + 3]
            hi = fl[:, 8 * half + 3:8 * half + 6]
            assert np.all(lo >= box_lo) and np.all(hi <= box_hi), "child box escapes parent box"
            ref = nd[:, 8 * half + 6].view(np.int32).astype(np.int64)
            leaf = ref < 0
            x = (~ref[leaf]).astype(np.int64)
            first, cnt = x >> 3, (x & 7) + 1
            for k in range(4):
                sel = cnt > k
                np.add.at(covered, first[sel] + k, 1)
            tl = f.dtype  # noqa
            nxt.append(ref[~leaf]); nlo.append(lo[~leaf]); nhi.append(hi[~leaf])
        frontier = np.concatenate(nxt); box_lo = np.concatenate(nlo); box_hi = np.concatenate(nhi)
        depth += 1
        assert depth < 200
    assert np.all(covered == 1), "a triangle is not in exactly one leaf"
    assert depth == info.max_depth
    print("build 2M tris:", t)


def test_edge_cases(rt, ctx, oracle):
    """Empty geometry, empty TLAS, masked instance, singular instance transform, duplicate coincident
    triangles (equal-t tie-break), degenerate triangles, ray parameters."""
    S = scenes
    quad_v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    quad_i = np.array([[0, 1, 3], [1, 2, 3]], dtype=np.uint32)
    empty = S.Geometry(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint32), None)
    degenerate = S.Geometry(np.array([[0, 0, 1], [0, 0, 1], [0, 0, 1], [0, 0, 1], [1, 1, 1], [2, 2, 1]], dtype=np.float32), None, None)
    dup = [S.Geometry(quad_v, quad_i, None), S.Geometry(quad_v.copy(), quad_i.copy(), None)]      # coincident geometries
    singular = np.array([1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0], dtype=np.float32)
    inst = [
        S.Instance(S.translation(-2.5, 0, 0), 1, 0xFF, 0, 1, 0),      # duplicate quads: geometry 0 must win the tie
        S.Instance(S.translation(-2.5, 0, 0), 2, 0xFF, 0, 1, 0),      # coincident instance: instance 0 must win
        S.Instance(S.translation(2.5, 0, 0), 3, 0x00, 0, 1, 0),       # masked out
        S.Instance(singular, 4, 0xFF, 0, 1, 0),                       # singular transform: never hit
        S.Instance(S.translation(0, 2.5, 0), 5, 0xFF, 1, 1, 1),       # BLAS with an empty + a degenerate geometry + a quad
        S.Instance(S.translation(0, -2.5, 0), 6, 0xFF, 0, 1, 2),      # completely empty BLAS
    ]
    scene = S.Scene("edge", [dup, [empty, degenerate, S.Geometry(quad_v, quad_i, None)], [empty]], inst,
                    S.SAMPLE_HIT_RECORDS.copy(), width=400, height=300, bounces=1)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE)
    rp, rs, rc = assert_parity(g, r, what="edge")
    prim = g[1]
    hit = prim["instance_id"] != MISS
    assert set(np.unique(prim["instance_id"][hit])) == {0, 4}
    m0 = prim["instance_id"] == 0
    assert np.all(prim["geometry_index"][m0] == 0)
    assert np.all(prim["geometry_index"][prim["instance_id"] == 4] == 2)
    # empty TLAS: every ray misses
    tl = ctx.build_tlas([], [])
    ctx.set_hit_records(scene.hit_records)
    rgba, p, _ = ctx.trace(tl, ctx.camera((0, 0, 10), 60), 64, 48, 0, want_hits=True)
    tl.free()
    assert np.all(p["instance_id"] == MISS) and np.all(rgba == np.array((0, 0, 51, 0), dtype=np.uint8))
    # ray parameters: tmax just short of the plane -> all miss; cull mask 0x00 -> all miss; then restore
    sh = rt.SceneHandles(ctx, S.sample_scene(300, 200))
    ctx.set_ray_params(tmax=9.99)
    _, p, _ = sh.trace(want_hits=True)
    assert np.all(p["instance_id"] == MISS)
    ctx.set_ray_params(tmin=10.5)
    _, p, _ = sh.trace(want_hits=True)
    assert np.all(p["instance_id"] == MISS)
    ctx.set_ray_params(cull_mask=0x00)
    _, p, _ = sh.trace(want_hits=True)
    assert np.all(p["instance_id"] == MISS)
    ctx.set_ray_params()
    _, p, _ = sh.trace(want_hits=True)
    assert (p["instance_id"] != MISS).sum() > 1000
    # releasing the cached scratch / staging buffers changes nothing but the memory footprint
    ctx.release_scratch()
    _, p2, _ = sh.trace(want_hits=True)
    assert p2.tobytes() == p.tobytes()
    # SBT range: fewer records than addressed -> refused, never silently wrong
    ctx.set_hit_records(S.SAMPLE_HIT_RECORDS[:3])
    with pytest.raises(rt.RtError) as e:
        sh.trace()
    assert e.value.code == rt.RT_ERROR_SBT_RANGE
    sh.free()


def test_row_partition_and_unpack(rt, ctx, oracle):
    """Image-space split used for multi-GPU: the packed bands of every part, unpacked, equal the full frame."""
    import torch
    scene = scenes.instanced_scene(n_side=4, quads=12, width=322, height=250, bounces=1)
    sh = rt.SceneHandles(ctx, scene)
    full, _, _ = sh.trace(want_hits=False)
    for parts in (2, 3, 8):
        px = ctx.rows_packed_pixels(scene.width, scene.height, 8, parts)
        packed = torch.zeros((parts, px, 4), dtype=torch.uint8, device="cuda:0")
        for p in range(parts):
            ctx.trace_rows(sh.tlas, sh.cam, scene.width, scene.height, 1, 8, p, parts, packed[p], device=True)
        out = torch.zeros((scene.height, scene.width, 4), dtype=torch.uint8, device="cuda:0")
        ctx.unpack_rows(packed, scene.width, scene.height, 8, parts, out)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), full), f"parts={parts}"
        # RT_TRACE_OUT_FULL_FRAME: every part stores its pixels at their final position of ONE frame (what the ranks do with
        # rank 0's shared framebuffer over NVLink): no packing, no unpack
        direct = torch.zeros((scene.height, scene.width, 4), dtype=torch.uint8, device="cuda:0")
        for p in range(parts):
            ctx.trace_rows(sh.tlas, sh.cam, scene.width, scene.height, 1, 8, p, parts, direct, device=True, full_frame=True)
        torch.cuda.synchronize()
        assert np.array_equal(direct.cpu().numpy(), full), f"full-frame parts={parts}"
    # rt_trace_rows_range: the same frame in three row ranges of two parts (what the chunk-pipelined multi-GPU e2e path does)
    direct = torch.zeros((scene.height, scene.width, 4), dtype=torch.uint8, device="cuda:0")
    for p in range(2):
        for a, b in ((0, 64), (64, 104), (104, 128)):       # 250 rows / 2 parts -> 128 packed rows per part
            ctx.trace_rows_range(sh.tlas, sh.cam, scene.width, scene.height, 1, 8, p, 2, a, b - a, direct, full_frame=True)
    ctx.sync()
    assert np.array_equal(direct.cpu().numpy(), full)
    # a shareable framebuffer (cudaMalloc + IPC handle) works as an output like any device buffer
    ptr, handle = ctx.frame_share_create(scene.width * scene.height * 4)
    assert len(handle) == 64 and any(handle)
    ctx.trace_rows(sh.tlas, sh.cam, scene.width, scene.height, 1, 8, 0, 1, ptr, device=True, full_frame=True)
    view = rt.device_view(ptr, scene.width * scene.height * 4, "cuda:0").view(scene.height, scene.width, 4)
    assert np.array_equal(view.cpu().numpy(), full)
    del view
    ctx.frame_share_free(ptr)
    # stream-ordered flags (the multi-GPU frame handshake): add / wait on a counter in the tail of a shared frame
    ptr, _ = ctx.frame_share_create(scene.width * scene.height * 4 + 256)
    ctr = ptr + scene.width * scene.height * 4
    ctx.flag_wait_ge(ctr, 0)
    ctx.flag_add(ctr)
    ctx.flag_add(ctr)
    ctx.flag_wait_ge(ctr, 2)
    ctx.sync()
    assert int(rt.device_view(ctr, 4, "cuda:0").view(torch.int32).cpu()[0]) == 2
    ctx.flag_wait_ge(ctr, 3)              # nobody will ever add the third: the wait gives up after ~4 s and reports instead of hanging
    with pytest.raises(rt.RtError) as e:
        ctx.sync()
    assert e.value.code == rt.RT_ERROR_INTERNAL
    ctx.sync()                            # the error state is cleared
    ctx.frame_share_free(ptr)
    sh.free()


def test_tlas_update_and_device_inputs(rt, ctx, oracle):
    """Per-frame path: rt_update_tlas with moved instances; device-pointer geometry inputs."""
    import torch
    scene = scenes.sample_scene(300, 200)
    sh = rt.SceneHandles(ctx, scene)
    moved = [scenes.Instance(scenes.translation(0.5, 2, 0), 100, 0xFF, 0, 1, 0), scenes.Instance(scenes.translation(-0.5, -2, 0), 100, 0xFF, 2, 1, 0)]
    ctx.update_tlas(sh.tlas, moved, sh.blases)
    g = sh.trace(want_hits=True)
    sh.free()
    scene2 = scenes.sample_scene(300, 200)
    scene2.instances = moved
    o = oracle.OracleScene(scene2)
    r = o.trace(mode=oracle.MODE_BRUTE)
    assert_parity(g, r, what="tlas-update")
    # device-resident vertex/index buffers (torch tensors) feed the same build
    geo = scenes.heightfield(40, 30, -3, 3, -2, 2, 0.8, 5)
    dv = torch.from_numpy(geo.vertices).cuda()
    di = torch.from_numpy(geo.indices.astype(np.int32)).cuda()
    b_dev = ctx.build_blas([scenes.Geometry(dv, di, None)], device=True)
    n_dev, t_dev = b_dev.export()
    b_host = ctx.build_blas([geo])
    n_host, t_host = b_host.export()
    assert np.array_equal(t_dev, t_host) and b_dev.info().root_ref == b_host.info().root_ref
    walk_compare_bvh(n_dev, b_dev.info().root_ref, n_host, b_host.info().root_ref)
    b_dev.free(); b_host.free()


def test_obj_mesh_build_and_trace(rt, ctx, oracle, tmp_path):
    """a Wavefront .obj (two groups -> two geometries of one BLAS, quads fan-triangulated by the
    loader) goes through rt_obj_load -> rt_build_blas -> rt_trace; the oracle brute-forces the same arrays as parsed by
    the independent Python statement of the grammar (tests/test_io_formats.py)."""
    from test_io_formats import py_parse_obj
    hf = scenes.heightfield(24, 18, -3.0, 3.0, -2.0, 2.0, 0.6, 21)
    v = hf.vertices.astype(np.float64)
    quads = hf.indices.reshape(-1, 2, 3)                     # grid_indices emits two triangles per quad
    lines = [f"v {x!r} {y!r} {z!r}" for x, y, z in v.tolist()]
    half = len(quads) // 2
    for gi, (a, b) in enumerate(((0, half), (half, len(quads)))):
        lines.append(f"g part{gi}")
        for q in quads[a:b]:
            lines.append("f " + " ".join(str(int(i) + 1) for i in q[0]))
            lines.append("f " + " ".join(f"{int(i) + 1}//1" for i in q[1]))
    text = "\n".join(lines) + "\n"
    path = tmp_path / "hf.obj"
    path.write_text(text)
    mesh = rt.ObjMesh(path=str(path))
    pv, pt, pg = py_parse_obj(text)
    assert np.array_equal(mesh.vertices.view(np.uint32), pv.view(np.uint32)) and np.array_equal(mesh.indices, pt) and len(pg) == 2
    geoms = mesh.geometries()
    assert sum(g.triangle_count for g in geoms) == hf.triangle_count
    inst = scenes.Instance(scenes.rotation_3x4(np.array([1.0, 0.3, 0.0]), 0.4, np.array([0.0, 0.0, 0.0])), 7, 0xFF, 0,
                           scenes.INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0)
    scene = scenes.Scene("obj", [geoms], [inst], scenes.SAMPLE_HIT_RECORDS[:2].copy(), width=320, height=200, bounces=1)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE)
    rp, rs, rc = assert_parity(g, r, what="obj")
    assert rp["hits"] > 5000 and set(np.unique(g[1]["geometry_index"][g[1]["instance_id"] != MISS])) == {0, 1}
    print("obj", rp, rs, rc)


def _flag_scene(seed=9):
    """Fuzz scene with every flag that feeds the GENERAL trace variant: non-opaque geometries, instances with
    FLIP_FACING / FORCE_OPAQUE / FORCE_NO_OPAQUE, some with facing culls disabled."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=500, n_instances=10, seed=seed, width=320, height=200, bounces=1,
                                shared_edges=True)
    for b, geoms in enumerate(scene.blases):
        for g, geo in enumerate(geoms):
            geo.flags = 1 if (b + g) % 2 == 0 else 0
    for i, I in enumerate(scene.instances):
        I.flags = [0x0, 0x1, 0x2, 0x4, 0x8, 0x2 | 0x8, 0x0, 0x4 | 0x2, 0x1 | 0x8, 0x0][i % 10]
        I.mask = 0xFF
    return scene


@pytest.mark.parametrize("ray_flags", [0x10, 0x20, 0x40, 0x80, 0x1 | 0x10, 0x2 | 0x20 | 0x80, 0x10 | 0x40, 0x1 | 0x8, 0x8 | 0x20, 0x0, 0x2],
                         ids=lambda f: f"rayflags{f:#04x}")
def test_ray_flags_vs_brute_force(rt, ctx, oracle, ray_flags):
    """face culls, opacity culls, instance FORCE_*/FLIP flags, SkipClosestHitShader — ids, t, u, v
    bit-exact and RGBA8 within +-1 LSB against the oracle's brute force under the same flags."""
    scene = _flag_scene()
    sh = rt.SceneHandles(ctx, scene)
    try:
        ctx.set_ray_params(ray_flags=ray_flags)
        g = sh.trace(want_hits=True)
        ctx.set_ray_params()
        base = sh.trace(want_hits=True)
    finally:
        ctx.set_ray_params()
        sh.free()
    o = oracle.OracleScene(scene)
    r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=ray_flags))
    o.close()
    rp, rs, rc = assert_parity(g, r, what=f"rayflags{ray_flags:#x}")
    changed = int((g[1]["primitive_id"] != base[1]["primitive_id"]).sum())
    if ray_flags & 0xF0:
        assert changed > 200, "the culls must actually change the image of this scene"
    elif not (ray_flags & 0x8):
        assert changed == 0
    print("rayflags", hex(ray_flags), rp, rs, rc, "changed", changed)


def test_terminate_on_first_hit_and_miss_records(rt, ctx, oracle):
    """TerminateOnFirstHit: which hit is undefined, so only the hit/miss mask (primary) is compared; with
    SkipClosestHitShader it is the classic shadow ray. Miss records: rt_ray_params.miss_index selects the colour."""
    scene = _flag_scene(seed=10)
    scene.bounces = 0
    sh = rt.SceneHandles(ctx, scene)
    o = oracle.OracleScene(scene)
    try:
        for rf in (0x1 | 0x4, 0x4 | 0x8 | 0x10, 0x4 | 0x40):
            ctx.set_ray_params(ray_flags=rf)
            rgba, prim, _ = sh.trace(want_hits=True)
            r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=rf))
            mg, mr = prim["instance_id"] != MISS, r[1]["instance_id"] != MISS
            assert np.array_equal(mg, mr) and mg.sum() > 1000
            if rf & 0x8:
                assert np.array_equal(rgba, r[0])                      # hit -> (0,0,0,0), miss -> miss colour: fully defined
        miss = np.array([[0, 0, 0.2], [0.25, 0.5, 1.0]], dtype=np.float32)
        ctx.set_miss_records(miss)
        o.set_miss_records(miss)
        ctx.set_ray_params(miss_index=1)
        g = sh.trace(want_hits=True)
        r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(miss_index=1))
        assert_parity(g, r, what="miss1")
        assert tuple(g[0][prim["instance_id"] == MISS][0]) == (64, 128, 255, 0)
        ctx.set_ray_params(miss_index=2)
        with pytest.raises(rt.RtError) as e:
            sh.trace()
        assert e.value.code == rt.RT_ERROR_SBT_RANGE
    finally:
        ctx.set_ray_params()
        ctx.set_miss_records(np.array([[0, 0, 0.2]], dtype=np.float32))
        o.close()
        sh.free()


def test_blas_update_animated_mesh(rt, ctx, oracle):
    """the per-frame loop of an animated mesh. rt_update_blas re-builds the BLAS in place (same handle,
    same device address), rt_update_tlas re-reads it; every frame equals a from-scratch build of that frame's geometry."""
    S = scenes
    frames = [S.heightfield(30, 20, -3.0, 3.0, -2.0, 2.0, 0.3 + 0.25 * f, 40 + f) for f in range(3)]
    inst = [S.Instance(S.rotation_3x4(np.array([0.2, 1.0, 0.0]), 0.3, np.array([0.0, 0.0, 0.0])), 3, 0xFF, 0, 1, 0),
            S.Instance(S.translation(0.0, 0.0, -3.0), 4, 0xFF, 1, 1, 0)]
    scene = S.Scene("anim", [[frames[0]]], inst, S.SAMPLE_HIT_RECORDS[:2].copy(), width=320, height=200, bounces=1)
    sh = rt.SceneHandles(ctx, scene)
    handle0, info0 = sh.blases[0].handle, sh.blases[0].info()
    try:
        for f in (1, 2):
            ctx.update_blas(sh.blases[0], [frames[f]])
            assert sh.blases[0].handle == handle0
            assert sh.blases[0].info().device_storage == info0.device_storage
            ctx.update_tlas(sh.tlas, inst, sh.blases)
            g = sh.trace(want_hits=True)
            sc = S.Scene("anim", [[frames[f]]], inst, scene.hit_records, width=320, height=200, bounces=1)
            o = oracle.OracleScene(sc)
            r = o.trace(mode=oracle.MODE_BRUTE)
            o.close()
            rp, rs, rc = assert_parity(g, r, what=f"anim{f}")
            assert rp["hits"] > 5000
        # counts must match the original build
        with pytest.raises(rt.RtError):
            ctx.update_blas(sh.blases[0], [S.heightfield(10, 10, -1, 1, -1, 1, 0.1, 1)])
    finally:
        sh.free()


def test_cpp_host_program(rt, ctx, oracle, tmp_path):
    """The C++ host program above the C ABI (host/sample_scene.cpp, the headless mirror of the reference's main()):
    its PPM equals the oracle's frame of the sample scene; with --obj it builds from a Wavefront file."""
    import subprocess
    from build_up_phase_b200 import build as b
    exe = b.build_host_sample()
    out = tmp_path / "frame.ppm"
    p = subprocess.run([exe, str(out), "600", "400"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    raw = out.read_bytes()
    hdr = b"P6\n600 400\n255\n"
    assert raw.startswith(hdr)
    img = np.frombuffer(raw[len(hdr):], dtype=np.uint8).reshape(400, 600, 3)
    o = oracle.OracleScene(scenes.sample_scene(600, 400))
    ref = o.trace(mode=oracle.MODE_BRUTE)[0]
    o.close()
    assert np.abs(img.astype(np.int16) - ref[:, :, :3].astype(np.int16)).max() <= 1
    # --obj: a unit quad made of two groups fills the middle of the frame with the first two hit-record colours
    objf = tmp_path / "quad.obj"
    objf.write_text("v -1 -1 0\nv 1 -1 0\nv 1 1 0\nv -1 1 0\ng a\nf 1 2 4\ng b\nf 2 3 4\n")
    p = subprocess.run([exe, str(out), "300", "200", "--obj", str(objf), "--srgb"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr
    assert "4 vertices, 2 triangles, 2 group(s)" in p.stdout
    raw = out.read_bytes()
    img = np.frombuffer(raw[len(b"P6\n300 200\n255\n"):], dtype=np.uint8).reshape(200, 300, 3)
    lut = rt.srgb8_table()
    cols = {tuple(c) for c in img.reshape(-1, 3).tolist()}
    assert cols == {tuple(lut[[0, 0, 51]].tolist()), tuple(lut[[153, 26, 51]].tolist()), tuple(lut[[26, 204, 102]].tolist())}


def test_batched_build_paths_agree(rt, ctx, oracle):
    """A batch whose largest BLAS exceeds one CTA's shared memory takes the global onesweep sort with the BLAS id as key
    prefix; a batch of small BLASes takes the fused per-BLAS kernel. Both must equal the one-by-one builds bit for bit
    (nodes, sorted triangles) and trace like the oracle."""
    S = scenes
    big = S.heightfield(90, 80, -2.0, 2.0, -2.0, 2.0, 0.5, 3)            # 14,400 triangles > SEG_SORT_CAPACITY
    small = [S.heightfield(20 + 3 * k, 17, -1.0, 1.0, -1.0, 1.0, 0.4, 10 + k) for k in range(3)]
    for blases in ([[big], [small[0]], [small[1], small[2]]], [[g] for g in small]):
        batch = ctx.build_blas_batch(blases)
        single = [ctx.build_blas(b) for b in blases]
        for hb, hs in zip(batch, single):
            nb, tb = hb.export()
            ns, ts = hs.export()
            ib, is_ = hb.info(), hs.info()
            assert ib.root_ref == is_.root_ref and ib.max_depth == is_.max_depth and ib.triangle_count == is_.triangle_count
            assert np.array_equal(tb, ts)
            walk_compare_bvh(nb, ib.root_ref, ns, is_.root_ref)
        for h in batch + single:
            h.free()
    inst = [S.Instance(S.translation(-2.5 + 2.5 * k, 0.0, 0.0), k, 0xFF, 0, 1, k) for k in range(3)]
    scene = S.Scene("mixed-batch", [[big], [small[0]], [small[1], small[2]]], inst, S.SAMPLE_HIT_RECORDS[:2].copy(), width=320, height=200, bounces=1)
    g, r, _ = _run(rt, ctx, oracle, scene, oracle.MODE_BRUTE, batch=True)
    rp, rs, rc = assert_parity(g, r, what="mixed-batch")
    assert rp["hits"] > 5000


# ---- round 2: the surfaces VERDICT r01 listed as implemented but untested ----------------------------------------------
def _sbt_scene(seed=21, n_geoms=3):
    """Several geometries per BLAS and distinct instance SBT offsets, so that sbtRecordStride / sbtRecordOffset
    (main.cpp:1260-1262) select visibly different hit records."""
    scene = scenes.random_scene(n_blas=2, tris_per_blas=360, n_instances=6, seed=seed, width=320, height=200, bounces=1,
                                n_geoms=n_geoms, shared_edges=True)
    for i, I in enumerate(scene.instances):
        I.mask = 0xFF
        I.sbt_offset = [0, 3, 1, 5, 2, 4][i % 6]
    rng = np.random.default_rng(seed)
    scene.hit_records = rng.uniform(0.05, 0.95, size=(32, 3)).astype(np.float32)
    return scene


@pytest.mark.parametrize("stride,offset", [(2, 0), (3, 0), (1, 5), (2, 3), (3, 7), (0, 2)])
def test_sbt_stride_and_offset_vs_brute_force(rt, ctx, oracle, stride, offset):
    """record = instanceSbtOffset + geometryIndex * sbtRecordStride + sbtRecordOffset (main.cpp:1260-1262) for strides and
    offsets other than the sample's (1, 0): ids/t/u/v bit-exact, RGBA8 within +-1 LSB of the oracle's brute force, and the
    image must really differ from the default-parameter image (the records are all distinct)."""
    scene = _sbt_scene()
    sh = rt.SceneHandles(ctx, scene)
    try:
        base = sh.trace(want_hits=True)
        ctx.set_ray_params(sbt_record_stride=stride, sbt_record_offset=offset)
        g = sh.trace(want_hits=True)
    finally:
        ctx.set_ray_params()
        sh.free()
    o = oracle.OracleScene(scene)
    r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(sbt_record_stride=stride, sbt_record_offset=offset))
    o.close()
    rp, rs, rc = assert_parity(g, r, what=f"sbt stride {stride} offset {offset}")
    assert rp["hits"] > 3000 and rs["hits"] > 50
    assert g[1].tobytes() == base[1].tobytes()                       # the SBT rule selects shading, never geometry
    hit = g[1]["instance_id"] != MISS
    assert (np.abs(g[0].astype(np.int16) - base[0].astype(np.int16)).max(axis=-1)[hit] > 0).mean() > 0.5
    # spot-check the rule itself on the GPU output: pixels whose secondary ray missed show 0.5 * record + 0.5 * miss
    sec_miss = hit & (g[2]["instance_id"] == MISS)
    inst_sbt = np.array([I.sbt_offset for I in scene.instances], dtype=np.int64)
    rec = inst_sbt[g[1]["instance_id"][sec_miss]] + g[1]["geometry_index"][sec_miss].astype(np.int64) * stride + offset
    special = (g[1]["primitive_id"][sec_miss] == 1) & (g[1]["instance_id"][sec_miss] == 1) & (g[1]["custom_index"][sec_miss] == 100) & \
              (g[1]["geometry_index"][sec_miss] == 1)
    want = 0.5 * scene.hit_records[rec] + 0.5 * scene.miss_color[None, :]
    got = g[0][sec_miss][:, :3].astype(np.float64) / 255.0
    assert np.abs(got - want)[~special].max() < 1.01 / 255.0


def test_sbt_range_check_is_exact(rt, ctx):
    """the static range check must reduce max_i(sbt_i + (n_geoms_i - 1) * stride) per instance, not combine maxima
    of different instances: instance A (sbt 0, 3 geometries) and B (sbt 5, 1 geometry) with stride 2 address records 0..5."""
    S = scenes
    quad_v = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)
    quad_i = np.array([[0, 1, 3], [1, 2, 3]], dtype=np.uint32)
    three = [S.Geometry(quad_v, quad_i, S.translation(-3 + 3 * k, 1.5, 0)) for k in range(3)]
    one = [S.Geometry(quad_v, quad_i, S.translation(0, -1.5, 0))]
    inst = [S.Instance(S.IDENTITY_3X4.copy(), 1, 0xFF, 0, 1, 0), S.Instance(S.IDENTITY_3X4.copy(), 2, 0xFF, 5, 1, 1)]
    rec = np.linspace(0.1, 0.9, 18, dtype=np.float32).reshape(6, 3)
    scene = S.Scene("sbt-exact", [three, one], inst, rec, width=200, height=120, bounces=0)
    sh = rt.SceneHandles(ctx, scene)
    try:
        ctx.set_ray_params(sbt_record_stride=2)
        rgba, prim, _ = sh.trace(want_hits=True)                    # six records are enough (the r01 bound said nine)
        hit = prim["instance_id"] != MISS
        got = {(int(i), int(g)): tuple(c) for i, g, c in zip(prim["instance_id"][hit], prim["geometry_index"][hit], rgba[hit][:, :3].tolist())}
        for (i, g), c in got.items():
            r = [0, 5][i] + 2 * g
            assert np.abs(np.array(c) / 255.0 - rec[r]).max() < 1.01 / 255.0
        assert set(got) == {(0, 0), (0, 1), (0, 2), (1, 0)}
        ctx.set_hit_records(rec[:5])
        with pytest.raises(rt.RtError) as e:
            sh.trace()
        assert e.value.code == rt.RT_ERROR_SBT_RANGE
        ctx.set_ray_params(sbt_record_stride=2, ray_flags=0x1 | 0x8)   # SkipClosestHitShader: no hit record is read, nothing to refuse
        rgba2, _, _ = sh.trace(want_hits=True)
        assert np.all(rgba2[hit][:, :3] == 0)
        ctx.set_ray_params(sbt_record_stride=3)                       # 0 + 2*3 = 6 > 5 records even with all six set
        ctx.set_hit_records(rec)
        with pytest.raises(rt.RtError):
            sh.trace()
    finally:
        ctx.set_ray_params()
        sh.free()


@pytest.mark.parametrize("kind", ["sample", "tess1m"])
def test_build_sizes_bound_real_allocations(rt, ctx, kind):
    """vkGetAccelerationStructureBuildSizesKHR (main.cpp:756-762, 892-898): the reported sizes bound what the builds really
    allocate — for the sample's {2, 2} triangle counts / 2 instances and for a 1 M-triangle mesh."""
    if kind == "sample":
        scene = scenes.sample_scene(64, 64)
    else:
        scene = scenes.tess_scene(nx=1000, ny=500, width=64, height=64, bounces=0)
    geoms = scene.blases[0]
    sizes = ctx.blas_build_sizes([g.triangle_count for g in geoms])
    blas = ctx.build_blas(geoms)
    info = blas.info()
    assert info.storage_bytes > 0 and sizes.acceleration_structure_size >= info.storage_bytes
    assert sizes.build_scratch_size >= ctx.build_scratch_bytes() > 0
    assert sizes.acceleration_structure_size < 2 * info.storage_bytes + 4096      # a bound, not a wild guess
    assert sizes.build_scratch_size < 2 * ctx.build_scratch_bytes() + 65536
    tsizes = ctx.tlas_build_sizes(len(scene.instances))
    tlas = ctx.build_tlas(scene.instances, [blas])
    assert tsizes.acceleration_structure_size >= tlas.storage_bytes() > 0
    assert tsizes.build_scratch_size >= ctx.build_scratch_bytes() > 0
    tlas.free(); blas.free()
    t1k = ctx.tlas_build_sizes(1024)
    assert t1k.acceleration_structure_size >= 1024 * (64 + 96)


def test_blas_import_roundtrip_traces_identically(rt, ctx, oracle):
    """rt_blas_get_info().device_storage -> (copy, as an NCCL broadcast would deliver it) -> rt_blas_import on ANOTHER context:
    the adopted BLAS has the same root/bounds/depth, exports the same nodes and triangles and traces to the same frame and hits
    (the cfg5 per-GPU build + broadcast path)."""
    import torch
    S = scenes
    geoms = [S.heightfield(50, 40, -2.5, 2.5, -2.0, 2.0, 0.7, 31), S.heightfield(20, 20, -1.0, 1.0, -1.0, 1.0, 0.3, 32)]
    geoms[1].transform = S.translation(0.0, 0.0, 1.5)
    inst = [S.Instance(S.rotation_3x4(np.array([0.3, 1.0, 0.2]), 0.5, np.array([0.5, 0.0, 0.0])), 9, 0xFF, 1, 1, 0),
            S.Instance(S.translation(-1.0, 0.5, -2.0), 10, 0xFF, 0, 1, 0)]
    scene = S.Scene("import", [geoms], inst, S.SAMPLE_HIT_RECORDS.copy(), width=320, height=200, bounces=1)
    sh = rt.SceneHandles(ctx, scene)
    g1 = sh.trace(want_hits=True)
    src = sh.blases[0]
    info = src.info()
    blob = rt.device_view(info.device_storage, int(info.storage_bytes), "cuda:0").clone()      # the broadcast's receive buffer
    torch.cuda.synchronize()
    n1, t1 = src.export()
    with rt.Context(0) as ctx2:
        b2 = ctx2.import_blas(info, blob)
        del blob
        i2 = b2.info()
        assert (i2.root_ref, i2.max_depth, i2.triangle_count) == (info.root_ref, info.max_depth, info.triangle_count)
        assert list(i2.bounds_lo) == list(info.bounds_lo) and list(i2.bounds_hi) == list(info.bounds_hi)
        assert i2.device_storage != info.device_storage
        n2, t2 = b2.export()
        assert np.array_equal(n1, n2) and np.array_equal(t1, t2)
        tl2 = ctx2.build_tlas(scene.instances, [b2])
        ctx2.set_hit_records(scene.hit_records); ctx2.set_miss_color(scene.miss_color)
        g2 = ctx2.trace(tl2, ctx2.camera(scene.camera_pos, scene.yfov_deg), scene.width, scene.height, 1, want_hits=True)
        tl2.free(); b2.free()
    sh.free()
    assert np.array_equal(g1[0], g2[0]) and g1[1].tobytes() == g2[1].tobytes() and g1[2].tobytes() == g2[2].tobytes()
    o = oracle.OracleScene(scene)
    r = o.trace(mode=oracle.MODE_BRUTE)
    o.close()
    rp, rs, _ = assert_parity(g2, r, what="imported blas")
    assert rp["hits"] > 5000


def test_tlas_from_device_resident_instances(rt, ctx, oracle):
    """RT_BUILD_INSTANCES_ON_DEVICE: the 64-byte instance records live in device memory, like the reference's instance buffer
    (main.cpp:860-868), and carry rt_blas_device_reference() values (its vk.blasAddress, main.cpp:785,853)."""
    import ctypes as C
    import torch
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=7, seed=5, width=320, height=200, bounces=1)
    sh = rt.SceneHandles(ctx, scene)
    g_host = sh.trace(want_hits=True)
    arr = ctx.instance_array(scene.instances, sh.blases)
    for i, I in enumerate(scene.instances):
        ref = sh.blases[I.blas].device_reference()
        assert ref != 0 and ref != sh.blases[I.blas].handle
        arr[i].blas = ref
    raw = np.frombuffer(C.string_at(C.addressof(arr), 64 * len(scene.instances)), dtype=np.uint8).copy()
    dev = torch.from_numpy(raw).cuda()
    tl = ctx.build_tlas_device(dev, len(scene.instances))
    g_dev = ctx.trace(tl, sh.cam, scene.width, scene.height, 1, want_hits=True)
    tl.free(); sh.free()
    assert np.array_equal(g_host[0], g_dev[0]) and g_host[1].tobytes() == g_dev[1].tobytes() and g_host[2].tobytes() == g_dev[2].tobytes()
    o = oracle.OracleScene(scene)
    r = o.trace(mode=oracle.MODE_BRUTE)
    o.close()
    assert_parity(g_dev, r, what="device instances")


def test_rows_range_needs_whole_bands(rt, ctx):
    """with block_rows = 16 a chunk of 8 packed rows is half a band; the header's promise (packed rows [a, b) of all
    parts = image rows [a * parts, b * parts)) holds for whole bands only, so the call is refused."""
    import torch
    scene = scenes.sample_scene(160, 128)
    sh = rt.SceneHandles(ctx, scene)
    out = torch.zeros((128, 160, 4), dtype=torch.uint8, device="cuda:0")
    try:
        with pytest.raises(rt.RtError):
            ctx.trace_rows_range(sh.tlas, sh.cam, 160, 128, 0, 16, 0, 2, 8, 16, out, full_frame=True)
        with pytest.raises(rt.RtError):
            ctx.trace_rows_range(sh.tlas, sh.cam, 160, 128, 0, 16, 0, 2, 0, 24, out, full_frame=True)
        full, _, _ = sh.trace(want_hits=False)
        for p in range(2):
            for a, b in ((0, 32), (32, 64)):
                ctx.trace_rows_range(sh.tlas, sh.cam, 160, 128, 0, 16, p, 2, a, b - a, out, full_frame=True)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy(), full)
    finally:
        sh.free()


def _anyhit_records(n, seed=5):
    """One any-hit record per hit group: ACCEPT, or an alpha mask of 2^k x 2^k cells (k = 0..5) with ~55 % of the bits set."""
    rng = np.random.default_rng(seed)
    recs = []
    for i in range(n):
        if i % 4 == 3:
            recs.append((0, 0, 0, None))                                   # RT_ANYHIT_ACCEPT: no any-hit shader in this hit group
            continue
        k = [3, 5, 0, 2, 4, 1][i % 6]
        bits = rng.random((1 << k) * (1 << k)) < 0.55
        if k == 0:
            bits[:] = (i % 2 == 0)                                         # a 1x1 mask: everything ignored / everything accepted
        words = np.zeros(((1 << (2 * k)) + 31) // 32, dtype=np.uint32)
        for b in np.nonzero(bits)[0]:
            words[b >> 5] |= np.uint32(1 << (int(b) & 31))
        recs.append((1, k, 0, words))
    return recs


@pytest.mark.gpu
@pytest.mark.parametrize("ray_flags", [0x0, 0x2, 0x10, 0x80, 0x1], ids=lambda f: f"rayflags{f:#04x}")
def test_anyhit_alpha_masks_vs_brute_force(rt, ctx, oracle, ray_flags):
    """SURVEY 8(f) row 2, the any-hit stage (shader_module.h:90): alpha-mask records on the non-opaque geometries of the flag
    scene. ids, t, u, v bit-exact and RGBA8 within +-1 LSB against the oracle's brute force, primary and bounce rays."""
    scene = _flag_scene(seed=12)
    recs = _anyhit_records(len(scene.hit_records))
    sh = rt.SceneHandles(ctx, scene)
    o = oracle.OracleScene(scene)
    try:
        ctx.set_ray_params(ray_flags=ray_flags)
        base = sh.trace(want_hits=True)
        ctx.set_anyhit_records(recs)
        g = sh.trace(want_hits=True)
        o.set_anyhit_records(recs)
        r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=ray_flags))
    finally:
        ctx.set_anyhit_records([])
        ctx.set_ray_params()
        o.close()
        sh.free()
    rp, rs, rc = assert_parity(g, r, what=f"anyhit{ray_flags:#x}")
    changed = int((g[1]["primitive_id"] != base[1]["primitive_id"]).sum())
    if ray_flags & 0x1 or ray_flags & 0x80:
        assert changed == 0, "Opaque ray flag / CullNoOpaque: no candidate reaches the any-hit stage"
    else:
        assert changed > 300, "the masks must actually cut holes into this scene"
    print("anyhit", hex(ray_flags), rp, rs, rc, "changed", changed)


@pytest.mark.gpu
def test_anyhit_terminate_ray(rt, ctx, oracle):
    """terminateRayEXT from an any-hit record: WHICH hit is undefined, the hit/miss mask is not."""
    scene = _flag_scene(seed=13)
    scene.bounces = 0
    recs = [(k, l, 1, m) for k, l, _, m in _anyhit_records(len(scene.hit_records), seed=6)]
    sh = rt.SceneHandles(ctx, scene)
    o = oracle.OracleScene(scene)
    try:
        ctx.set_ray_params(ray_flags=0x2)                                   # NoOpaque: every candidate runs its any-hit record
        ctx.set_anyhit_records(recs)
        o.set_anyhit_records(recs)
        _, prim, _ = sh.trace(want_hits=True)
        r = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=0x2))
        mg, mr = prim["instance_id"] != MISS, r[1]["instance_id"] != MISS
        assert np.array_equal(mg, mr) and mg.sum() > 1000
        with pytest.raises(rt.RtError):
            ctx.set_anyhit_records([(1, 11, 0, np.zeros(4, dtype=np.uint32))])     # log2_res > 10
    finally:
        ctx.set_anyhit_records([])
        ctx.set_ray_params()
        o.close()
        sh.free()



def _animated(f):
    """Frame f of a moving height field: same grid, the bumps drift (what a refit is for) - not a new random surface."""
    S = scenes
    g = S.heightfield(60, 40, -3.0, 3.0, -2.0, 2.0, 0.5, 40)
    v = g.vertices.copy()
    v[:, 2] = (v[:, 2] * np.float32(1.0 + 0.15 * f) + np.float32(0.05 * f) * np.sin(v[:, 0] * np.float32(2.0) + np.float32(f))).astype(np.float32)
    v[:, 0] = (v[:, 0] + np.float32(0.02 * f) * np.cos(v[:, 1] * np.float32(3.0))).astype(np.float32)
    return S.Geometry(np.ascontiguousarray(v), g.indices, None)


def test_blas_refit_update_keeps_topology(rt, ctx, oracle):
    """SURVEY 8(f) row 3, VK_BUILD_ACCELERATION_STRUCTURE_MODE_UPDATE_KHR proper: RT_BUILD_MODE_REFIT keeps the sorted order / tree of the
    last full build (RT_BUILD_ALLOW_UPDATE) and re-fits every box. Bit-for-bit equal to the oracle's refit (same nodes, same sorted
    triangles), traces like a brute force over the new geometry, and differs from a full rebuild of the same frame."""
    S = scenes
    inst = [S.Instance(S.rotation_3x4(np.array([0.2, 1.0, 0.0]), 0.3, np.array([0.0, 0.0, 0.0])), 3, 0xFF, 0, 1, 0)]
    scene = S.Scene("refit", [[_animated(0)]], inst, S.SAMPLE_HIT_RECORDS[:1].copy(), width=320, height=200, bounces=1)
    sh = rt.SceneHandles(ctx, scene, build_flags=rt.RT_BUILD_ALLOW_UPDATE)
    o = oracle.OracleScene(scene)
    try:
        n0, t0 = sh.blases[0].export()
        prim_order0 = t0[:, 10].copy()
        for f in (1, 2):
            geo = _animated(f)
            ctx.update_blas(sh.blases[0], [geo], flags=rt.RT_BUILD_ALLOW_UPDATE | rt.RT_BUILD_MODE_REFIT)
            timing = ctx.build_timing()
            assert timing["sort_ms"] < 0.01 and timing["morton_ms"] < 0.01, "a refit runs no Morton pass and no sort"
            ctx.update_tlas(sh.tlas, inst, sh.blases)
            n1, t1 = sh.blases[0].export()
            assert np.array_equal(t1[:, 10], prim_order0), "a refit keeps the sorted triangle order"
            assert not np.array_equal(t1[:, :9], t0[:, :9])
            o.refit_blas(0, [geo])
            _, on, ot, _, _ = o.blas_export(0)
            info, oinfo = sh.blases[0].info(), o.blas_info(0)
            assert info.root_ref == oinfo.root_ref and info.max_depth == oinfo.max_depth
            assert walk_compare_bvh(n1, info.root_ref, on, oinfo.root_ref) > 1000
            assert np.array_equal(t1[:, :11], ot[:, :11])
            g = sh.trace(want_hits=True)
            sc = S.Scene("refit", [[geo]], inst, scene.hit_records, width=320, height=200, bounces=1)
            ob = oracle.OracleScene(sc)
            r = ob.trace(mode=oracle.MODE_BRUTE)
            ob.close()
            rp, rs, rc = assert_parity(g, r, what=f"refit{f}")
            assert rp["hits"] > 5000
        # the same frame, fully rebuilt: a different (re-sorted) tree, the same image
        ctx.update_blas(sh.blases[0], [_animated(2)], flags=rt.RT_BUILD_ALLOW_UPDATE)
        ctx.update_tlas(sh.tlas, inst, sh.blases)
        n2, t2 = sh.blases[0].export()
        assert not np.array_equal(t2[:, 10], prim_order0)
        g2 = sh.trace(want_hits=True)
        assert np.array_equal(g2[0], g[0]) and g2[1].tobytes() == g[1].tobytes()
        # a BLAS built without ALLOW_UPDATE has nothing to refit from
        plain = ctx.build_blas([_animated(0)])
        with pytest.raises(rt.RtError):
            ctx.update_blas(plain, [_animated(1)], flags=rt.RT_BUILD_MODE_REFIT)
        plain.free()
    finally:
        o.close()
        sh.free()


@pytest.mark.parametrize("kind", ["single", "batch"])
def test_blas_compaction(rt, ctx, oracle, kind):
    """SURVEY 8(f) row 3, VK_COPY_ACCELERATION_STRUCTURE_MODE_COMPACT_KHR: rt_compact_blas packs the live nodes (order-preserving, dense)
    into a right-sized allocation. Same tree (lock-step walk: boxes bit-exact, leaf refs equal), same frame and hit records."""
    if kind == "single":
        scene = scenes.tess_scene(nx=120, ny=80, width=320, height=200, bounces=1)
    else:
        scene = scenes.random_scene(n_blas=5, tris_per_blas=700, n_instances=12, seed=31, width=320, height=200, bounces=1, shared_edges=True)
    sh = rt.SceneHandles(ctx, scene, build_flags=rt.RT_BUILD_ALLOW_COMPACTION)
    try:
        before = sh.trace(want_hits=True)
        exported = [(b.export()[0], b.info().root_ref, b.info().node_count) for b in sh.blases]
        b0, b1 = ctx.compact_blas(sh.blases[0])
        assert b1 < b0
        sh.rebuild_tlas()
        after = sh.trace(want_hits=True)
        assert np.array_equal(before[0], after[0]) and before[1].tobytes() == after[1].tobytes() and before[2].tobytes() == after[2].tobytes()
        live_total = 0
        for b, (n_old, root_old, slots_old) in zip(sh.blases, exported):
            info = b.info()
            n_new = b.export()[0]
            pairs = walk_compare_bvh_renumbered(n_old, root_old, n_new, info.root_ref)
            assert info.node_count == len(pairs) <= slots_old, "exactly the reachable nodes are kept"
            assert sorted(p[1] for p in pairs) == list(range(len(pairs))), "dense numbering"
            by_old = sorted(pairs)
            assert all(by_old[i][1] < by_old[i + 1][1] for i in range(len(by_old) - 1)), "order-preserving: siblings stay adjacent"
            live_total += len(pairs)
            assert info.triangle_count == slots_old
        n_tris = sum(b.info().triangle_count for b in sh.blases)
        assert 0.3 * n_tris < live_total < 0.8 * n_tris
        print("compaction", kind, "bytes", b0, "->", b1, "live nodes", live_total, "of", n_tris, "slots")
        assert ctx.compact_blas(sh.blases[0]) == (b1, b1)                  # idempotent
        if kind == "single":
            with pytest.raises(rt.RtError):
                ctx.update_blas(sh.blases[0], scene.blases[0])             # a compacted BLAS cannot be updated
        plain = ctx.build_blas(scene.blases[0])
        with pytest.raises(rt.RtError):
            ctx.compact_blas(plain)                                         # needs RT_BUILD_ALLOW_COMPACTION
        plain.free()
    finally:
        sh.free()


def _padded_copy(scene, cols, extra_rows):
    """The same scene with every vertex array stored as [nv + extra_rows, cols]: x y z first, NaN in the padding columns (vertexStride =
    4 * cols, main.cpp:733), and unreferenced rows of 1e30 behind the real vertices (maxVertex counts them, no index refers to them)."""
    import copy
    s = copy.deepcopy(scene)
    for geoms in s.blases:
        for g in geoms:
            v = np.asarray(g.vertices, dtype=np.float32)
            pv = np.full((v.shape[0] + extra_rows, cols), np.nan, dtype=np.float32)
            pv[:v.shape[0], :3] = v
            pv[v.shape[0]:, :3] = 1e30
            if g.indices is None and extra_rows:                      # a non-indexed list uses every vertex: no spare rows there
                pv = pv[:v.shape[0]]
            g.vertices = pv
    return s


@pytest.mark.parametrize("n_geoms", [1, 2], ids=["one-geometry-blas", "two-geometry-blas"])
def test_vertex_stride_and_unreferenced_vertices(rt, ctx, oracle, n_geoms):
    """a1 (main.cpp:726-746): vertexStride and maxVertex. A BLAS that is ONE indexed geometry has its vertex array staged in shared memory by
    the fused per-BLAS kernel (stride and all), a BLAS of several geometries gathers from global memory: both must read x y z at the stride and
    nothing else, whatever lies in the padding or behind the last referenced vertex. Third case: a one-geometry BLAS whose vertex array is too
    large for the staging area falls back to the global gather."""
    base = scenes.random_scene(n_blas=4, tris_per_blas=400, n_instances=6, seed=13, width=320, height=200, bounces=1, n_geoms=n_geoms)
    g0, r0, _ = _run(rt, ctx, oracle, base, oracle.MODE_BRUTE)
    assert_parity(g0, r0, what=f"stride base {n_geoms}")
    for cols, extra in ((4, 0), (5, 7), (8, 3)):
        padded = _padded_copy(base, cols, extra)
        g1, r1, _ = _run(rt, ctx, oracle, padded, oracle.MODE_BRUTE)
        assert g1[1].tobytes() == g0[1].tobytes() and g1[2].tobytes() == g0[2].tobytes() and np.array_equal(g1[0], g0[0]), (cols, extra)
        assert r1[1].tobytes() == r0[1].tobytes() and np.array_equal(r1[0], r0[0])            # the oracle reads the stride the same way
    if n_geoms == 1:
        big = _padded_copy(base, 3, 9000)                              # (nv + 9000) * 12 B > the 85 KB staging area: global gather
        assert all((g.vertices.shape[0] * 12) > 87000 for geoms in big.blases for g in geoms)
        g2, _, _ = _run(rt, ctx, oracle, big, oracle.MODE_BRUTE)
        assert g2[1].tobytes() == g0[1].tobytes() and np.array_equal(g2[0], g0[0])
