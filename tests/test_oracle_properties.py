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


def test_ray_flag_semantics_known_answers(oracle):
    """Pins the ray-flag conventions on the reference's own scene. The sample's quad triangles
    (0,1,3) and (1,2,3) (main.cpp:689-695) have the normal (v1-v0)x(v2-v0) = +z, towards the camera at z = 10: they
    appear counter-clockwise from the ray origin, i.e. BACK facing under the default "front = clockwise" rule."""
    S = scenes
    base = S.sample_scene(300, 200)

    def hits(scene, **rp):
        o = oracle.OracleScene(scene)
        rgba, prim, _, st = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(**rp))
        o.close()
        return rgba, prim, st["primary_hits"]

    _, p0, n0 = hits(base)
    assert n0 > 4000
    # the sample sets TRIANGLE_FACING_CULL_DISABLE (main.cpp:852): facing culls change nothing
    assert hits(base, ray_flags=0x1 | 0x10)[2] == n0 and hits(base, ray_flags=0x1 | 0x20)[2] == n0
    # without that flag: back-facing quads vanish under CullBack, stay under CullFront; FLIP_FACING swaps the two
    for flags, back, front in ((0x0, 0, n0), (0x2, n0, 0)):
        sc = S.sample_scene(300, 200)
        for I in sc.instances:
            I.flags = flags
        assert hits(sc, ray_flags=0x1 | 0x10)[2] == back
        assert hits(sc, ray_flags=0x1 | 0x20)[2] == front
    # opacity: geometry OPAQUE (main.cpp:741) -> CullOpaque removes everything, CullNoOpaque nothing; the ray's NoOpaque
    # flag overrides the geometry, the instance FORCE_NO_OPAQUE flag does too, and the ray flag beats the instance flag
    assert hits(base, ray_flags=0x40)[2] == 0 and hits(base, ray_flags=0x80)[2] == n0
    assert hits(base, ray_flags=0x2 | 0x80)[2] == 0 and hits(base, ray_flags=0x2 | 0x40)[2] == n0
    sc = S.sample_scene(300, 200)
    sc.instances[0].flags |= 0x8                      # FORCE_NO_OPAQUE on the upper instance only
    _, p, n = hits(sc, ray_flags=0x80)
    assert 0 < n < n0 and set(np.unique(p["instance_id"][p["instance_id"] != MISS])) == {1}
    assert hits(sc, ray_flags=0x1 | 0x80)[2] == n0
    sc.blases[0][1].flags = 0                         # geometry 1 not opaque
    _, p, n = hits(sc, ray_flags=0x40)                # CullOpaque: instance 0 (forced non-opaque) keeps both, instance 1 keeps geometry 1
    m = p["instance_id"] != MISS
    assert set(zip(p["instance_id"][m].tolist(), p["geometry_index"][m].tolist())) == {(0, 0), (0, 1), (1, 1)}
    # SkipClosestHitShader: hit pixels keep the payload's initial (0,0,0); misses still run the miss shader
    rgba, p, n = hits(base, ray_flags=0x1 | 0x8)
    m = p["instance_id"] != MISS
    assert n == n0 and np.all(rgba[m] == 0) and np.all(rgba[~m] == np.array((0, 0, 51, 0), dtype=np.uint8))
    # TerminateOnFirstHit: same hit/miss mask (which hit is undefined)
    _, p, n = hits(base, ray_flags=0x1 | 0x4)
    assert np.array_equal(p["instance_id"] != MISS, p0["instance_id"] != MISS)
    # miss records
    o = oracle.OracleScene(base)
    o.set_miss_records(np.array([[0, 0, 0.2], [1.0, 0.5, 0.0]], dtype=np.float32))
    rgba = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(miss_index=1))[0]
    assert tuple(rgba[0, 0]) == (255, 128, 0, 0)
    with pytest.raises(RuntimeError):
        o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(miss_index=2))
    o.close()


def _world_triangles(scene, cull_mask=0xFF):
    """All instanced triangles in WORLD space, float64: geometry transform (baked at BLAS build) then instance transform; plus ids.
    Instances whose mask has no bit in common with the ray's cullMask do not exist for the ray [spec: instance culling]."""
    P, ids = [], []
    for ii, I in enumerate(scene.instances):
        if (I.mask & cull_mask) == 0:
            continue
        M = np.asarray(I.transform, dtype=np.float64).reshape(3, 4)
        for gi, g in enumerate(scene.blases[I.blas]):
            v = np.asarray(g.vertices, dtype=np.float64)
            idx = np.asarray(g.indices, dtype=np.int64) if g.indices is not None else np.arange(v.shape[0]).reshape(-1, 3)
            tri = v[idx]                                                     # [nt, 3, 3]
            if g.transform is not None:
                G = np.asarray(g.transform, dtype=np.float64).reshape(3, 4)
                tri = tri @ G[:, :3].T + G[:, 3]
            tri = tri @ M[:, :3].T + M[:, 3]
            P.append(tri)
            n = tri.shape[0]
            ids.append(np.stack([np.full(n, ii), np.full(n, gi), np.arange(n), np.full(n, I.custom_index), np.full(n, I.sbt_offset)], axis=1))
    return np.concatenate(P), np.concatenate(ids)


def _closest_hits_f64(O, D, tri, tmin=0.0, tmax=100.0, t_eps=0.0):
    """Moeller-Trumbore of rays (O, D) against all triangles in float64. Returns (t, triangle index, clear) where clear marks the
    rays whose answer is unambiguous: the closest hit is separated from the runner-up and no candidate lies within 1e-4 of an edge
    (t_eps > 0: nor within t_eps of an end of the ray interval)."""
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    n = D.shape[0]
    best_t = np.full(n, np.inf); second_t = np.full(n, np.inf)
    best_k = np.full(n, -1); best_margin = np.zeros(n)
    rows = np.arange(n)
    for k0 in range(0, tri.shape[0], 256):                                    # chunks of triangles against all rays
        E1, E2, V0 = e1[k0:k0 + 256], e2[k0:k0 + 256], tri[k0:k0 + 256, 0]
        pvec = np.cross(D[:, None, :], E2[None, :, :])
        det = np.einsum("rkc,kc->rk", pvec, E1)
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tvec = O[:, None, :] - V0[None, :, :]
            u = np.einsum("rkc,rkc->rk", tvec, pvec) * inv
            qvec = np.cross(tvec, E1[None, :, :])
            v = np.einsum("rc,rkc->rk", D, qvec) * inv
            t = np.einsum("kc,rkc->rk", E2, qvec) * inv
            margin = np.minimum(np.minimum(u, v), 1.0 - u - v)
            ok = (np.abs(det) > 1e-12) & (margin > 0) & (t > tmin) & (t < tmax)
            near = (np.abs(margin) < 1e-4) & (np.abs(det) > 1e-12) & (t > tmin - 1e-3) & (t < tmax + 1e-3)
            if t_eps > 0.0:
                near |= (margin > -1e-4) & (np.abs(det) > 1e-12) & ((np.abs(t - tmin) < t_eps) | (np.abs(t - tmax) < t_eps))
        t = np.where(ok, t, np.inf)
        kk = np.argmin(t, axis=1)
        tt = t[rows, kk]
        t2 = np.partition(t, 1, axis=1)[:, 1] if t.shape[1] > 1 else np.full(n, np.inf)
        better = tt < best_t
        second_t = np.where(better, np.minimum(best_t, t2), np.minimum(second_t, tt))
        best_margin = np.where(better, margin[rows, kk], best_margin)
        best_k = np.where(better, k0 + kk, best_k)
        best_t = np.where(better, tt, best_t)
        second_t = np.where(near.any(axis=1), -np.inf, second_t)              # a near-edge candidate anywhere: ambiguous ray
    hit = np.isfinite(best_t)
    with np.errstate(invalid="ignore"):
        clear = np.where(hit, (best_margin > 1e-4) & (second_t - best_t > 1e-4 * np.maximum(1.0, best_t)), second_t == np.inf)
    return best_t, best_k, clear


def test_oracle_vs_float64_world_space_intersection(oracle):
    """An INDEPENDENT statement of the whole primary-ray path for a fuzz scene: fp32 raygen as written in the shader, then everything
    else in float64 in world space (geometry transform, instance transform, Moeller-Trumbore instead of the watertight test, no
    world->object matrices at all), SBT rule rec = instanceSbtOffset + geometryIndex. Every ray whose float64 answer is unambiguous
    (closest hit separated from the runner-up, not within 1e-4 of an edge) must get the same instance / geometry / primitive /
    customIndex, t within 1e-4 and the same hit-record colour from the oracle."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=8, seed=5, width=160, height=100, bounces=0, shared_edges=True)
    o = oracle.OracleScene(scene)
    rgba, prim, _, st = o.trace(mode=oracle.MODE_BRUTE)
    o.close()
    W, H = scene.width, scene.height
    f32 = np.float32
    ay = f32(oracle.lib().orc_aspect_y(f32(scene.yfov_deg)))
    ax = f32(ay * f32(W) / f32(H))
    ndcx = ((np.arange(W, dtype=np.float32) + f32(0.5)) / f32(W) * f32(2.0) - f32(1.0)).astype(np.float32)
    ndcy = ((np.arange(H, dtype=np.float32) + f32(0.5)) / f32(H) * f32(2.0) - f32(1.0)).astype(np.float32)
    D = np.empty((H, W, 3))
    D[..., 0] = (ndcx * ax).astype(np.float64)[None, :]
    D[..., 1] = (-(ndcy * ay)).astype(np.float64)[:, None]
    D[..., 2] = -1.0
    D = D.reshape(-1, 3)
    O = np.asarray(scene.camera_pos, dtype=np.float64)
    tri, ids = _world_triangles(scene)
    best_t, best_k, clear = _closest_hits_f64(np.broadcast_to(O, D.shape), D, tri)
    hit64 = np.isfinite(best_t)
    p = prim.reshape(-1)
    ohit = p["instance_id"] != MISS
    n_clear = int(clear.sum())
    assert n_clear > 0.9 * D.shape[0] and int((clear & hit64).sum()) > 1500
    assert np.array_equal(ohit[clear], hit64[clear])
    c = clear & hit64
    want = ids[best_k[c]]
    assert np.array_equal(p["instance_id"][c], want[:, 0]) and np.array_equal(p["geometry_index"][c], want[:, 1])
    assert np.array_equal(p["primitive_id"][c], want[:, 2]) and np.array_equal(p["custom_index"][c], want[:, 3])
    assert np.abs(p["t"][c] - best_t[c]).max() < 1e-4 * best_t[c].max()
    # closest-hit colour: hit record instanceSbtOffset + geometryIndex (main.cpp:1260-1262), except the shader's barycentric special case
    rec = want[:, 4] + want[:, 1]
    special = (want[:, 2] == 1) & (want[:, 0] == 1) & (want[:, 3] == 100) & (want[:, 1] == 1)
    col = np.clip(np.asarray(scene.hit_records, dtype=np.float32)[rec], 0, 1) * np.float32(255)
    got = rgba.reshape(-1, 4)[c][:, :3].astype(np.float64)
    assert np.abs(got - np.rint(col))[~special].max() <= 1.0


def test_interval_cullmask_and_sbt_rule_vs_float64_world_space(oracle):
    """Three more rules the GLSL leaves to the driver, pinned by the same independent float64 world-space statement: the ray interval is OPEN
    (tmin < t < tmax; the sample passes 0.0 / 100.0, main.cpp:1050-1052), an instance exists for a ray iff (mask & cullMask) != 0
    (main.cpp:851 / 1048 pass 0xFF / 0xFF), and the hit record is instanceSbtOffset + geometryIndex * sbtRecordStride + sbtRecordOffset
    (main.cpp:1260-1262 with stride 1 / offset 0 in the sample). The interval is chosen to cut through the scene's own hit distances, the masks
    so that the cullMask removes some instances and keeps others."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=10, seed=21, width=160, height=100, bounces=0, shared_edges=True)
    masks = [0x01, 0x02, 0x04, 0x08, 0x10, 0xFF, 0x03, 0x80, 0x0C, 0x00]
    for i, I in enumerate(scene.instances):
        I.mask = masks[i % len(masks)]
    rng = np.random.default_rng(3)
    scene.hit_records = rng.uniform(0.05, 0.95, size=(12, 3)).astype(np.float32)
    cull_mask, stride, offset = 0x0B, 3, 2
    o = oracle.OracleScene(scene)
    _, p_all, _, _ = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(cull_mask=cull_mask))
    t_all = p_all["t"][p_all["instance_id"] != MISS]
    tmin, tmax = float(np.float32(np.percentile(t_all, 30))), float(np.float32(np.percentile(t_all, 80)))
    rgba, prim, _, _ = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(tmin=tmin, tmax=tmax, cull_mask=cull_mask,
                                                                                   sbt_record_offset=offset, sbt_record_stride=stride))
    o.close()
    O, D = _primary_rays(oracle, scene)
    tri, ids = _world_triangles(scene, cull_mask)
    assert set(np.unique(ids[:, 0])) == {i for i, I in enumerate(scene.instances) if I.mask & cull_mask} != set(range(len(scene.instances)))
    best_t, best_k, clear = _closest_hits_f64(O, D, tri, tmin, tmax, t_eps=1e-3)
    hit64 = np.isfinite(best_t)
    p = prim.reshape(-1)
    ohit = p["instance_id"] != MISS
    assert clear.sum() > 0.85 * D.shape[0] and (clear & hit64).sum() > 500
    # the interval really cuts: some rays lose their nearest surface to tmin and hit something behind it, others lose everything to tmax
    t0, _, _ = _closest_hits_f64(O, D, tri, 0.0, 100.0)
    assert (np.isfinite(t0) & (t0 <= tmin) & hit64).sum() > 50 and (np.isfinite(t0) & ~hit64).sum() > 50
    assert np.array_equal(ohit[clear], hit64[clear])
    c = clear & hit64
    want = ids[best_k[c]]
    assert np.array_equal(p["instance_id"][c], want[:, 0]) and np.array_equal(p["geometry_index"][c], want[:, 1])
    assert np.array_equal(p["primitive_id"][c], want[:, 2]) and np.array_equal(p["custom_index"][c], want[:, 3])
    assert np.all(p["t"][c] > np.float32(tmin)) and np.all(p["t"][c] < np.float32(tmax))
    assert np.abs(p["t"][c] - best_t[c]).max() < 1e-4 * best_t[c].max()
    rec = want[:, 4] + want[:, 1] * stride + offset
    assert rec.max() < len(scene.hit_records) and len(np.unique(rec)) >= 4
    special = (want[:, 2] == 1) & (want[:, 0] == 1) & (want[:, 3] == 100) & (want[:, 1] == 1)
    col = np.clip(np.asarray(scene.hit_records, dtype=np.float32)[rec], 0, 1) * np.float32(255)
    got = rgba.reshape(-1, 4)[c][:, :3].astype(np.float64)
    assert np.abs(got - np.rint(col))[~special].max() <= 1.0
    miss = clear & ~hit64
    assert np.all(rgba.reshape(-1, 4)[miss] == np.array((0, 0, 51, 0), dtype=np.uint8))


def _pcg(v):
    v = np.asarray(v, dtype=np.uint32)
    with np.errstate(over="ignore"):
        state = v * np.uint32(747796405) + np.uint32(2891336453)
        word = ((state >> ((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737)
    return ((word >> np.uint32(22)) ^ word).astype(np.uint32)


def test_bounce_definition_restated(oracle):
    """The diffuse bounce is OUR definition (the reference's recursion depth is 1): restated here from that text in numpy —
    hit point, geometric normal through the instance transform, flip against the ray, first of <= 8 cube-rejection samples from
    pcg_hash(pixel + pcg_hash(seed + 0x9E3779B9)), normalize(n + s), origin p + n * 2^-10 — and traced in float64 world space. On every
    unambiguous secondary ray the oracle must report the same hit / miss, ids and t, and the 0.5 / 0.5 blend of the two colours."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=8, seed=5, width=160, height=100, bounces=1, shared_edges=True)
    o = oracle.OracleScene(scene)
    rgba, prim, sec, st = o.trace(mode=oracle.MODE_BRUTE)
    o.close()
    W, H = scene.width, scene.height
    f32 = np.float32
    ay = f32(oracle.lib().orc_aspect_y(f32(scene.yfov_deg)))
    ax = f32(ay * f32(W) / f32(H))
    ndcx = ((np.arange(W, dtype=np.float32) + f32(0.5)) / f32(W) * f32(2.0) - f32(1.0)).astype(np.float32)
    ndcy = ((np.arange(H, dtype=np.float32) + f32(0.5)) / f32(H) * f32(2.0) - f32(1.0)).astype(np.float32)
    D = np.empty((H, W, 3), dtype=np.float32)
    D[..., 0] = (ndcx * ax)[None, :]; D[..., 1] = (-(ndcy * ay))[:, None]; D[..., 2] = f32(-1.0)
    D = D.reshape(-1, 3)
    O = np.asarray(scene.camera_pos, dtype=np.float32)
    tri, ids = _world_triangles(scene)
    p = prim.reshape(-1); s2 = sec.reshape(-1)
    hit = p["instance_id"] != MISS
    rays = np.nonzero(hit)[0]
    # world-space triangle of every primary hit (looked up by ids), its geometric normal; the oracle derives the same normal from the
    # object-space triangle and the transposed world->object matrix: n_world ~ M^-T (e1 x e2), which is parallel to the world-space
    # cross product up to the sign of det(M) — restated here through the float64 inverse transpose
    key = {tuple(r[:3]): k for k, r in enumerate(ids)}
    k_hit = np.array([key[(int(a), int(b), int(c))] for a, b, c in zip(p["instance_id"][rays], p["geometry_index"][rays], p["primitive_id"][rays])])
    n = np.empty((rays.size, 3))
    for ii, I in enumerate(scene.instances):
        sel = p["instance_id"][rays] == ii
        if not sel.any():
            continue
        M = np.asarray(I.transform, dtype=np.float64).reshape(3, 4)[:, :3]
        Minv = np.linalg.inv(M)
        T = tri[k_hit[sel]]
        obj = (T - np.asarray(I.transform, dtype=np.float64).reshape(3, 4)[:, 3]) @ Minv.T          # back to object space
        n_obj = np.cross(obj[:, 1] - obj[:, 0], obj[:, 2] - obj[:, 0])
        n[sel] = n_obj @ Minv                                                                       # (w2o)^T n
    n /= np.linalg.norm(n, axis=1, keepdims=True)
    d = D[rays].astype(np.float64)
    n = np.where((np.einsum("rc,rc->r", n, d) > 0)[:, None], -n, n)
    P = O.astype(np.float64) + p["t"][rays].astype(np.float64)[:, None] * d
    # cube rejection, hashes exactly as specified (integer arithmetic), samples in fp32 like the product
    h = _pcg(rays.astype(np.uint32) + _pcg(np.uint32(1) + np.uint32(0x9E3779B9)))
    s = np.zeros((rays.size, 3)); found = np.zeros(rays.size, dtype=bool)
    for _ in range(8):
        ha = _pcg(h); hb = _pcg(ha); hc = _pcg(hb); h = hc
        q = np.stack([(x >> np.uint32(8)).astype(np.float32) * f32(2.0 ** -24) * f32(2.0) - f32(1.0) for x in (ha, hb, hc)], axis=1)
        qq = (q[:, 0] * q[:, 0] + q[:, 1] * q[:, 1]) + q[:, 2] * q[:, 2]
        take = ~found & (qq <= f32(1.0)) & (qq > f32(1e-8))
        s[take] = (q[take] / np.sqrt(qq[take])[:, None]).astype(np.float64)
        found |= take
    dirv = n + s
    dl = np.linalg.norm(dirv, axis=1)
    dirv = np.where((dl * dl < 1e-12)[:, None], n, dirv / np.maximum(dl, 1e-30)[:, None])
    O2 = P + n * (2.0 ** -10)
    t2, k2, clear = _closest_hits_f64(O2, dirv, tri)
    hit2 = np.isfinite(t2)
    assert int(clear.sum()) > 0.85 * rays.size and int((clear & hit2).sum()) > 100
    ohit2 = s2["instance_id"][rays] != MISS
    assert np.array_equal(ohit2[clear], hit2[clear])
    c = clear & hit2
    want = ids[k2[c]]
    got = s2[rays][c]
    assert np.array_equal(got["instance_id"], want[:, 0]) and np.array_equal(got["geometry_index"], want[:, 1]) and np.array_equal(got["primitive_id"], want[:, 2])
    assert np.abs(got["t"] - t2[c]).max() < 2e-3          # the ray itself is rebuilt from fp32 t and a float64 normal: looser than the primary check
    # shading: 0.5 * primary colour + 0.5 * (secondary hit colour | miss colour), special barycentric case excluded
    rec1 = ids[k_hit][:, 4] + ids[k_hit][:, 1]
    spec1 = (ids[k_hit][:, 2] == 1) & (ids[k_hit][:, 0] == 1) & (ids[k_hit][:, 3] == 100) & (ids[k_hit][:, 1] == 1)
    col1 = np.asarray(scene.hit_records, dtype=np.float64)[rec1]
    col2 = np.tile(np.asarray(scene.miss_color, dtype=np.float64), (rays.size, 1))
    rec2 = ids[np.maximum(k2, 0)][:, 4] + ids[np.maximum(k2, 0)][:, 1]
    spec2 = hit2 & (ids[np.maximum(k2, 0)][:, 2] == 1) & (ids[np.maximum(k2, 0)][:, 0] == 1) & (ids[np.maximum(k2, 0)][:, 3] == 100) & (ids[np.maximum(k2, 0)][:, 1] == 1)
    col2[hit2] = np.asarray(scene.hit_records, dtype=np.float64)[rec2[hit2]]
    ok = clear & ~spec1 & ~spec2
    expect = np.clip(0.5 * col1 + 0.5 * col2, 0, 1) * 255.0
    assert np.abs(rgba.reshape(-1, 4)[rays][ok][:, :3].astype(np.float64) - expect[ok]).max() <= 1.0


# ---- the any-hit stage (rt_oracle.h orc_anyhit_record; SURVEY 8(f) row 2, stage mapping shader_module.h:90) ------------------------------
def _anyhit_scene(seed=14):
    scene = scenes.random_scene(n_blas=3, tris_per_blas=300, n_instances=8, seed=seed, width=160, height=100, bounces=1, shared_edges=True)
    for b, geoms in enumerate(scene.blases):
        for g, geo in enumerate(geoms):
            geo.flags = 1 if (b + g) % 3 == 0 else 0                      # two thirds of the geometries are NOT opaque
    for i, I in enumerate(scene.instances):
        I.flags = [0x1, 0x1 | 0x8, 0x1, 0x1 | 0x4, 0x1, 0x1, 0x1 | 0x8, 0x1][i % 8]    # some FORCE_NO_OPAQUE / FORCE_OPAQUE
        I.mask = 0xFF
    return scene


def _mask_records(n, seed=3):
    rng = np.random.default_rng(seed)
    recs, bits_of = [], []
    for i in range(n):
        if i % 5 == 4:
            recs.append((0, 0, 0, None)); bits_of.append(None)
            continue
        k = [3, 5, 2, 4, 1][i % 5]
        bits = rng.random((1 << k, 1 << k)) < 0.5                           # [cell_v, cell_u]
        words = np.zeros(((1 << (2 * k)) + 31) // 32, dtype=np.uint32)
        for b in np.nonzero(bits.reshape(-1))[0]:
            words[b >> 5] |= np.uint32(1 << (int(b) & 31))
        recs.append((1, k, 0, words)); bits_of.append(bits)
    return recs, bits_of


def test_anyhit_brute_force_equals_bvh(oracle):
    """With alpha-mask any-hit records the closest ACCEPTED hit is still independent of the traversal order."""
    scene = _anyhit_scene()
    recs, _ = _mask_records(len(scene.hit_records))
    o = oracle.OracleScene(scene)
    rp = oracle.ray_params(ray_flags=0x0)
    base = o.trace(mode=oracle.MODE_BRUTE, ray_params=rp)
    o.set_anyhit_records(recs)
    a = o.trace(mode=oracle.MODE_BRUTE, ray_params=rp)
    b = o.trace(mode=oracle.MODE_BVH, ray_params=rp)
    opaque_ray = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=0x1))
    o.close()
    assert a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes() and np.array_equal(a[0], b[0])
    assert int((a[1]["primitive_id"] != base[1]["primitive_id"]).sum()) > 300, "the masks must cut holes"
    assert opaque_ray[1].tobytes() == base[1].tobytes(), "gl_RayFlagsOpaqueEXT: the any-hit stage never runs"


def test_anyhit_vs_float64_world_space(oracle):
    """INDEPENDENT statement of the any-hit semantics: all candidates of every primary ray in float64 world space (Moeller-Trumbore),
    opacity = geometry flag overridden by the instance FORCE_* flags, non-opaque candidates looked up in the hit group's bit mask at
    cell (int(u * res), int(v * res)) with u -> vertex 1, v -> vertex 2, closest accepted candidate wins. Rays with a candidate within
    1e-4 of an edge or of a mask-cell boundary, or with two accepted candidates closer than 1e-4 in t, are skipped as ambiguous."""
    scene = _anyhit_scene()
    recs, bits_of = _mask_records(len(scene.hit_records))
    o = oracle.OracleScene(scene)
    o.set_anyhit_records(recs)
    _, prim, _, _ = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=0x0), bounces=0)
    o.close()
    W, H = scene.width, scene.height
    f32 = np.float32
    ay = f32(oracle.lib().orc_aspect_y(f32(scene.yfov_deg)))
    ax = f32(ay * f32(W) / f32(H))
    ndcx = ((np.arange(W, dtype=np.float32) + f32(0.5)) / f32(W) * f32(2.0) - f32(1.0)).astype(np.float32)
    ndcy = ((np.arange(H, dtype=np.float32) + f32(0.5)) / f32(H) * f32(2.0) - f32(1.0)).astype(np.float32)
    D = np.empty((H, W, 3))
    D[..., 0] = (ndcx * ax).astype(np.float64)[None, :]
    D[..., 1] = (-(ndcy * ay)).astype(np.float64)[:, None]
    D[..., 2] = -1.0
    D = D.reshape(-1, 3)
    O = np.broadcast_to(np.asarray(scene.camera_pos, dtype=np.float64), D.shape)
    tri, ids = _world_triangles(scene)
    # per world triangle: opaque? which any-hit record?
    geo_flags = {(b, g): geo.flags for b, geoms in enumerate(scene.blases) for g, geo in enumerate(geoms)}
    opaque = np.zeros(tri.shape[0], dtype=bool)
    for k in range(tri.shape[0]):
        I = scene.instances[ids[k, 0]]
        op = bool(geo_flags[(I.blas, int(ids[k, 1]))] & 1)
        if I.flags & 0x4: op = True
        elif I.flags & 0x8: op = False
        opaque[k] = op
    rec = ids[:, 4] + ids[:, 1]                                             # instanceSbtOffset + geometryIndex (stride 1, offset 0)
    n = D.shape[0]
    best_t = np.full(n, np.inf); second_t = np.full(n, np.inf); best_k = np.full(n, -1); ambiguous = np.zeros(n, dtype=bool)
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    for k in range(tri.shape[0]):
        pvec = np.cross(D, e2[k]); det = pvec @ e1[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tvec = O - tri[k, 0]
            u = np.einsum("rc,rc->r", tvec, pvec) * inv
            qvec = np.cross(tvec, e1[k])
            v = np.einsum("rc,rc->r", D, qvec) * inv
            t = (qvec @ e2[k]) * inv
        margin = np.minimum(np.minimum(u, v), 1.0 - u - v)
        in_t = (t > 0.0) & (t < 100.0) & (np.abs(det) > 1e-12)
        ambiguous |= in_t & (np.abs(margin) < 1e-4)
        cand = in_t & (margin > 0)
        if not opaque[k] and rec[k] < len(recs) and bits_of[rec[k]] is not None:
            bits = bits_of[rec[k]]; res = bits.shape[0]
            fu, fv = u * res, v * res
            near_cell = (np.abs(fu - np.rint(fu)) < 1e-4) | (np.abs(fv - np.rint(fv)) < 1e-4)
            ambiguous |= cand & near_cell
            cu = np.clip(np.floor(np.where(cand, fu, 0)).astype(np.int64), 0, res - 1)
            cv = np.clip(np.floor(np.where(cand, fv, 0)).astype(np.int64), 0, res - 1)
            cand &= bits[cv, cu]
        tt = np.where(cand, t, np.inf)
        better = tt < best_t
        second_t = np.where(better, best_t, np.minimum(second_t, tt))
        best_k = np.where(better, k, best_k)
        best_t = np.where(better, tt, best_t)
    hit = np.isfinite(best_t)
    with np.errstate(invalid="ignore"):
        clear = ~ambiguous & np.where(hit, second_t - best_t > 1e-4 * np.maximum(1.0, best_t), True)
    p = prim.reshape(-1)
    ohit = p["instance_id"] != MISS
    assert clear.sum() > 0.85 * n and (clear & hit).sum() > 1200
    assert np.array_equal(ohit[clear], hit[clear])
    c = clear & hit
    want = ids[best_k[c]]
    assert np.array_equal(p["instance_id"][c], want[:, 0]) and np.array_equal(p["geometry_index"][c], want[:, 1])
    assert np.array_equal(p["primitive_id"][c], want[:, 2])
    assert np.abs(p["t"][c] - best_t[c]).max() < 1e-4 * best_t[c].max()
    # the masks really decided: some clear rays pass THROUGH a masked-out candidate that lies in front of their accepted hit
    print("anyhit f64 pin: clear rays", int(clear.sum()), "of", n, "hits", int(c.sum()))


def test_oracle_refit_keeps_topology_and_stays_exact(oracle):
    """orc_refit_blas (the restatement of RT_BUILD_MODE_REFIT): the sorted order of the last full build is kept, every box is re-fitted
    to the moved triangles, and the BVH traversal still equals the brute force over the new geometry."""
    S = scenes
    g0 = S.heightfield(40, 30, -3.0, 3.0, -2.0, 2.0, 0.5, 40)
    v = g0.vertices.copy()
    v[:, 2] = (v[:, 2] * np.float32(1.4) + np.float32(0.1) * np.sin(v[:, 0] * np.float32(2.0))).astype(np.float32)
    g1 = S.Geometry(np.ascontiguousarray(v), g0.indices, None)
    inst = [S.Instance(S.rotation_3x4(np.array([0.2, 1.0, 0.0]), 0.3, np.array([0.0, 0.0, 0.0])), 3, 0xFF, 0, 1, 0)]
    scene = S.Scene("refit", [[g0]], inst, S.SAMPLE_HIT_RECORDS[:1].copy(), width=200, height=120, bounces=1)
    o = oracle.OracleScene(scene)
    _, n0, t0, k0, p0 = o.blas_export(0)
    o.refit_blas(0, [g1])
    _, n1, t1, k1, p1 = o.blas_export(0)
    assert np.array_equal(k0, k1) and np.array_equal(p0, p1), "keys and order of the last full build are kept"
    assert not np.array_equal(n0, n1) and not np.array_equal(t0, t1)
    a = o.trace(mode=oracle.MODE_BVH)
    o.close()
    fresh = oracle.OracleScene(S.Scene("refit", [[g1]], inst, scene.hit_records, width=200, height=120, bounces=1))
    b = fresh.trace(mode=oracle.MODE_BRUTE)
    fresh.close()
    assert a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes() and np.array_equal(a[0], b[0])
    assert a[3]["primary_hits"] > 2000


# ---- float64 pin of the whole flag surface: opacity resolution, opacity culls, facing culls with FLIP / CULL_DISABLE -------------------------
def _primary_rays(oracle, scene):
    W, H = scene.width, scene.height
    f32 = np.float32
    ay = f32(oracle.lib().orc_aspect_y(f32(scene.yfov_deg)))
    ax = f32(ay * f32(W) / f32(H))
    ndcx = ((np.arange(W, dtype=np.float32) + f32(0.5)) / f32(W) * f32(2.0) - f32(1.0)).astype(np.float32)
    ndcy = ((np.arange(H, dtype=np.float32) + f32(0.5)) / f32(H) * f32(2.0) - f32(1.0)).astype(np.float32)
    D = np.empty((H, W, 3))
    D[..., 0] = (ndcx * ax).astype(np.float64)[None, :]
    D[..., 1] = (-(ndcy * ay)).astype(np.float64)[:, None]
    D[..., 2] = -1.0
    D = D.reshape(-1, 3)
    return np.broadcast_to(np.asarray(scene.camera_pos, dtype=np.float64), D.shape), D


def _flagged_closest_f64(scene, O, D, ray_flags):
    """Closest ACCEPTED candidate per ray, float64, world space, straight from the Vulkan rules: opacity = geometry OPAQUE bit, overridden by
    the instance FORCE_OPAQUE / FORCE_NO_OPAQUE, overridden by the ray Opaque / NoOpaque; CullOpaque / CullNoOpaque; a triangle is front
    facing when ((v1-v0) x (v2-v0)) . d > 0 in OBJECT space (= world space times the sign of the instance matrix determinant), inverted by
    FLIP_FACING; facing culls are off for instances with TRIANGLE_FACING_CULL_DISABLE. Returns (t, triangle index, clear, ids)."""
    tri, ids = _world_triangles(scene)
    geo_flags = {(b, g): geo.flags for b, geoms in enumerate(scene.blases) for g, geo in enumerate(geoms)}
    n = D.shape[0]
    best_t = np.full(n, np.inf); second_t = np.full(n, np.inf); best_k = np.full(n, -1); ambiguous = np.zeros(n, dtype=bool)
    e1, e2 = tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    for k in range(tri.shape[0]):
        I = scene.instances[ids[k, 0]]
        opaque = bool(geo_flags[(I.blas, int(ids[k, 1]))] & 1)
        if I.flags & 0x4: opaque = True
        elif I.flags & 0x8: opaque = False
        if ray_flags & 0x1: opaque = True
        elif ray_flags & 0x2: opaque = False
        if (opaque and ray_flags & 0x40) or (not opaque and ray_flags & 0x80):
            continue
        pvec = np.cross(D, e2[k]); det = pvec @ e1[k]
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = 1.0 / det
            tvec = O - tri[k, 0]
            u = np.einsum("rc,rc->r", tvec, pvec) * inv
            qvec = np.cross(tvec, e1[k])
            v = np.einsum("rc,rc->r", D, qvec) * inv
            t = (qvec @ e2[k]) * inv
        margin = np.minimum(np.minimum(u, v), 1.0 - u - v)
        in_t = (t > 0.0) & (t < 100.0) & (np.abs(det) > 1e-12)
        ambiguous |= in_t & (np.abs(margin) < 1e-4)
        cand = in_t & (margin > 0)
        if (ray_flags & 0x30) and not (I.flags & 0x1):
            A = np.asarray(I.transform, dtype=np.float64).reshape(3, 4)[:, :3]
            facing = np.sign(np.linalg.det(A)) * (D @ np.cross(e1[k], e2[k]))        # > 0: front (clockwise from the ray origin), object space
            ambiguous |= cand & (np.abs(facing) < 1e-9)
            front = (facing > 0) != bool(I.flags & 0x2)
            cand &= ~np.where(front, bool(ray_flags & 0x20), bool(ray_flags & 0x10))
        tt = np.where(cand, t, np.inf)
        better = tt < best_t
        second_t = np.where(better, best_t, np.minimum(second_t, tt))
        best_k = np.where(better, k, best_k)
        best_t = np.where(better, tt, best_t)
    hit = np.isfinite(best_t)
    with np.errstate(invalid="ignore"):
        clear = ~ambiguous & np.where(hit, second_t - best_t > 1e-4 * np.maximum(1.0, best_t), True)
    return best_t, best_k, clear, ids


@pytest.mark.parametrize("ray_flags", [0x10, 0x20, 0x40, 0x80, 0x2 | 0x20 | 0x80, 0x2 | 0x20, 0x1 | 0x10, 0x0], ids=lambda f: f"rayflags{f:#04x}")
def test_ray_flags_vs_float64_world_space(oracle, ray_flags):
    """The flag semantics pinned independently of the oracle's own arithmetic (the culls had known-answer pins on the
    8-triangle scene only): every unambiguous primary ray of a fuzz scene with non-opaque geometries and FLIP / FORCE_* / CULL_DISABLE
    instances gets the same hit / miss and the same ids from the float64 statement above and from the oracle."""
    scene = scenes.random_scene(n_blas=3, tris_per_blas=250, n_instances=10, seed=9, width=160, height=100, bounces=0, shared_edges=True)
    for b, geoms in enumerate(scene.blases):
        for g, geo in enumerate(geoms):
            geo.flags = 1 if (b + g) % 2 == 0 else 0
    for i, I in enumerate(scene.instances):
        I.flags = [0x0, 0x1, 0x2, 0x4, 0x8, 0x2 | 0x8, 0x0, 0x4 | 0x2, 0x1 | 0x8, 0x0][i % 10]
        I.mask = 0xFF
    o = oracle.OracleScene(scene)
    _, prim, _, _ = o.trace(mode=oracle.MODE_BRUTE, ray_params=oracle.ray_params(ray_flags=ray_flags))
    o.close()
    O, D = _primary_rays(oracle, scene)
    best_t, best_k, clear, ids = _flagged_closest_f64(scene, O, D, ray_flags)
    hit = np.isfinite(best_t)
    p = prim.reshape(-1)
    ohit = p["instance_id"] != MISS
    if (ray_flags & 0x2) and (ray_flags & 0x80):
        assert not hit.any() and not ohit.any(), "NoOpaque makes every candidate non-opaque and CullNoOpaque drops them all"
        return
    assert clear.sum() > 0.85 * D.shape[0] and (clear & hit).sum() > 250
    assert np.array_equal(ohit[clear], hit[clear])
    c = clear & hit
    want = ids[best_k[c]]
    assert np.array_equal(p["instance_id"][c], want[:, 0]) and np.array_equal(p["geometry_index"][c], want[:, 1])
    assert np.array_equal(p["primitive_id"][c], want[:, 2])
    assert np.abs(p["t"][c] - best_t[c]).max() < 1e-4 * best_t[c].max()
