"""Comparison helpers shared by the GPU parity tests, smoke() and bench.py."""
import numpy as np

ID_FIELDS = ("instance_id", "geometry_index", "primitive_id", "custom_index")
MISS = 0xFFFFFFFF
EDGE_EPS = 2.0 ** -20          # "within epsilon of a shared edge": min barycentric weight below this


def compare_hits(gpu, ref, rows=None):
    """Bit-exact comparison of rt_hit arrays. Returns a report dict; rows = slice of image rows that
    the reference traced (others are ignored)."""
    if rows is not None:
        gpu, ref = gpu[rows], ref[rows]
    rep = {"rays": int(gpu.size)}
    bad = np.zeros(gpu.shape, dtype=bool)
    for f in ID_FIELDS:
        bad |= gpu[f] != ref[f]
    rep["id_mismatches"] = int(bad.sum())
    hit = ref["instance_id"] != MISS
    rep["hits"] = int(hit.sum())
    w0 = 1.0 - ref["u"].astype(np.float64) - ref["v"].astype(np.float64)
    near = hit & (np.minimum(np.minimum(ref["u"], ref["v"]), w0) < EDGE_EPS)
    rep["near_edge_rays"] = int(near.sum())
    rep["id_mismatches_near_edge"] = int((bad & near).sum())
    for f in ("t", "u", "v"):
        rep[f + "_bit_mismatches"] = int((gpu[f].view(np.uint32) != ref[f].view(np.uint32)).sum())
    anybad = bad.copy()
    for f in ("t", "u", "v"):
        anybad |= gpu[f].view(np.uint32) != ref[f].view(np.uint32)
    if anybad.any():
        ys, xs = np.nonzero(anybad)
        rep["first_mismatches"] = [(int(y), int(x), tuple(gpu[y, x].tolist()), tuple(ref[y, x].tolist()),
                                    [hex(int(v)) for v in gpu[y, x].tolist()[:0]] + [hex(int(np.float32(gpu[y, x][f]).view(np.uint32))) for f in ("t", "u", "v")],
                                    [hex(int(np.float32(ref[y, x][f]).view(np.uint32))) for f in ("t", "u", "v")])
                                   for y, x in list(zip(ys, xs))[:5]]
    return rep


def compare_rgba(gpu, ref, rows=None):
    if rows is not None:
        gpu, ref = gpu[rows], ref[rows]
    d = np.abs(gpu.astype(np.int16) - ref.astype(np.int16))
    return {"max_abs_diff": int(d.max()) if d.size else 0, "pixels_differing": int((d.max(axis=-1) > 0).sum()) if d.size else 0}


def assert_parity(gpu_out, ref_out, rows=None, what=""):
    """ids / t / u / v bit-exact; RGBA8 within +-1 LSB (the Vulkan UNORM rounding allowance)."""
    rgba_g, prim_g, sec_g = gpu_out[:3]
    rgba_r, prim_r, sec_r = ref_out[:3]
    rp = compare_hits(prim_g, prim_r, rows)
    assert rp["id_mismatches"] == 0, f"{what} primary ids: {rp}"
    assert rp["t_bit_mismatches"] == 0 and rp["u_bit_mismatches"] == 0 and rp["v_bit_mismatches"] == 0, f"{what} primary t/u/v: {rp}"
    rs = None
    if sec_g is not None and sec_r is not None:
        rs = compare_hits(sec_g, sec_r, rows)
        assert rs["id_mismatches"] == 0, f"{what} secondary ids: {rs}"
        assert rs["t_bit_mismatches"] == 0 and rs["u_bit_mismatches"] == 0 and rs["v_bit_mismatches"] == 0, f"{what} secondary t/u/v: {rs}"
    rc = compare_rgba(rgba_g, rgba_r, rows)
    assert rc["max_abs_diff"] <= 1, f"{what} rgba: {rc}"      # tolerance stated by north_star: +-1 LSB of the 8-bit image
    return rp, rs, rc


def walk_compare_bvh(nodes_a, root_a, nodes_b, root_b):
    """Lock-step walk of two exported BVHs (uint32[n,16] node arrays); returns the number of reachable
    internal nodes if identical (boxes bit-exact, refs equal), else raises AssertionError."""
    if root_a != root_b:
        raise AssertionError(f"root {root_a} != {root_b}")
    count = 0
    stack = [root_a] if 0 <= root_a < 0x7FFFFFF0 else []
    while stack:
        r = stack.pop()
        a, b = nodes_a[r], nodes_b[r]
        for half in (0, 1):
            ha, hb = a[8 * half:8 * half + 8], b[8 * half:8 * half + 8]
            if not np.array_equal(ha[:7], hb[:7]):
                raise AssertionError(f"node {r} half {half}: {ha} != {hb}")
            ref = int(np.int32(ha[6]))
            if 0 <= ref < 0x7FFFFFF0:
                stack.append(ref)
        count += 1
    return count


def _popcount24(v):
    v = v.astype(np.int64)
    c = np.zeros(v.shape, dtype=np.int64)
    for b in range(24):
        c += (v >> b) & 1
    return c


def check_wide_bvh(wnodes, tris, root, bounds_lo, bounds_hi):
    """Structural invariants of an exported 8-wide BVH (uint32[n,20] nodes, uint32[m,12] triangles in wide order):
    every node referenced exactly once, every triangle in exactly one leaf slot, imask/prim_valid consistent, and every
    triangle inside the dequantised box of EVERY ancestor slot on its path (so no conservative box can cull it).
    Returns the number of reachable nodes. Vectorised level by level."""
    n_nodes, n_tris = wnodes.shape[0], tris.shape[0]
    if root < 0 or root >= 0x7FFFFFF0:
        assert n_tris == 0
        return 0
    b = np.ascontiguousarray(wnodes).view(np.uint8).reshape(n_nodes, 80)
    p = np.ascontiguousarray(wnodes[:, 0:3]).view(np.float32).astype(np.float64)
    e = b[:, 12:15].astype(np.int64)
    step = np.ldexp(1.0, e - 127)
    imask = b[:, 15].astype(np.int64)
    child_base, prim_base = wnodes[:, 4].astype(np.int64), wnodes[:, 5].astype(np.int64)
    prim_valid = wnodes[:, 6].astype(np.int64)
    q = b[:, 32:80].reshape(n_nodes, 6, 8).astype(np.float64)          # qlo.x qlo.y qlo.z qhi.x qhi.y qhi.z  x 8 slots
    lo = p[:, :, None] + q[:, 0:3, :] * step[:, :, None]               # [n, 3, 8]
    hi = p[:, :, None] + q[:, 3:6, :] * step[:, :, None]
    tv = np.ascontiguousarray(tris[:, :9]).view(np.float32).reshape(n_tris, 3, 3).astype(np.float64)
    tlo, thi = tv.min(axis=1), tv.max(axis=1)
    seen_nodes = np.zeros(n_nodes, dtype=np.int32)
    seen_tris = np.zeros(n_tris, dtype=np.int32)
    frontier = np.array([root], dtype=np.int64)
    acc_lo = np.array([bounds_lo], dtype=np.float64)
    acc_hi = np.array([bounds_hi], dtype=np.float64)
    depth = 0
    while frontier.size:
        np.add.at(seen_nodes, frontier, 1)
        nxt, nlo, nhi = [], [], []
        rank = np.zeros(frontier.size, dtype=np.int64)
        pv = prim_valid[frontier]
        assert np.all(pv < (1 << 24))
        for s in range(8):
            inner = ((imask[frontier] >> s) & 1) == 1
            f3 = (pv >> (3 * s)) & 7
            assert np.all(np.isin(f3, (0, 1, 3, 7))), "bad primitive field"
            assert np.all(f3[inner] == 0), "internal slot with primitive bits"
            leaf = f3 != 0
            clo = np.maximum(lo[frontier, :, s], acc_lo)
            chi = np.minimum(hi[frontier, :, s], acc_hi)
            cnt = np.where(f3 == 1, 1, np.where(f3 == 3, 2, np.where(f3 == 7, 3, 0)))
            below = pv & ((1 << (3 * s)) - 1)
            off = _popcount24(below)
            for k in range(3):
                sel = leaf & (cnt > k)
                ti = prim_base[frontier][sel] + off[sel] + k
                assert np.all(ti < n_tris)
                np.add.at(seen_tris, ti, 1)
                assert np.all(tlo[ti] >= clo[sel]) and np.all(thi[ti] <= chi[sel]), "a triangle escapes a box on its path"
            nxt.append(child_base[frontier][inner] + rank[inner]); nlo.append(clo[inner]); nhi.append(chi[inner])
            rank += inner
        frontier = np.concatenate(nxt); acc_lo = np.concatenate(nlo); acc_hi = np.concatenate(nhi)
        assert np.all(frontier < n_nodes)
        depth += 1
        assert depth < 200
    assert np.all(seen_tris == 1), "a triangle is not in exactly one leaf slot"
    assert np.all(seen_nodes <= 1), "a wide node is referenced twice"
    return int((seen_nodes > 0).sum())
