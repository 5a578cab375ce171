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


def walk_compare_bvh_renumbered(nodes_a, root_a, nodes_b, root_b):
    """Lock-step walk of two BVHs that may number their internal nodes differently (e.g. before / after compaction): boxes and heights
    bit-exact, leaf refs equal, internal refs followed pairwise. Returns the list of (index_a, index_b) pairs of the reachable nodes."""
    def internal(r):
        return 0 <= r < 0x7FFFFFF0
    if internal(root_a) != internal(root_b) or (not internal(root_a) and root_a != root_b):
        raise AssertionError(f"root {root_a} vs {root_b}")
    pairs = []
    stack = [(root_a, root_b)] if internal(root_a) else []
    while stack:
        ra, rb = stack.pop()
        pairs.append((ra, rb))
        a, b = nodes_a[ra], nodes_b[rb]
        for half in (0, 1):
            ha, hb = a[8 * half:8 * half + 8], b[8 * half:8 * half + 8]
            if not (np.array_equal(ha[:6], hb[:6]) and ha[7] == hb[7]):
                raise AssertionError(f"node {ra}/{rb} half {half}: {ha} != {hb}")
            fa, fb = int(np.int32(ha[6])), int(np.int32(hb[6]))
            if internal(fa) != internal(fb) or (not internal(fa) and fa != fb):
                raise AssertionError(f"node {ra}/{rb} half {half}: refs {fa} vs {fb}")
            if internal(fa):
                stack.append((fa, fb))
    return pairs
