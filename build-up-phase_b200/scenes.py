"""Scene descriptions: the reference sample scene (literal mirror of the host arrays in
vulkan-raytracing-basic/main.cpp) and the synthetic procedural scenes of BASELINE.json's configs.

A Scene is plain numpy data — exactly what the reference's createBLAS/createTLAS/createUniformBuffer/
createShaderBindingTable put into host-visible buffers — and is consumed unchanged by the CUDA
product (rtcore.py) and by the CPU checker used in tests. Nothing here
computes ray tracing.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

IDENTITY_3X4 = np.array([1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0], dtype=np.float32)

# VK_GEOMETRY_INSTANCE_TRIANGLE_FACING_CULL_DISABLE_BIT_KHR (main.cpp:852)
INSTANCE_TRIANGLE_FACING_CULL_DISABLE = 0x1


@dataclass
class Geometry:
    """One VkAccelerationStructureGeometryKHR of type TRIANGLES (main.cpp:726-742)."""
    vertices: np.ndarray                     # float32 [nv, 3] (stride 12, R32G32B32_SFLOAT)
    indices: Optional[np.ndarray]            # uint32 [nt, 3] or None for a non-indexed list
    transform: Optional[np.ndarray] = None   # float32 [12], row-major 3x4 (VkTransformMatrixKHR)
    flags: int = 1                           # VkGeometryFlagsKHR; 1 = OPAQUE (main.cpp:741)

    @property
    def triangle_count(self) -> int:
        return int(self.indices.shape[0]) if self.indices is not None else int(self.vertices.shape[0] // 3)


@dataclass
class Instance:
    """One VkAccelerationStructureInstanceKHR (main.cpp:848-858)."""
    transform: np.ndarray        # float32 [12]
    custom_index: int
    mask: int
    sbt_offset: int
    flags: int
    blas: int                    # index into Scene.blases


@dataclass
class Scene:
    name: str
    blases: List[List[Geometry]]
    instances: List[Instance]
    hit_records: np.ndarray                      # float32 [n, 3] — SBT hit-group payloads (main.cpp:1310-1317)
    miss_color: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 0.2], dtype=np.float32))  # main.cpp:1065
    camera_pos: np.ndarray = field(default_factory=lambda: np.array([0.0, 0.0, 10.0], dtype=np.float32))  # main.cpp:1015
    yfov_deg: float = 60.0
    width: int = 1200                            # main.cpp:13
    height: int = 800                            # main.cpp:14
    bounces: int = 0

    @property
    def triangle_count(self) -> int:
        return sum(g.triangle_count for b in self.blases for g in b)

    @property
    def instanced_triangle_count(self) -> int:
        per = [sum(g.triangle_count for g in b) for b in self.blases]
        return sum(per[i.blas] for i in self.instances)


def translation(tx: float, ty: float, tz: float) -> np.ndarray:
    return np.array([1, 0, 0, tx, 0, 1, 0, ty, 0, 0, 1, tz], dtype=np.float32)


SAMPLE_HIT_RECORDS = np.array(
    [[0.6, 0.1, 0.2],    # Deep Red Wine      main.cpp:1311
     [0.1, 0.8, 0.4],    # Emerald Green      main.cpp:1313
     [0.9, 0.7, 0.1],    # Golden Yellow      main.cpp:1315
     [0.3, 0.6, 0.9]],   # Dawn Sky Blue      main.cpp:1317
    dtype=np.float32)


def sample_scene(width: int = 1200, height: int = 800) -> Scene:
    """The scene both RT samples render (their main.cpp are byte-identical): one BLAS of two
    geometries sharing a quad (main.cpp:676-695,743), two instances of it (main.cpp:835-858)."""
    vertices = np.array([[-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.float32)   # main.cpp:676-681
    indices = np.array([[0, 1, 3], [1, 2, 3]], dtype=np.uint32)                               # main.cpp:682
    geo_transforms = [translation(-2, 0, 0), translation(2, 0, 0)]                            # main.cpp:684-695
    blas = [Geometry(vertices, indices, geo_transforms[0]), Geometry(vertices, indices, geo_transforms[1])]
    ins_transforms = [translation(0, 2, 0), translation(0, -2, 0)]                            # main.cpp:835-846
    instances = [
        Instance(ins_transforms[0], 100, 0xFF, 0, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0),  # main.cpp:848-856
        Instance(ins_transforms[1], 100, 0xFF, 2, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0),  # main.cpp:857-858
    ]
    return Scene("sample", [blas], instances, SAMPLE_HIT_RECORDS.copy(), width=width, height=height)


def single_triangle_scene(width: int = 1200, height: int = 800) -> Scene:
    """Literal reading of BASELINE.json config 1 ("single-triangle BLAS/TLAS")."""
    vertices = np.array([[-1, -1, 0], [1, -1, 0], [0, 1, 0]], dtype=np.float32)
    indices = np.array([[0, 1, 2]], dtype=np.uint32)
    inst = Instance(IDENTITY_3X4.copy(), 7, 0xFF, 0, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0)
    return Scene("triangle", [[Geometry(vertices, indices, None)]], [inst], SAMPLE_HIT_RECORDS[:1].copy(),
                 width=width, height=height)


# ------------------------------------------------------------------------------------------------
# procedural helpers (generation only; both arms receive the resulting arrays)
# ------------------------------------------------------------------------------------------------
def pcg_hash(v: np.ndarray) -> np.ndarray:
    v = np.asarray(v, dtype=np.uint32)
    with np.errstate(over="ignore"):
        state = v * np.uint32(747796405) + np.uint32(2891336453)
        word = ((state >> ((state >> np.uint32(28)) + np.uint32(4))) ^ state) * np.uint32(277803737)
        return (word >> np.uint32(22)) ^ word


def _hash01(ix: np.ndarray, iy: np.ndarray, octave: int, seed: int) -> np.ndarray:
    with np.errstate(over="ignore"):
        h = pcg_hash(ix.astype(np.uint32) + pcg_hash(iy.astype(np.uint32) + pcg_hash(np.uint32(seed * 131 + octave))))
    return (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def value_noise(x: np.ndarray, y: np.ndarray, seed: int, octaves: int = 4, base_freq: float = 0.5) -> np.ndarray:
    """Sum of `octaves` octaves of lattice value noise in [0, ~1)."""
    out = np.zeros(np.broadcast(x, y).shape, dtype=np.float32)
    amp = np.float32(0.5)
    freq = np.float32(base_freq)
    for o in range(octaves):
        fx = (x * freq).astype(np.float32)
        fy = (y * freq).astype(np.float32)
        ix = np.floor(fx)
        iy = np.floor(fy)
        tx = (fx - ix).astype(np.float32)
        ty = (fy - iy).astype(np.float32)
        tx = tx * tx * (np.float32(3) - np.float32(2) * tx)
        ty = ty * ty * (np.float32(3) - np.float32(2) * ty)
        ixi = ix.astype(np.int64) + 100000
        iyi = iy.astype(np.int64) + 100000
        h00 = _hash01(ixi, iyi, o, seed)
        h10 = _hash01(ixi + 1, iyi, o, seed)
        h01 = _hash01(ixi, iyi + 1, o, seed)
        h11 = _hash01(ixi + 1, iyi + 1, o, seed)
        a = h00 + (h10 - h00) * tx
        b = h01 + (h11 - h01) * tx
        out = out + amp * (a + (b - a) * ty)
        amp = amp * np.float32(0.5)
        freq = freq * np.float32(2.0)
    return out.astype(np.float32)


def grid_indices(nx: int, ny: int) -> np.ndarray:
    """Two triangles per quad of an (nx x ny)-quad grid with (nx+1) vertices per row, split along the
    v1-v3 diagonal like the sample's quad (indices {0,1,3, 1,2,3}, main.cpp:682)."""
    i, j = np.meshgrid(np.arange(nx, dtype=np.uint32), np.arange(ny, dtype=np.uint32), indexing="xy")
    v0 = (j * np.uint32(nx + 1) + i).ravel()
    v1 = v0 + np.uint32(1)
    v2 = v0 + np.uint32(nx + 2)
    v3 = v0 + np.uint32(nx + 1)
    tris = np.empty((nx * ny, 2, 3), dtype=np.uint32)
    tris[:, 0, 0], tris[:, 0, 1], tris[:, 0, 2] = v0, v1, v3
    tris[:, 1, 0], tris[:, 1, 1], tris[:, 1, 2] = v1, v2, v3
    return tris.reshape(-1, 3)


def heightfield(nx: int, ny: int, x0: float, x1: float, y0: float, y1: float, amp: float, seed: int,
                base_freq: float = 0.5) -> Geometry:
    xs = np.linspace(x0, x1, nx + 1, dtype=np.float32)
    ys = np.linspace(y0, y1, ny + 1, dtype=np.float32)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    Z = (np.float32(amp) * (value_noise(X, Y, seed, 4, base_freq) - np.float32(0.45))).astype(np.float32)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    return Geometry(np.ascontiguousarray(verts), grid_indices(nx, ny), None)


def displaced_sphere(n: int, radius: float, amp: float, seed: int) -> Geometry:
    """(n x n)-quad latitude/longitude sphere with radial value-noise displacement: 2*n*n triangles."""
    u = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
    v = np.linspace(0.0, 1.0, n + 1, dtype=np.float32)
    U, V = np.meshgrid(u, v, indexing="xy")
    theta = (U * np.float32(2 * np.pi)).astype(np.float32)
    phi = (V * np.float32(np.pi)).astype(np.float32)
    r = (np.float32(radius) * (np.float32(1.0) + np.float32(amp) * (value_noise(U * 8, V * 8, seed, 3, 1.0) - np.float32(0.45)))).astype(np.float32)
    X = r * np.sin(phi) * np.cos(theta)
    Y = r * np.sin(phi) * np.sin(theta)
    Z = r * np.cos(phi)
    verts = np.stack([X, Y, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    return Geometry(np.ascontiguousarray(verts), grid_indices(n, n), None)


def rotation_3x4(axis: np.ndarray, angle: float, t: np.ndarray, scale: float = 1.0) -> np.ndarray:
    a = np.asarray(axis, dtype=np.float64)
    a = a / np.linalg.norm(a)
    c, s = np.cos(angle), np.sin(angle)
    x, y, z = a
    R = np.array([[c + x * x * (1 - c), x * y * (1 - c) - z * s, x * z * (1 - c) + y * s],
                  [y * x * (1 - c) + z * s, c + y * y * (1 - c), y * z * (1 - c) - x * s],
                  [z * x * (1 - c) - y * s, z * y * (1 - c) + x * s, c + z * z * (1 - c)]]) * scale
    m = np.concatenate([R, np.asarray(t, dtype=np.float64).reshape(3, 1)], axis=1)
    return m.reshape(-1).astype(np.float32)


def tess_scene(nx: int = 1000, ny: int = 500, width: int = 3840, height: int = 2160, bounces: int = 1, seed: int = 1) -> Scene:
    """cfg3 "tess1m": height-field grid of nx x ny quads (1000 x 500 -> exactly 1,000,000 triangles,
    501,501 vertices) over [-8,8] x [-4,4], one geometry, one identity instance, one hit record."""
    geo = heightfield(nx, ny, -8.0, 8.0, -4.0, 4.0, 1.0, seed)
    inst = Instance(IDENTITY_3X4.copy(), 3, 0xFF, 0, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0)
    return Scene(f"tess{nx * ny * 2}", [[geo]], [inst], SAMPLE_HIT_RECORDS[1:2].copy(), width=width, height=height, bounces=bounces)


def instanced_scene(n_side: int = 32, quads: int = 70, width: int = 3840, height: int = 2160, bounces: int = 1,
                    seed: int = 100) -> Scene:
    """cfg4 "inst10m": n_side^2 DISTINCT BLASes (32^2 = 1024), each 2*quads^2 triangles (70 -> 9,800;
    total 10,035,200), one instance each on an n_side x n_side grid in the z=0 plane with a
    per-instance rotation, sbt_offset = instance % 4 and the reference's four hit records."""
    n_blas = n_side * n_side
    blases: List[List[Geometry]] = []
    instances: List[Instance] = []
    span_x, span_y = 19.0, 10.6
    cell_x, cell_y = span_x / n_side, span_y / n_side
    rad = 0.46 * min(cell_x, cell_y)
    rng = np.random.default_rng(seed)
    for b in range(n_blas):
        if b % 2 == 0:
            g = heightfield(quads, quads, -rad, rad, -rad, rad, 0.6 * rad, seed + b, base_freq=3.0 / rad)
        else:
            g = displaced_sphere(quads, 0.9 * rad, 0.35, seed + b)
        blases.append([g])
        ix, iy = b % n_side, b // n_side
        cx = -span_x / 2 + (ix + 0.5) * cell_x
        cy = -span_y / 2 + (iy + 0.5) * cell_y
        axis = rng.normal(size=3)
        angle = float(rng.uniform(-0.6, 0.6))
        m = rotation_3x4(axis, angle, np.array([cx, cy, 0.0]))
        instances.append(Instance(m, b, 0xFF, b % 4, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, b))
    return Scene(f"inst{n_blas}x{2 * quads * quads}", blases, instances, SAMPLE_HIT_RECORDS.copy(),
                 width=width, height=height, bounces=bounces)


SOUP_PARTS = 8     # cfg5 is always cut into 8 contiguous index ranges = 8 BLASes, whatever the GPU count


def soup_part(n_tris: int, part: int, parts: int = SOUP_PARTS, seed: int = 7, edge: float = 0.01, split: str = "slab") -> Geometry:
    """Triangles [part*n/parts, (part+1)*n/parts) of the cfg5 soup as one non-indexed geometry. Every part has its own
    generator stream, so a rank can generate just the parts it builds and the scene is identical for every GPU count.
    split="slab": part p holds the triangles whose centroid lies in the p-th of `parts` equal x-slabs of the volume (a
    spatially partitioned data set: the BLASes do not overlap, the overall density is still uniform). split="index":
    every part is uniform over the whole volume (all BLASes overlap completely; every ray has to traverse all of them)."""
    first, last = part * n_tris // parts, (part + 1) * n_tris // parts
    n = last - first
    rng = np.random.default_rng([seed, part])
    verts = np.empty((n, 3, 3), dtype=np.float32)
    chunk = 4_000_000
    lo = np.array([-10.0, -5.6, -6.0], dtype=np.float32)
    hi = np.array([10.0, 5.6, 2.0], dtype=np.float32)
    if split == "slab":
        w = (hi[0] - lo[0]) / np.float32(parts)
        lo, hi = lo.copy(), hi.copy()
        lo[0], hi[0] = lo[0] + np.float32(part) * w, lo[0] + np.float32(part + 1) * w
    elif split != "index":
        raise ValueError(split)
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        c = rng.random((e - s, 3), dtype=np.float32) * (hi - lo) + lo
        e1 = (rng.random((e - s, 3), dtype=np.float32) * np.float32(2) - np.float32(1)) * np.float32(edge)
        e2 = (rng.random((e - s, 3), dtype=np.float32) * np.float32(2) - np.float32(1)) * np.float32(edge)
        verts[s:e, 0] = c
        verts[s:e, 1] = c + e1
        verts[s:e, 2] = c + e2
    return Geometry(verts.reshape(-1, 3), None, None)


def soup_scene(n_tris: int = 100_000_000, width: int = 7680, height: int = 4320, bounces: int = 0, seed: int = 7,
               edge: float = 0.01, parts: int = SOUP_PARTS, only_parts=None, split: str = "slab") -> Scene:
    """cfg5 "soup100m": n independent triangles; centroid uniform in a [-10,10] x [-5.6,5.6] x [-6,2] slab in
    front of the camera, two edge vectors uniform in [-edge, edge]^3; non-indexed (indices = None). The soup is
    cut into `parts` contiguous index ranges, one BLAS each (per-GPU BLAS builds), instanced with
    the identity. only_parts: generate just these parts (the others get an empty placeholder geometry)."""
    blases, instances = [], []
    for p in range(parts):
        if only_parts is None or p in only_parts:
            geo = soup_part(n_tris, p, parts, seed, edge, split)
        else:
            geo = Geometry(np.zeros((0, 3), dtype=np.float32), None, None)
        blases.append([geo])
        instances.append(Instance(IDENTITY_3X4.copy(), 5, 0xFF, 0, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, p))
    return Scene(f"soup{n_tris}-{split}", blases, instances, SAMPLE_HIT_RECORDS[3:4].copy(), width=width, height=height, bounces=bounces)


def duplicate_key_scene(n_tris: int = 20_000, clusters: int = 37, seed: int = 11, width: int = 64, height: int = 64) -> Scene:
    """Build stress case: triangles piled onto a few cluster centres, so that long runs of primitives share one Morton
    code (the index-augmented part of the radix-tree definition decides the topology there), plus two far outliers that
    stretch the quantisation grid."""
    rng = np.random.default_rng(seed)
    centres = (rng.random((clusters, 3), dtype=np.float32) * np.float32(2) - np.float32(1)).astype(np.float32)
    which = rng.integers(0, clusters, size=n_tris)
    c = centres[which]
    e1 = (rng.random((n_tris, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(1e-5)
    e2 = (rng.random((n_tris, 3), dtype=np.float32) - np.float32(0.5)) * np.float32(1e-5)
    verts = np.empty((n_tris, 3, 3), dtype=np.float32)
    verts[:, 0] = c; verts[:, 1] = c + e1; verts[:, 2] = c + e2
    verts[0] += np.float32(50.0); verts[1] -= np.float32(50.0)
    geo = Geometry(verts.reshape(-1, 3), None, None)
    inst = Instance(IDENTITY_3X4.copy(), 1, 0xFF, 0, INSTANCE_TRIANGLE_FACING_CULL_DISABLE, 0)
    return Scene(f"dupkeys{n_tris}", [[geo]], [inst], SAMPLE_HIT_RECORDS[:1].copy(), width=width, height=height, bounces=0)


def random_scene(n_blas: int, tris_per_blas: int, n_instances: int, seed: int, width: int = 256, height: int = 160,
                 bounces: int = 1, n_geoms: int = 2, shared_edges: bool = True) -> Scene:
    """Small fuzz scene for brute-force parity: random triangle clusters (optionally as a connected strip
    so that many rays cross shared edges), several geometries per BLAS with transforms, random
    rotated/scaled instances, masks and SBT offsets."""
    rng = np.random.default_rng(seed)
    blases: List[List[Geometry]] = []
    for b in range(n_blas):
        geoms = []
        for g in range(n_geoms):
            nt = max(1, tris_per_blas // n_geoms)
            if shared_edges:
                side = max(1, int(np.sqrt(nt / 2)))
                geo = heightfield(side, side, -1.0, 1.0, -1.0, 1.0, 0.7, seed * 1000 + b * 10 + g, base_freq=2.0)
            else:
                c = rng.uniform(-1, 1, size=(nt, 1, 3))
                v = (c + rng.uniform(-0.25, 0.25, size=(nt, 3, 3))).astype(np.float32)
                geo = Geometry(np.ascontiguousarray(v.reshape(-1, 3)), None, None)
            if g > 0:
                geo.transform = rotation_3x4(rng.normal(size=3), float(rng.uniform(-1, 1)), rng.uniform(-0.5, 0.5, size=3))
            geoms.append(geo)
        blases.append(geoms)
    n_records = 4 + n_geoms
    records = rng.uniform(0, 1, size=(n_records, 3)).astype(np.float32)
    instances = []
    for i in range(n_instances):
        t = np.array([rng.uniform(-4, 4), rng.uniform(-2.5, 2.5), rng.uniform(-3, 3)])
        m = rotation_3x4(rng.normal(size=3), float(rng.uniform(-3, 3)), t, scale=float(rng.uniform(0.5, 1.5)))
        mask = 0xFF if i % 7 != 6 else 0x00          # every 7th instance is invisible to cullMask 0xff
        instances.append(Instance(m, 100 if i == 1 else i, mask, int(rng.integers(0, 4)), 1, int(rng.integers(0, n_blas))))
    return Scene(f"random{seed}", blases, instances, records, width=width, height=height, bounces=bounces)


def save_scene(scene: Scene, path: str) -> None:
    """Writes the scene's host arrays as an "RTSCENE1" file, the input format of the C++ multi-GPU host program
    (host/sample_scene_mgpu.cpp --scene): header {n_blas, n_instances, n_records, width, height, bounces} (uint32),
    {camera xyz, fov, miss rgb} (float32); per BLAS: n_geoms, then per geometry {n_verts, n_indexed_tris, has_xform, flags}, the
    vertices (float32 x 3), the indices (uint32 x 3, absent for a non-indexed list), the 3x4 transform (if any); the 64-byte
    instance records (VkAccelerationStructureInstanceKHR layout, main.cpp:848-858, BLAS index in the reference field); hit records."""
    import struct
    with open(path, "wb") as f:
        f.write(b"RTSCENE1")
        f.write(struct.pack("<6I", len(scene.blases), len(scene.instances), int(scene.hit_records.shape[0]), scene.width, scene.height, scene.bounces))
        f.write(struct.pack("<7f", *[float(x) for x in scene.camera_pos], float(scene.yfov_deg), *[float(x) for x in scene.miss_color]))
        for geoms in scene.blases:
            f.write(struct.pack("<I", len(geoms)))
            for g in geoms:
                v = np.ascontiguousarray(g.vertices, dtype=np.float32).reshape(-1, 3)
                nt = 0 if g.indices is None else int(g.indices.shape[0])
                f.write(struct.pack("<4I", v.shape[0], nt, 0 if g.transform is None else 1, int(getattr(g, "flags", 1)) & 0xFF))
                f.write(v.tobytes())
                if g.indices is not None:
                    f.write(np.ascontiguousarray(g.indices, dtype=np.uint32).tobytes())
                if g.transform is not None:
                    f.write(np.ascontiguousarray(g.transform, dtype=np.float32).tobytes())
        for I in scene.instances:
            f.write(np.ascontiguousarray(I.transform, dtype=np.float32).tobytes())
            f.write(struct.pack("<IIQ", (I.custom_index & 0xFFFFFF) | ((I.mask & 0xFF) << 24), (I.sbt_offset & 0xFFFFFF) | ((I.flags & 0xFF) << 24), int(I.blas)))
        f.write(np.ascontiguousarray(scene.hit_records, dtype=np.float32).tobytes())
