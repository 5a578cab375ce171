"""ctypes binding of the CPU ORACLE (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("rt_oracle.cpp", "rt_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _LIB_PATH


class OrcGeometry(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_count", C.c_uint32), ("vertex_stride_bytes", C.c_uint32),
                ("indices", C.c_void_p), ("triangle_count", C.c_uint32), ("transform3x4", C.c_void_p),
                ("flags", C.c_uint32)]


class OrcInstance(C.Structure):
    _fields_ = [("transform", C.c_float * 12), ("custom_index_and_mask", C.c_uint32),
                ("sbt_offset_and_flags", C.c_uint32), ("blas", C.c_void_p)]


class OrcCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("yfov_deg", C.c_float)]


class OrcRayParams(C.Structure):
    _fields_ = [("tmin", C.c_float), ("tmax", C.c_float), ("cull_mask", C.c_uint32),
                ("sbt_record_offset", C.c_uint32), ("sbt_record_stride", C.c_uint32), ("bounce_seed", C.c_uint32),
                ("ray_flags", C.c_uint32), ("miss_index", C.c_uint32)]


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_secondary", "nodes_visited", "triangles_tested",
                                          "instances_entered", "primary_hits", "secondary_hits", "near_edge_hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class OrcAnyHitRecord(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("log2_res", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32), ("mask", C.c_void_p)]


class OrcShaderData(C.Structure):
    _fields_ = [("hit_records_rgb", C.c_void_p), ("hit_record_count", C.c_uint32), ("miss_rgb", C.c_float * 3),
                ("miss_records_rgb", C.c_void_p), ("miss_record_count", C.c_uint32),
                ("anyhit_records", C.c_void_p), ("anyhit_record_count", C.c_uint32)]


class OrcBlasInfo(C.Structure):
    _fields_ = [("triangle_count", C.c_uint32), ("node_count", C.c_uint32), ("root_ref", C.c_int32),
                ("max_depth", C.c_uint32), ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3)]


HIT_DTYPE = np.dtype([("instance_id", "<u4"), ("geometry_index", "<u4"), ("primitive_id", "<u4"),
                      ("custom_index", "<u4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
assert HIT_DTYPE.itemsize == 28

MODE_BRUTE, MODE_BVH = 0, 1

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_build_blas.restype = C.c_void_p
        L.orc_build_blas.argtypes = [C.POINTER(OrcGeometry), C.c_uint32, C.c_int, C.c_int]
        L.orc_free_blas.argtypes = [C.c_void_p]
        L.orc_refit_blas.restype = C.c_int
        L.orc_refit_blas.argtypes = [C.c_void_p, C.POINTER(OrcGeometry), C.c_uint32]
        L.orc_build_tlas.restype = C.c_void_p
        L.orc_build_tlas.argtypes = [C.POINTER(OrcInstance), C.c_uint32, C.c_int]
        L.orc_free_tlas.argtypes = [C.c_void_p]
        L.orc_trace.restype = C.c_int
        L.orc_trace.argtypes = [C.c_void_p, C.POINTER(OrcCamera), C.POINTER(OrcRayParams), C.POINTER(OrcShaderData),
                                C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32,
                                C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(OrcStats)]
        L.orc_blas_get_info.argtypes = [C.c_void_p, C.POINTER(OrcBlasInfo)]
        L.orc_blas_export.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_time_blas_build.restype = C.c_double
        L.orc_time_blas_build.argtypes = [C.POINTER(OrcGeometry), C.c_uint32, C.c_int]
        L.orc_aspect_y.restype = C.c_float
        L.orc_aspect_y.argtypes = [C.c_float]
        L.orc_pcg_hash.restype = C.c_uint32
        L.orc_pcg_hash.argtypes = [C.c_uint32]
        L.orc_morton30.restype = C.c_uint32
        L.orc_morton30.argtypes = [C.c_float, C.c_float, C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
        L.orc_unorm8.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_uint8)]
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _geom_array(geoms, keep: list):
    arr = (OrcGeometry * len(geoms))()
    for i, g in enumerate(geoms):
        v = np.ascontiguousarray(g.vertices, dtype=np.float32)
        keep.append(v)
        arr[i].vertices = v.ctypes.data
        arr[i].vertex_count = v.shape[0]
        arr[i].vertex_stride_bytes = 4 * int(v.shape[1]) if v.ndim == 2 else 12      # [nv, k >= 3]: x y z first, k - 3 floats of padding
        if g.indices is not None:
            idx = np.ascontiguousarray(g.indices, dtype=np.uint32)
            keep.append(idx)
            arr[i].indices = idx.ctypes.data
        else:
            arr[i].indices = None
        arr[i].triangle_count = g.triangle_count
        if g.transform is not None:
            t = np.ascontiguousarray(g.transform, dtype=np.float32)
            keep.append(t)
            arr[i].transform3x4 = t.ctypes.data
        else:
            arr[i].transform3x4 = None
        arr[i].flags = getattr(g, "flags", 1) & 0xFF
    return arr


class OracleScene:
    """Builds the oracle's BLAS/TLAS for a scenes.Scene and traces it on the CPU."""

    def __init__(self, scene, build_bvh: bool = True):
        self.scene = scene
        self._keep: list = []
        L = lib()
        self.blases = []
        for geoms in scene.blases:
            arr = _geom_array(geoms, self._keep)
            self.blases.append(L.orc_build_blas(arr, len(geoms), 1 if build_bvh else 0, 30))
        inst = (OrcInstance * max(1, len(scene.instances)))()
        for i, I in enumerate(scene.instances):
            for k in range(12):
                inst[i].transform[k] = float(I.transform[k])
            inst[i].custom_index_and_mask = (I.custom_index & 0xFFFFFF) | ((I.mask & 0xFF) << 24)
            inst[i].sbt_offset_and_flags = (I.sbt_offset & 0xFFFFFF) | ((I.flags & 0xFF) << 24)
            inst[i].blas = self.blases[I.blas]
        self.tlas = L.orc_build_tlas(inst, len(scene.instances), 1 if build_bvh else 0)
        self.records = np.ascontiguousarray(scene.hit_records, dtype=np.float32)
        self.sd = OrcShaderData()
        self.sd.hit_records_rgb = self.records.ctypes.data
        self.sd.hit_record_count = self.records.shape[0]
        for k in range(3):
            self.sd.miss_rgb[k] = float(scene.miss_color[k])
        self.sd.miss_records_rgb = None
        self.sd.miss_record_count = 0
        self.sd.anyhit_records = None
        self.sd.anyhit_record_count = 0

    def refit_blas(self, i: int, geoms):
        """orc_refit_blas: new vertices for BLAS i, topology of its last full build."""
        arr = _geom_array(geoms, self._keep)
        rc = lib().orc_refit_blas(self.blases[i], arr, len(geoms))
        if rc != 0:
            raise RuntimeError(f"orc_refit_blas failed: {rc}")

    def set_miss_records(self, rgb):
        self.miss_records = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1, 3)
        self.sd.miss_records_rgb = self.miss_records.ctypes.data
        self.sd.miss_record_count = self.miss_records.shape[0]

    def set_anyhit_records(self, records):
        """records: list of (kind, log2_res, flags, mask words as a uint32 array or None); [] removes the table."""
        self._anyhit_masks = [None if m is None else np.ascontiguousarray(m, dtype=np.uint32) for _, _, _, m in records]
        self._anyhit = (OrcAnyHitRecord * max(1, len(records)))()
        for i, (kind, log2_res, flags, _) in enumerate(records):
            self._anyhit[i].kind, self._anyhit[i].log2_res, self._anyhit[i].flags = kind, log2_res, flags
            self._anyhit[i].mask = None if self._anyhit_masks[i] is None else self._anyhit_masks[i].ctypes.data
        self.sd.anyhit_records = C.cast(self._anyhit, C.c_void_p) if records else None
        self.sd.anyhit_record_count = len(records)

    def trace(self, width: Optional[int] = None, height: Optional[int] = None, bounces: Optional[int] = None,
              mode: int = MODE_BVH, rows=(0, None, 1), ray_params: Optional[OrcRayParams] = None,
              want_hits: bool = True):
        s = self.scene
        w = width or s.width
        h = height or s.height
        b = s.bounces if bounces is None else bounces
        r0, r1, rs = rows
        r1 = h if r1 is None else r1
        cam = OrcCamera()
        for k in range(3):
            cam.pos[k] = float(s.camera_pos[k])
        cam.yfov_deg = float(s.yfov_deg)
        rgba = np.zeros((h, w, 4), dtype=np.uint8)
        prim = np.zeros((h, w), dtype=HIT_DTYPE) if want_hits else None
        sec = np.zeros((h, w), dtype=HIT_DTYPE) if want_hits else None
        stats = OrcStats()
        rc = lib().orc_trace(self.tlas, C.byref(cam), C.byref(ray_params) if ray_params is not None else None,
                             C.byref(self.sd), w, h, b, mode, r0, r1, rs, rgba.ctypes.data,
                             prim.ctypes.data if prim is not None else None,
                             sec.ctypes.data if sec is not None else None, C.byref(stats))
        if rc != 0:
            raise RuntimeError(f"orc_trace failed: {rc}")
        return rgba, prim, sec, stats.as_dict()

    def blas_info(self, i: int = 0) -> OrcBlasInfo:
        info = OrcBlasInfo()
        lib().orc_blas_get_info(self.blases[i], C.byref(info))
        return info

    def blas_export(self, i: int = 0):
        info = self.blas_info(i)
        nodes = np.zeros((info.node_count, 16), dtype=np.uint32)
        tris = np.zeros((info.triangle_count, 12), dtype=np.uint32)
        keys = np.zeros(info.triangle_count, dtype=np.uint64)
        prims = np.zeros(info.triangle_count, dtype=np.uint32)
        rc = lib().orc_blas_export(self.blases[i], nodes.ctypes.data, tris.ctypes.data, keys.ctypes.data, prims.ctypes.data)
        if rc != 0:
            raise RuntimeError("orc_blas_export failed")
        return info, nodes, tris, keys, prims

    def close(self):
        L = lib()
        if self.tlas:
            L.orc_free_tlas(self.tlas)
            self.tlas = None
        for b in self.blases:
            L.orc_free_blas(b)
        self.blases = []

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def time_blas_build(geoms) -> float:
    keep: list = []
    arr = _geom_array(geoms, keep)
    return float(lib().orc_time_blas_build(arr, len(geoms), 30))


def ray_params(tmin=0.0, tmax=100.0, cull_mask=0xFF, sbt_record_offset=0, sbt_record_stride=1, bounce_seed=1,
               ray_flags=0x1, miss_index=0) -> OrcRayParams:
    return OrcRayParams(tmin, tmax, cull_mask, sbt_record_offset, sbt_record_stride, bounce_seed, ray_flags, miss_index)


def num_threads() -> int:
    return int(lib().orc_num_threads())
