"""ctypes mirror of the C ABI in include/rtcore.h (librtcore.so, CUDA sm_100a).

The calls map 1:1 onto the reference sample's setup steps (vulkan-raytracing-basic/main.cpp):
createBLAS -> Context.build_blas, createTLAS -> Context.build_tlas, createUniformBuffer /
createShaderBindingTable -> camera argument / Context.set_hit_records, render -> Context.trace.

There is no CPU fallback: loading fails loudly if librtcore.so is missing, and Context() raises if
no CUDA device is usable.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional, Sequence

import numpy as np

from . import build as _build

_HERE = os.path.dirname(os.path.abspath(__file__))

RT_SUCCESS = 0
RT_ERROR_INVALID_ARG, RT_ERROR_CUDA, RT_ERROR_OUT_OF_MEMORY = -1, -2, -3
RT_ERROR_STACK_DEPTH, RT_ERROR_SBT_RANGE, RT_ERROR_INTERNAL = -4, -5, -6
RT_GEOMETRY_OPAQUE = 0x1
RT_GEOMETRY_DEVICE_POINTERS = 0x100
RT_BUILD_ALLOW_UPDATE = 0x1
RT_BUILD_ALLOW_COMPACTION = 0x2
RT_BUILD_PREFER_FAST_TRACE = 0x4
RT_BUILD_PREFER_FAST_BUILD = 0x8
RT_BUILD_MODE_REFIT = 0x1000
RT_BUILD_INSTANCES_ON_DEVICE = 0x100
RT_BUILD_NO_PACKED_SORT = 0x200
RT_TRACE_OUT_DEVICE = 0x1
RT_TRACE_STATS = 0x2
RT_TRACE_ASYNC = 0x4
RT_TRACE_OUT_FULL_FRAME = 0x8
RT_TRACE_OUT_BGRA = 0x10
RT_REF_EMPTY = 0x7FFFFFFD

EXPORTED_SYMBOLS = [
    "rt_create", "rt_destroy", "rt_last_error", "rt_device_info", "rt_set_stream", "rt_sync", "rt_release_scratch",
    "rt_blas_build_sizes", "rt_tlas_build_sizes", "rt_build_blas", "rt_build_blas_batch", "rt_build_tlas",
    "rt_update_tlas", "rt_update_blas", "rt_compact_blas", "rt_free_blas", "rt_free_tlas", "rt_last_build_timing", "rt_last_build_ms",
    "rt_last_build_scratch_bytes", "rt_tlas_storage_bytes", "rt_blas_device_reference",
    "rt_blas_get_info", "rt_blas_export", "rt_debug_last_sorted_keys", "rt_blas_import", "rt_tlas_get_info",
    "rt_set_hit_records", "rt_set_miss_color", "rt_set_miss_records", "rt_set_anyhit_records", "rt_set_ray_params", "rt_trace", "rt_trace_rows", "rt_trace_rows_range",
    "rt_rows_packed_pixels", "rt_unpack_rows", "rt_frame_share_create", "rt_frame_share_open", "rt_frame_share_close", "rt_frame_share_free", "rt_flag_add", "rt_flag_wait_ge", "rt_last_trace_stats", "rt_last_trace_ms",
    "rt_kernel_launch_count", "rt_version", "rt_copy_to_host",
    "rt_group_create", "rt_group_destroy", "rt_group_trace", "rt_group_flush_host", "rt_host_frame_wait", "rt_group_sync", "rt_group_join", "rt_group_barrier", "rt_group_rank", "rt_group_world",
    "rt_group_share_blas", "rt_group_share_finish", "rt_group_last_share_ms", "rt_group_last_share_host_ms", "rt_group_host_frame_begin", "rt_group_host_frame_end", "rt_group_last_error",
    # include/rtcore_io.h
    "rt_obj_load", "rt_obj_parse", "rt_obj_free", "rt_obj_last_error", "rt_obj_vertex_count", "rt_obj_triangle_count",
    "rt_obj_group_count", "rt_obj_vertices", "rt_obj_indices", "rt_obj_group_name", "rt_obj_group_first_triangle",
    "rt_obj_group_triangle_count", "rt_obj_geometry", "rt_write_ppm", "rt_write_png", "rt_srgb8_table",
]


ANYHIT_ACCEPT, ANYHIT_ALPHA_MASK, ANYHIT_TERMINATE_RAY = 0, 1, 1


class RtAnyHitRecord(C.Structure):
    _fields_ = [("kind", C.c_uint32), ("log2_res", C.c_uint32), ("flags", C.c_uint32), ("reserved", C.c_uint32), ("mask", C.c_void_p)]


class RtGeometry(C.Structure):
    _fields_ = [("vertices", C.c_void_p), ("vertex_count", C.c_uint32), ("vertex_stride_bytes", C.c_uint32),
                ("indices", C.c_void_p), ("triangle_count", C.c_uint32), ("transform3x4", C.c_void_p),
                ("flags", C.c_uint32)]


class RtInstance(C.Structure):
    # bit-fields of rt_instance packed by hand: custom_index:24 | mask:8, sbt_offset:24 | flags:8
    _fields_ = [("transform", C.c_float * 12), ("custom_index_and_mask", C.c_uint32),
                ("sbt_offset_and_flags", C.c_uint32), ("blas", C.c_void_p)]


assert C.sizeof(RtInstance) == 64


class RtCamera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("yfov_deg", C.c_float)]


class RtRayParams(C.Structure):
    _fields_ = [("tmin", C.c_float), ("tmax", C.c_float), ("cull_mask", C.c_uint32),
                ("sbt_record_offset", C.c_uint32), ("sbt_record_stride", C.c_uint32), ("bounce_seed", C.c_uint32),
                ("ray_flags", C.c_uint32), ("miss_index", C.c_uint32)]


# RT_RAY_FLAG_* / RT_INSTANCE_* (include/rtcore.h)
RAY_FLAG_OPAQUE, RAY_FLAG_NO_OPAQUE, RAY_FLAG_TERMINATE_ON_FIRST_HIT, RAY_FLAG_SKIP_CLOSEST_HIT_SHADER = 0x01, 0x02, 0x04, 0x08
RAY_FLAG_CULL_BACK_FACING, RAY_FLAG_CULL_FRONT_FACING, RAY_FLAG_CULL_OPAQUE, RAY_FLAG_CULL_NO_OPAQUE = 0x10, 0x20, 0x40, 0x80
INSTANCE_FACING_CULL_DISABLE, INSTANCE_FLIP_FACING, INSTANCE_FORCE_OPAQUE, INSTANCE_FORCE_NO_OPAQUE = 0x1, 0x2, 0x4, 0x8


class RtBuildSizes(C.Structure):
    _fields_ = [("acceleration_structure_size", C.c_uint64), ("build_scratch_size", C.c_uint64)]


class RtTraceStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays_primary", "rays_secondary", "nodes_visited", "triangles_tested",
                                          "instances_entered", "primary_hits", "secondary_hits", "near_edge_hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class RtBuildTiming(C.Structure):
    _fields_ = [("total_ms", C.c_float), ("setup_ms", C.c_float), ("morton_ms", C.c_float), ("sort_ms", C.c_float),
                ("hierarchy_ms", C.c_float), ("refit_ms", C.c_float), ("h2d_ms", C.c_float), ("primitives", C.c_uint64)]

    def as_dict(self):
        return {n: (float(getattr(self, n)) if n != "primitives" else int(self.primitives)) for n, _ in self._fields_}


class RtBlasInfo(C.Structure):
    _fields_ = [("triangle_count", C.c_uint32), ("node_count", C.c_uint32), ("root_ref", C.c_int32),
                ("max_depth", C.c_uint32), ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3),
                ("storage_bytes", C.c_uint64), ("device_storage", C.c_void_p)]


class RtTlasInfo(C.Structure):
    _fields_ = [("instance_count", C.c_uint32), ("node_count", C.c_uint32), ("root_ref", C.c_int32),
                ("max_depth", C.c_uint32), ("bounds_lo", C.c_float * 3), ("bounds_hi", C.c_float * 3)]


HIT_DTYPE = np.dtype([("instance_id", "<u4"), ("geometry_index", "<u4"), ("primitive_id", "<u4"),
                      ("custom_index", "<u4"), ("t", "<f4"), ("u", "<f4"), ("v", "<f4")])
assert HIT_DTYPE.itemsize == 28


class RtError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"rtcore error {code}: {msg}")
        self.code = code


_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Loads librtcore.so (building it with nvcc first when absent/stale). Raises if that fails."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("RTCORE_LIB") or lib_path()      # RTCORE_LIB: tuning variants built by build.build_variant()
    if build_if_missing and not os.environ.get("RTCORE_LIB") and _build.needs_build():
        _build.build_rtcore()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: the CUDA extension must be built (python -m build_up_phase_b200.build); there is no CPU fallback")
    L = C.CDLL(path)
    vp, u32, i32, u64 = C.c_void_p, C.c_uint32, C.c_int, C.c_uint64
    L.rt_create.argtypes = [i32, C.POINTER(vp)]
    L.rt_destroy.argtypes = [vp]
    L.rt_destroy.restype = None
    L.rt_last_error.argtypes = [vp]
    L.rt_last_error.restype = C.c_char_p
    L.rt_device_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_size_t)]
    L.rt_set_stream.argtypes = [vp, vp]
    L.rt_sync.argtypes = [vp]
    L.rt_release_scratch.argtypes = [vp]
    L.rt_blas_build_sizes.argtypes = [vp, C.POINTER(u32), u32, C.POINTER(RtBuildSizes)]
    L.rt_tlas_build_sizes.argtypes = [vp, u32, C.POINTER(RtBuildSizes)]
    L.rt_build_blas.argtypes = [vp, C.POINTER(RtGeometry), u32, u32, C.POINTER(vp)]
    L.rt_build_blas_batch.argtypes = [vp, C.POINTER(RtGeometry), C.POINTER(u32), u32, u32, C.POINTER(vp)]
    L.rt_build_tlas.argtypes = [vp, vp, u32, u32, C.POINTER(vp)]
    L.rt_update_tlas.argtypes = [vp, vp, vp, u32, u32]
    L.rt_update_blas.argtypes = [vp, vp, C.POINTER(RtGeometry), u32, u32]
    L.rt_compact_blas.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(u64)]
    L.rt_free_blas.argtypes = [vp, vp]
    L.rt_free_blas.restype = None
    L.rt_free_tlas.argtypes = [vp, vp]
    L.rt_free_tlas.restype = None
    L.rt_last_build_timing.argtypes = [vp, C.POINTER(RtBuildTiming)]
    L.rt_last_build_ms.argtypes = [vp]
    L.rt_last_build_ms.restype = C.c_float
    L.rt_last_build_scratch_bytes.argtypes = [vp]
    L.rt_last_build_scratch_bytes.restype = u64
    L.rt_tlas_storage_bytes.argtypes = [vp, vp]
    L.rt_tlas_storage_bytes.restype = u64
    L.rt_blas_device_reference.argtypes = [vp, vp]
    L.rt_blas_device_reference.restype = u64
    L.rt_blas_get_info.argtypes = [vp, vp, C.POINTER(RtBlasInfo)]
    L.rt_blas_export.argtypes = [vp, vp, vp, vp]
    L.rt_debug_last_sorted_keys.argtypes = [vp, vp, vp, u32, C.POINTER(u32)]
    L.rt_blas_import.argtypes = [vp, C.POINTER(RtBlasInfo), vp, C.POINTER(vp)]
    L.rt_tlas_get_info.argtypes = [vp, vp, C.POINTER(RtTlasInfo)]
    L.rt_set_hit_records.argtypes = [vp, vp, u32]
    L.rt_set_miss_color.argtypes = [vp, C.POINTER(C.c_float)]
    L.rt_set_miss_records.argtypes = [vp, vp, u32]
    L.rt_set_ray_params.argtypes = [vp, C.POINTER(RtRayParams)]
    L.rt_set_anyhit_records.argtypes = [vp, vp, u32]
    L.rt_trace.argtypes = [vp, vp, C.POINTER(RtCamera), u32, u32, u32, u32, vp, vp, vp]
    L.rt_trace_rows.argtypes = [vp, vp, C.POINTER(RtCamera), u32, u32, u32, u32, u32, u32, u32, vp, vp, vp]
    L.rt_trace_rows_range.argtypes = [vp, vp, C.POINTER(RtCamera), u32, u32, u32, u32, u32, u32, u32, u32, u32, vp, vp, vp]
    L.rt_rows_packed_pixels.argtypes = [u32, u32, u32, u32]
    L.rt_rows_packed_pixels.restype = u64
    L.rt_unpack_rows.argtypes = [vp, vp, u32, u32, u32, u32, vp]
    L.rt_frame_share_create.argtypes = [vp, u64, C.POINTER(vp), vp]
    L.rt_frame_share_open.argtypes = [vp, vp, C.POINTER(vp)]
    L.rt_frame_share_close.argtypes = [vp, vp]
    L.rt_frame_share_free.argtypes = [vp, vp]
    L.rt_flag_add.argtypes = [vp, vp]
    L.rt_flag_wait_ge.argtypes = [vp, vp, u32]
    L.rt_last_trace_stats.argtypes = [vp, C.POINTER(RtTraceStats)]
    L.rt_last_trace_ms.argtypes = [vp]
    L.rt_last_trace_ms.restype = C.c_float
    L.rt_kernel_launch_count.argtypes = [vp]
    L.rt_kernel_launch_count.restype = u64
    L.rt_version.restype = C.c_char_p
    L.rt_copy_to_host.argtypes = [vp, vp, vp, u64]
    L.rt_group_create.argtypes = [vp, C.c_char_p, i32, i32, u32, u32, C.POINTER(vp)]
    L.rt_group_destroy.argtypes = [vp]
    L.rt_group_destroy.restype = None
    L.rt_group_trace.argtypes = [vp, vp, C.POINTER(RtCamera), u32, u32, u32, u32, C.POINTER(vp)]
    L.rt_group_flush_host.argtypes = [vp, C.POINTER(vp)]
    L.rt_host_frame_wait.argtypes = [vp]
    for name in ("rt_group_sync", "rt_group_join", "rt_group_barrier", "rt_group_rank", "rt_group_world", "rt_group_share_finish"):
        getattr(L, name).argtypes = [vp]
    L.rt_group_share_blas.argtypes = [vp, u32, i32, vp, C.POINTER(vp)]
    L.rt_group_last_share_host_ms.argtypes = [vp, C.POINTER(C.c_float)]
    L.rt_group_last_share_ms.argtypes = [vp]
    L.rt_group_last_share_ms.restype = C.c_float
    L.rt_group_host_frame_begin.argtypes = [vp, C.POINTER(vp)]
    L.rt_group_host_frame_end.argtypes = [vp, C.POINTER(vp)]
    L.rt_group_last_error.argtypes = [vp]
    L.rt_group_last_error.restype = C.c_char_p
    # include/rtcore_io.h
    L.rt_obj_load.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.rt_obj_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(vp)]
    L.rt_obj_free.argtypes = [vp]
    L.rt_obj_free.restype = None
    L.rt_obj_last_error.restype = C.c_char_p
    for name in ("rt_obj_vertex_count", "rt_obj_triangle_count", "rt_obj_group_count"):
        getattr(L, name).argtypes = [vp]
        getattr(L, name).restype = u32
    L.rt_obj_vertices.argtypes = [vp]
    L.rt_obj_vertices.restype = C.POINTER(C.c_float)
    L.rt_obj_indices.argtypes = [vp]
    L.rt_obj_indices.restype = C.POINTER(u32)
    L.rt_obj_group_name.argtypes = [vp, u32]
    L.rt_obj_group_name.restype = C.c_char_p
    L.rt_obj_group_first_triangle.argtypes = [vp, u32]
    L.rt_obj_group_first_triangle.restype = u32
    L.rt_obj_group_triangle_count.argtypes = [vp, u32]
    L.rt_obj_group_triangle_count.restype = u32
    L.rt_obj_geometry.argtypes = [vp, u32, C.POINTER(RtGeometry)]
    L.rt_write_ppm.argtypes = [C.c_char_p, vp, u32, u32, u32]
    L.rt_write_png.argtypes = [C.c_char_p, vp, u32, u32, u32]
    L.rt_srgb8_table.argtypes = [vp]
    L.rt_srgb8_table.restype = None
    _lib = L
    return L


def _ptr(x) -> Optional[int]:
    """numpy array -> host pointer; torch tensor / int -> raw pointer."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    raise TypeError(type(x))


class _CudaPtr:
    """Minimal __cuda_array_interface__ holder: lets torch alias raw device memory owned by librtcore."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}


def device_view(ptr: int, nbytes: int, device):
    """uint8 torch tensor aliasing [ptr, ptr+nbytes) on `device` (no copy) — e.g. a BLAS blob handed to NCCL."""
    import torch
    return torch.as_tensor(_CudaPtr(ptr, nbytes), device=device)


class Blas:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.handle = ctx, handle

    def info(self) -> RtBlasInfo:
        info = RtBlasInfo()
        self.ctx._check(self.ctx.L.rt_blas_get_info(self.ctx.h, self.handle, C.byref(info)))
        return info

    def device_reference(self) -> int:
        """accelerationStructureReference: what a device-resident rt_instance array carries in its blas field."""
        return int(self.ctx.L.rt_blas_device_reference(self.ctx.h, self.handle))

    def export(self):
        """(nodes uint32[n,16], tris uint32[n,12]) raw 64-B nodes and 48-B triangles."""
        info = self.info()
        nodes = np.zeros((info.node_count, 16), dtype=np.uint32)
        tris = np.zeros((info.triangle_count, 12), dtype=np.uint32)
        self.ctx._check(self.ctx.L.rt_blas_export(self.ctx.h, self.handle, nodes.ctypes.data, tris.ctypes.data))
        return nodes, tris

    def free(self):
        if self.handle:
            self.ctx.L.rt_free_blas(self.ctx.h, self.handle)
            self.handle = None


class Tlas:
    def __init__(self, ctx: "Context", handle: int):
        self.ctx, self.handle = ctx, handle

    def info(self) -> RtTlasInfo:
        info = RtTlasInfo()
        self.ctx._check(self.ctx.L.rt_tlas_get_info(self.ctx.h, self.handle, C.byref(info)))
        return info

    def storage_bytes(self) -> int:
        return int(self.ctx.L.rt_tlas_storage_bytes(self.ctx.h, self.handle))

    def free(self):
        if self.handle:
            self.ctx.L.rt_free_tlas(self.ctx.h, self.handle)
            self.handle = None


class Context:
    """rt_context: one CUDA device, one stream."""

    def __init__(self, device: int = 0):
        self.L = load()
        h = C.c_void_p()
        rc = self.L.rt_create(device, C.byref(h))
        if rc != RT_SUCCESS:
            raise RtError(rc, "rt_create failed (no usable CUDA device? there is no CPU fallback)")
        self.h = h
        self.device = device
        self._keep: list = []

    # -- plumbing -------------------------------------------------------------------------------
    def _check(self, rc: int):
        if rc != RT_SUCCESS:
            raise RtError(rc, self.L.rt_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.L.rt_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def device_info(self):
        sm, ma, mi, mem = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._check(self.L.rt_device_info(self.h, C.byref(sm), C.byref(ma), C.byref(mi), C.byref(mem)))
        return {"sm_count": sm.value, "cc": (ma.value, mi.value), "total_mem": mem.value}

    def set_stream(self, cuda_stream: int):
        self._check(self.L.rt_set_stream(self.h, cuda_stream))

    def sync(self):
        self._check(self.L.rt_sync(self.h))

    def release_scratch(self):
        self._check(self.L.rt_release_scratch(self.h))

    def launch_count(self) -> int:
        return int(self.L.rt_kernel_launch_count(self.h))

    # -- acceleration structures --------------------------------------------------------------------
    def _geom_array(self, geoms, keep: list, device: bool = False):
        arr = (RtGeometry * max(1, len(geoms)))()
        for i, g in enumerate(geoms):
            if device:
                v, idx, t = g.vertices, g.indices, g.transform
                arr[i].vertices = _ptr(v)
                arr[i].vertex_count = int(v.shape[0])
                arr[i].indices = _ptr(idx)
                arr[i].triangle_count = int(idx.shape[0]) if idx is not None else int(v.shape[0]) // 3
                arr[i].transform3x4 = _ptr(t)
                arr[i].flags = (getattr(g, "flags", RT_GEOMETRY_OPAQUE) & 0xFF) | RT_GEOMETRY_DEVICE_POINTERS
                keep.extend([v, idx, t])
            else:
                v = np.ascontiguousarray(g.vertices, dtype=np.float32)
                keep.append(v)
                arr[i].vertices = v.ctypes.data
                arr[i].vertex_count = v.shape[0]
                if g.indices is not None:
                    idx = np.ascontiguousarray(g.indices, dtype=np.uint32)
                    keep.append(idx)
                    arr[i].indices = idx.ctypes.data
                arr[i].triangle_count = g.triangle_count
                if g.transform is not None:
                    t = np.ascontiguousarray(g.transform, dtype=np.float32)
                    keep.append(t)
                    arr[i].transform3x4 = t.ctypes.data
                arr[i].flags = getattr(g, "flags", RT_GEOMETRY_OPAQUE) & 0xFF
            vv = g.vertices
            arr[i].vertex_stride_bytes = 4 * int(vv.shape[1]) if getattr(vv, "ndim", 0) == 2 else 12   # [nv, k >= 3]: x y z first, k - 3 floats of padding
        return arr

    def blas_build_sizes(self, max_triangle_counts: Sequence[int]) -> RtBuildSizes:
        arr = (C.c_uint32 * len(max_triangle_counts))(*max_triangle_counts)
        out = RtBuildSizes()
        self._check(self.L.rt_blas_build_sizes(self.h, arr, len(max_triangle_counts), C.byref(out)))
        return out

    def tlas_build_sizes(self, max_instances: int) -> RtBuildSizes:
        out = RtBuildSizes()
        self._check(self.L.rt_tlas_build_sizes(self.h, max_instances, C.byref(out)))
        return out

    def build_blas(self, geoms, device: bool = False, flags: int = 0) -> Blas:
        keep: list = []
        arr = self._geom_array(geoms, keep, device)
        h = C.c_void_p()
        self._check(self.L.rt_build_blas(self.h, arr, len(geoms), RT_BUILD_PREFER_FAST_TRACE | flags, C.byref(h)))
        return Blas(self, h.value)

    def update_blas(self, blas: Blas, geoms, device: bool = False, flags: int = 0):
        """rt_update_blas: same counts, new vertex data; the handle (and what TLAS instances point at) stays valid."""
        keep: list = []
        arr = self._geom_array(geoms, keep, device)
        self._check(self.L.rt_update_blas(self.h, blas.handle, arr, len(geoms), RT_BUILD_PREFER_FAST_TRACE | flags))

    def compact_blas(self, blas: Blas):
        """rt_compact_blas: packs the live nodes of the BLAS (or of the whole batch it was built in) into a right-sized allocation.
        Returns (bytes_before, bytes_after). TLASes referencing it must be rebuilt."""
        b0, b1 = C.c_uint64(0), C.c_uint64(0)
        self._check(self.L.rt_compact_blas(self.h, blas.handle, C.byref(b0), C.byref(b1)))
        return int(b0.value), int(b1.value)

    def build_blas_batch(self, blases, device: bool = False, flags: int = 0) -> List[Blas]:
        keep: list = []
        flat = [g for b in blases for g in b]
        arr = self._geom_array(flat, keep, device)
        counts = (C.c_uint32 * len(blases))(*[len(b) for b in blases])
        out = (C.c_void_p * len(blases))()
        self._check(self.L.rt_build_blas_batch(self.h, arr, counts, len(blases), RT_BUILD_PREFER_FAST_TRACE | flags, out))
        return [Blas(self, out[i]) for i in range(len(blases))]

    def import_blas(self, info: RtBlasInfo, device_blob) -> Blas:
        """rt_blas_import: adopt (copy) a relocatable BLAS blob that another GPU built and broadcast."""
        h = C.c_void_p()
        self._check(self.L.rt_blas_import(self.h, C.byref(info), _ptr(device_blob), C.byref(h)))
        return Blas(self, h.value)

    @staticmethod
    def instance_array(instances, blas_handles: Sequence[Blas]):
        arr = (RtInstance * max(1, len(instances)))()
        for i, I in enumerate(instances):
            for k in range(12):
                arr[i].transform[k] = float(I.transform[k])
            arr[i].custom_index_and_mask = (I.custom_index & 0xFFFFFF) | ((I.mask & 0xFF) << 24)
            arr[i].sbt_offset_and_flags = (I.sbt_offset & 0xFFFFFF) | ((I.flags & 0xFF) << 24)
            arr[i].blas = blas_handles[I.blas].handle if I.blas is not None and I.blas >= 0 else None
        return arr

    def build_tlas(self, instances, blas_handles: Sequence[Blas]) -> Tlas:
        arr = self.instance_array(instances, blas_handles)
        h = C.c_void_p()
        self._check(self.L.rt_build_tlas(self.h, C.addressof(arr), len(instances), RT_BUILD_PREFER_FAST_TRACE, C.byref(h)))
        return Tlas(self, h.value)

    def build_tlas_device(self, instance_records_dev, n_instances: int) -> Tlas:
        """rt_build_tlas with RT_BUILD_INSTANCES_ON_DEVICE: `instance_records_dev` is a device buffer of n 64-byte rt_instance
        records whose blas fields hold Blas.device_reference() values (the reference's instance buffer, main.cpp:860-868)."""
        h = C.c_void_p()
        self._check(self.L.rt_build_tlas(self.h, _ptr(instance_records_dev), n_instances,
                                         RT_BUILD_PREFER_FAST_TRACE | RT_BUILD_INSTANCES_ON_DEVICE, C.byref(h)))
        return Tlas(self, h.value)

    def build_scratch_bytes(self) -> int:
        return int(self.L.rt_last_build_scratch_bytes(self.h))

    def update_tlas(self, tlas: Tlas, instances, blas_handles: Sequence[Blas]):
        arr = self.instance_array(instances, blas_handles)
        self._check(self.L.rt_update_tlas(self.h, tlas.handle, C.addressof(arr), len(instances), RT_BUILD_PREFER_FAST_TRACE))

    def build_timing(self) -> dict:
        t = RtBuildTiming()
        self._check(self.L.rt_last_build_timing(self.h, C.byref(t)))
        return t.as_dict()

    def last_sorted_keys(self):
        n = C.c_uint32()
        self._check(self.L.rt_debug_last_sorted_keys(self.h, None, None, 0, C.byref(n)))
        keys = np.zeros(n.value, dtype=np.uint64)
        prims = np.zeros(n.value, dtype=np.uint32)
        self._check(self.L.rt_debug_last_sorted_keys(self.h, keys.ctypes.data, prims.ctypes.data, n.value, C.byref(n)))
        return keys, prims

    # -- shader data ----------------------------------------------------------------------------------
    def set_hit_records(self, rgb: np.ndarray):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1, 3)
        self._check(self.L.rt_set_hit_records(self.h, rgb.ctypes.data, rgb.shape[0]))

    def set_miss_color(self, rgb):
        arr = (C.c_float * 3)(*[float(x) for x in rgb])
        self._check(self.L.rt_set_miss_color(self.h, arr))

    def set_miss_records(self, rgb: np.ndarray):
        rgb = np.ascontiguousarray(rgb, dtype=np.float32).reshape(-1, 3)
        self._check(self.L.rt_set_miss_records(self.h, rgb.ctypes.data, rgb.shape[0]))

    def set_anyhit_records(self, records):
        """Any-hit records of the hit groups: list of (kind, log2_res, flags, mask words as a uint32 array or None); [] removes the table."""
        masks = [None if m is None else np.ascontiguousarray(m, dtype=np.uint32) for _, _, _, m in records]
        arr = (RtAnyHitRecord * max(1, len(records)))()
        for i, (kind, log2_res, flags, _) in enumerate(records):
            arr[i].kind, arr[i].log2_res, arr[i].flags = kind, log2_res, flags
            arr[i].mask = None if masks[i] is None else masks[i].ctypes.data
        self._check(self.L.rt_set_anyhit_records(self.h, C.cast(arr, C.c_void_p) if records else None, len(records)))

    def set_ray_params(self, tmin=0.0, tmax=100.0, cull_mask=0xFF, sbt_record_offset=0, sbt_record_stride=1, bounce_seed=1,
                       ray_flags=RAY_FLAG_OPAQUE, miss_index=0):
        p = RtRayParams(tmin, tmax, cull_mask, sbt_record_offset, sbt_record_stride, bounce_seed, ray_flags, miss_index)
        self._check(self.L.rt_set_ray_params(self.h, C.byref(p)))

    # -- dispatch -----------------------------------------------------------------------------------------
    @staticmethod
    def camera(pos, yfov_deg) -> RtCamera:
        cam = RtCamera()
        for k in range(3):
            cam.pos[k] = float(pos[k])
        cam.yfov_deg = float(yfov_deg)
        return cam

    def trace(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int = 0, want_hits: bool = False,
              stats: bool = False, rgba_out: Optional[np.ndarray] = None, bgra: bool = False):
        """Host-buffer trace (the reference-facing call): returns (rgba[h,w,4], primary hits, secondary hits)."""
        rgba = rgba_out if rgba_out is not None else np.empty((height, width, 4), dtype=np.uint8)
        prim = np.empty((height, width), dtype=HIT_DTYPE) if want_hits else None
        sec = np.empty((height, width), dtype=HIT_DTYPE) if want_hits else None
        flags = (RT_TRACE_STATS if stats else 0) | (RT_TRACE_OUT_BGRA if bgra else 0)
        self._check(self.L.rt_trace(self.h, tlas.handle, C.byref(cam), width, height, bounces, flags, _ptr(rgba), _ptr(prim), _ptr(sec)))
        return rgba, prim, sec

    def trace_device(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int, rgba_dev, prim_dev=None,
                     sec_dev=None, stats: bool = False, async_: bool = False):
        flags = RT_TRACE_OUT_DEVICE | (RT_TRACE_STATS if stats else 0) | (RT_TRACE_ASYNC if async_ else 0)
        self._check(self.L.rt_trace(self.h, tlas.handle, C.byref(cam), width, height, bounces, flags, _ptr(rgba_dev), _ptr(prim_dev), _ptr(sec_dev)))

    def frame_share_create(self, nbytes: int):
        """-> (device pointer, 64-byte IPC handle as bytes) of a framebuffer other ranks can map (rt_frame_share_create)."""
        p = C.c_void_p()
        h = (C.c_uint8 * 64)()
        self._check(self.L.rt_frame_share_create(self.h, nbytes, C.byref(p), h))
        return p.value, bytes(h)

    def frame_share_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        h = (C.c_uint8 * 64).from_buffer_copy(handle)
        self._check(self.L.rt_frame_share_open(self.h, h, C.byref(p)))
        return p.value

    def frame_share_close(self, ptr: int):
        self._check(self.L.rt_frame_share_close(self.h, ptr))

    def frame_share_free(self, ptr: int):
        self._check(self.L.rt_frame_share_free(self.h, ptr))

    def flag_add(self, counter_ptr: int):
        self._check(self.L.rt_flag_add(self.h, counter_ptr))

    def flag_wait_ge(self, counter_ptr: int, target: int):
        self._check(self.L.rt_flag_wait_ge(self.h, counter_ptr, target & 0xFFFFFFFF))

    def trace_rows(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int, block_rows: int, part_index: int,
                   part_count: int, rgba, prim=None, sec=None, device: bool = False, stats: bool = False,
                   async_: bool = False, full_frame: bool = False):
        flags = (RT_TRACE_OUT_DEVICE if device else 0) | (RT_TRACE_STATS if stats else 0) | (RT_TRACE_ASYNC if async_ else 0) | \
                (RT_TRACE_OUT_FULL_FRAME if full_frame else 0)
        self._check(self.L.rt_trace_rows(self.h, tlas.handle, C.byref(cam), width, height, bounces, flags, block_rows, part_index,
                                         part_count, _ptr(rgba), _ptr(prim), _ptr(sec)))

    def trace_rows_range(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int, block_rows: int, part_index: int,
                         part_count: int, first_row: int, n_rows: int, rgba, full_frame: bool = False, async_: bool = True):
        """rt_trace_rows_range: only the packed rows [first_row, first_row + n_rows) of this part (device output)."""
        flags = RT_TRACE_OUT_DEVICE | (RT_TRACE_ASYNC if async_ else 0) | (RT_TRACE_OUT_FULL_FRAME if full_frame else 0)
        self._check(self.L.rt_trace_rows_range(self.h, tlas.handle, C.byref(cam), width, height, bounces, flags, block_rows, part_index,
                                               part_count, first_row, n_rows, _ptr(rgba), None, None))

    def rows_packed_pixels(self, width: int, height: int, block_rows: int, part_count: int) -> int:
        return int(self.L.rt_rows_packed_pixels(width, height, block_rows, part_count))

    def unpack_rows(self, packed_all_dev, width: int, height: int, block_rows: int, part_count: int, rgba_out_dev):
        self._check(self.L.rt_unpack_rows(self.h, _ptr(packed_all_dev), width, height, block_rows, part_count, _ptr(rgba_out_dev)))

    def trace_stats(self) -> dict:
        s = RtTraceStats()
        self._check(self.L.rt_last_trace_stats(self.h, C.byref(s)))
        return s.as_dict()

    def trace_ms(self) -> float:
        return float(self.L.rt_last_trace_ms(self.h))


GROUP_OUT_DEVICE, GROUP_OUT_HOST, GROUP_ASYNC, GROUP_PIPELINE = 0x1, 0x2, 0x4, 0x8


class Group:
    """rt_group: `world` processes (one per GPU of one box, scene replicated) rendering one frame together. The ranks meet in a
    POSIX shared-memory block named after `name`; no torch.distributed / NCCL is involved. ctx=None gives a host-only group
    (barrier + shared host frame), which is what the CPU tests of the handshake use."""

    def __init__(self, ctx: Optional["Context"], name: str, rank: int, world: int, max_width: int, max_height: int):
        self.L = load()
        self.ctx = ctx
        h = C.c_void_p()
        rc = self.L.rt_group_create(ctx.h if ctx is not None else None, name.encode(), rank, world, max_width, max_height, C.byref(h))
        if rc != RT_SUCCESS:
            raise RtError(rc, self.L.rt_last_error(ctx.h).decode() if ctx is not None else "rt_group_create failed")
        self.h, self.rank, self.world = h, rank, world
        self.max_width, self.max_height = max_width, max_height

    def _check(self, rc: int):
        if rc != RT_SUCCESS:
            raise RtError(rc, self.L.rt_group_last_error(self.h).decode())

    def trace(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int, flags: int) -> Optional[int]:
        """One group frame; returns the frame pointer on rank 0 (device or host, by flags), None elsewhere."""
        p = C.c_void_p()
        self._check(self.L.rt_group_trace(self.h, tlas.handle, C.byref(cam), width, height, bounces, flags, C.byref(p)))
        return p.value

    def trace_host(self, tlas: Tlas, cam: RtCamera, width: int, height: int, bounces: int, pipeline: bool = False) -> Optional[np.ndarray]:
        """RT_GROUP_OUT_HOST: every rank copies its bands over its own PCIe link; rank 0 gets a numpy VIEW of the shared pinned frame.
        pipeline=True (RT_GROUP_PIPELINE): two frames in flight; the view is the PREVIOUS frame (None for the first call), flush_host() the last."""
        p = self.trace(tlas, cam, width, height, bounces, GROUP_OUT_HOST | (GROUP_PIPELINE if pipeline else 0))
        if not p:
            return None
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(height, width, 4))

    def flush_host(self, width: int, height: int) -> Optional[np.ndarray]:
        """rt_group_flush_host: completes the pending frame of a pipelined host-output sequence; rank 0 gets its view."""
        p = C.c_void_p()
        self._check(self.L.rt_group_flush_host(self.h, C.byref(p)))
        if not p.value:
            return None
        return np.ctypeslib.as_array(C.cast(p.value, C.POINTER(C.c_uint8)), shape=(height, width, 4))

    def host_frame_begin(self) -> int:
        p = C.c_void_p()
        self._check(self.L.rt_group_host_frame_begin(self.h, C.byref(p)))
        return p.value

    def host_frame_end(self) -> Optional[int]:
        p = C.c_void_p()
        self._check(self.L.rt_group_host_frame_end(self.h, C.byref(p)))
        return p.value

    def sync(self):
        self._check(self.L.rt_group_sync(self.h))

    def join(self):
        self._check(self.L.rt_group_join(self.h))

    def barrier(self):
        self._check(self.L.rt_group_barrier(self.h))

    def share_blas(self, slot: int, owner_rank: int, mine: Optional[Blas]) -> Blas:
        h = C.c_void_p()
        self._check(self.L.rt_group_share_blas(self.h, slot, owner_rank, mine.handle if mine is not None else None, C.byref(h)))
        return mine if (mine is not None and self.rank == owner_rank) else Blas(self.ctx, h.value)

    def share_finish(self) -> float:
        self._check(self.L.rt_group_share_finish(self.h))
        return float(self.L.rt_group_last_share_ms(self.h))

    def share_host_ms(self) -> dict:
        a = (C.c_float * 3)()
        self._check(self.L.rt_group_last_share_host_ms(self.h, a))
        return {"wait_for_owner_ms": float(a[0]), "ipc_open_ms": float(a[1]), "alloc_ms": float(a[2])}

    def close(self):
        if getattr(self, "h", None):
            self.L.rt_group_destroy(self.h)
            self.h = None


class SceneHandles:
    """Builds a scenes.Scene on a Context the way the sample's main() does: BLASes (batched when there is
    more than one), TLAS, hit records, miss colour."""

    def __init__(self, ctx: Context, scene, batch: bool = True, build_flags: int = 0):
        self.ctx, self.scene = ctx, scene
        if len(scene.blases) > 1 and batch:
            self.blases = ctx.build_blas_batch(scene.blases, flags=build_flags)
        else:
            self.blases = [ctx.build_blas(g, flags=build_flags) for g in scene.blases]
        self.blas_timing = ctx.build_timing()
        self.tlas = ctx.build_tlas(scene.instances, self.blases)
        self.tlas_timing = ctx.build_timing()
        ctx.set_hit_records(scene.hit_records)
        ctx.set_miss_color(scene.miss_color)
        self.cam = ctx.camera(scene.camera_pos, scene.yfov_deg)

    def trace(self, width=None, height=None, bounces=None, want_hits=True, stats=False):
        s = self.scene
        return self.ctx.trace(self.tlas, self.cam, width or s.width, height or s.height,
                              s.bounces if bounces is None else bounces, want_hits=want_hits, stats=stats)

    def rebuild_tlas(self):
        """After rt_compact_blas (device addresses changed): a new TLAS over the same instances."""
        self.tlas.free()
        self.tlas = self.ctx.build_tlas(self.scene.instances, self.blases)

    def free(self):
        self.tlas.free()
        for b in self.blases:
            b.free()


# ---- include/rtcore_io.h: host-side formats either side of the path -------------------------------------------------
IMAGE_SRGB_ENCODE = 0x1
IMAGE_FLIP_Y = 0x2


class ObjMesh:
    """rt_obj_mesh: a Wavefront .obj parsed by the library (the reference's "load an obj file -> build the acceleration
    structure" assignment, vulkan-raytracing-basic/README.md:225-226). groups() -> [(name, first_triangle, count)]."""

    def __init__(self, path: Optional[str] = None, text: Optional[bytes] = None):
        L = load()
        h = C.c_void_p()
        if path is not None:
            rc = L.rt_obj_load(os.fsencode(path), C.byref(h))
        else:
            rc = L.rt_obj_parse(text, len(text), C.byref(h))
        if rc != 0:
            raise RtError(rc, L.rt_obj_last_error().decode())
        self._h = h
        nv, nt = L.rt_obj_vertex_count(h), L.rt_obj_triangle_count(h)
        self.vertices = (np.ctypeslib.as_array(L.rt_obj_vertices(h), shape=(nv, 3)).copy() if nv else np.zeros((0, 3), np.float32))
        self.indices = (np.ctypeslib.as_array(L.rt_obj_indices(h), shape=(nt, 3)).copy() if nt else np.zeros((0, 3), np.uint32))
        self._groups = [(L.rt_obj_group_name(h, g).decode(), L.rt_obj_group_first_triangle(h, g), L.rt_obj_group_triangle_count(h, g))
                        for g in range(L.rt_obj_group_count(h))]

    def groups(self):
        return list(self._groups)

    def geometries(self):
        """One scenes.Geometry per group (shared vertex array, the group's slice of the index array)."""
        from . import scenes
        return [scenes.Geometry(self.vertices, self.indices[f:f + n].copy(), None) for _, f, n in self._groups]

    def c_geometry(self, group: int) -> RtGeometry:
        g = RtGeometry()
        rc = load().rt_obj_geometry(self._h, group, C.byref(g))
        if rc != 0:
            raise RtError(rc, "rt_obj_geometry")
        return g

    def free(self):
        if self._h:
            load().rt_obj_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def write_ppm(path: str, rgba: np.ndarray, flags: int = 0) -> None:
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = a.shape[0], a.shape[1]
    rc = load().rt_write_ppm(os.fsencode(path), a.ctypes.data, w, h, flags)
    if rc != 0:
        raise RtError(rc, load().rt_obj_last_error().decode())


def write_png(path: str, rgba: np.ndarray, flags: int = 0) -> None:
    a = np.ascontiguousarray(rgba, dtype=np.uint8)
    h, w = a.shape[0], a.shape[1]
    rc = load().rt_write_png(os.fsencode(path), a.ctypes.data, w, h, flags)
    if rc != 0:
        raise RtError(rc, load().rt_obj_last_error().decode())


def srgb8_table() -> np.ndarray:
    t = np.zeros(256, dtype=np.uint8)
    load().rt_srgb8_table(t.ctypes.data)
    return t
