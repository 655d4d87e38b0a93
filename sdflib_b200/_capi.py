"""ctypes binding of the C-ABI in include/sdfb200.h. This is the ONLY way the Python host reaches the
compute path; if libsdfb200.so is missing or no CUDA device is present the calls raise — there is no
Python/CPU fallback."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsdfb200.so")

OK = 0
ERR_INVALID, ERR_CUDA, ERR_IO, ERR_UNSUPPORTED = -1, -2, -3, -4
FORMAT_OCTREE, FORMAT_EXACT_OCTREE = 1, 2
QUERY_DEVICE_POINTERS, QUERY_EXACT_ORDER = 1, 2


class SdfB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__(f"sdfb200 error {code}: {message}")
        self.code = code


class Info(C.Structure):
    _fields_ = [("format", C.c_int32), ("box_min", C.c_float * 3), ("box_max", C.c_float * 3),
                ("start_grid_size", C.c_int32), ("max_depth", C.c_uint32), ("value_range", C.c_float),
                ("min_border_value", C.c_float), ("start_depth", C.c_uint32), ("min_triangles_in_leafs", C.c_uint32),
                ("max_triangles_in_leafs", C.c_uint32), ("max_triangles_encoded_in_leafs", C.c_uint32),
                ("bit_encoding_start_depth", C.c_uint32), ("bits_per_index", C.c_uint32), ("octree_words", C.c_uint64),
                ("triangle_sets_words", C.c_uint64), ("triangle_masks_bytes", C.c_uint64), ("num_triangles", C.c_uint64),
                ("device", C.c_int32)]


class BuildStats(C.Structure):
    _fields_ = [(n, C.c_double) for n in ("total_ms", "triangle_data_ms", "bvh_ms", "upload_ms", "levels_ms", "layout_ms",
                                          "download_ms")] + \
               [(n, C.c_uint64) for n in ("nodes_processed", "leaves", "samples_evaluated", "kernel_launches")]


# every symbol include/sdfb200.h declares (tests check that the library exports all of them)
SYMBOLS = ["sdfb200_last_error", "sdfb200_version", "sdfb200_device_count", "sdfb200_set_device", "sdfb200_release_cached_memory", "sdfb200_build_octree",
           "sdfb200_build_exact", "sdfb200_build_octree_shard", "sdfb200_build_octree_collective", "sdfb200_build_exact_shard", "sdfb200_shard_sizes",
           "sdfb200_shard_finish", "sdfb200_shard_words", "sdfb200_shard_export",
           "sdfb200_assemble", "sdfb200_save", "sdfb200_load", "sdfb200_free", "sdfb200_get_info",
           "sdfb200_get_build_stats", "sdfb200_get_octree_data", "sdfb200_get_exact_arrays", "sdfb200_get_device_octree",
           "sdfb200_query", "sdfb200_triangle_data", "sdfb200_nearest_triangle", "sdfb200_point_triangle",
           "sdfb200_make_isosphere", "sdfb200_mesh_create", "sdfb200_mesh_free", "sdfb200_mesh_blob_bytes", "sdfb200_mesh_export",
           "sdfb200_mesh_import", "sdfb200_mesh_stats", "sdfb200_build_octree_from_mesh", "sdfb200_build_octree_collective_from_mesh",
           "sdfb200_build_exact_from_mesh", "sdfb200_build_octree_multi", "sdfb200_build_exact_multi", "sdfb200_nccl_available",
           "sdfb200_sphere_trace", "sdfb200_mesh_bvh", "sdfb200_bvh_host"]

# int (*sdfb200_allgather_fn)(void* user, const void* dSend, void* dRecv, uint64_t bytesPerRank)
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64)

_lib = None


def lib():
    """Loads libsdfb200.so; raises if it has not been built (python -m sdflib_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdfB200Error(ERR_CUDA, f"{LIB_PATH} is missing: build it with `python -m sdflib_b200.build` "
                                         "(the CUDA extension is required, there is no fallback)")
        L = C.CDLL(LIB_PATH)
        L.sdfb200_last_error.restype = C.c_char_p
        L.sdfb200_free.restype = None
        L.sdfb200_mesh_free.restype = None
        _lib = L
    return _lib


def check(code):
    if code != OK:
        raise SdfB200Error(code, lib().sdfb200_last_error().decode(errors="replace"))


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)
