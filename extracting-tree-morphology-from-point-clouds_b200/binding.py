"""ctypes binding of include/treemorph_nn.h — the only bridge between Python and the CUDA library.

Plain integers (``tensor.data_ptr()``) cross the boundary; no torch types.  There is no CPU
fallback: if the library is missing or no CUDA device is present, the calls raise.
"""
from __future__ import annotations

import ctypes
import os

from . import build as _build

TM_OK, TM_ERR_INVALID, TM_ERR_NO_CYLINDERS, TM_ERR_CUDA, TM_ERR_NOMEM, TM_ERR_STATE = range(6)
TM_MODE_AUTO, TM_MODE_BRUTE, TM_MODE_GRID = 0, 1, 2
TM_F32, TM_F64 = 0, 1
ABI_VERSION = 7
TM_PHASES = 9
TM_COMM_ID_BYTES = 128
PHASE_NAMES = ("bin", "scan", "scatter", "evaluate", "tree", "exhaustive", "pending", "epilogue", "total")

c_i64, c_i32, c_f32, c_vp = ctypes.c_int64, ctypes.c_int32, ctypes.c_float, ctypes.c_void_p


class TmParams(ctypes.Structure):
    _fields_ = [("perp_atol", c_f32), ("norm_eps", c_f32), ("move_to_mantle", c_i32), ("norm_fma", c_i32),
                ("mode", c_i32), ("cell_size", c_f32), ("reserved", c_i32 * 2)]


class TmStats(ctypes.Structure):
    _fields_ = [("pairs_evaluated", ctypes.c_uint64), ("cull_tests", ctypes.c_uint64),
                ("points_grid", ctypes.c_uint64), ("points_far", ctypes.c_uint64), ("points_ring", ctypes.c_uint64), ("points_tree", ctypes.c_uint64),
                ("points_brute", ctypes.c_uint64), ("index_entries", ctypes.c_uint64),
                ("voxels_occupied", ctypes.c_uint32), ("work_items", ctypes.c_uint32),
                ("mode_used", ctypes.c_uint32), ("cell_size", c_f32), ("reach", c_f32), ("near_reach", c_f32),
                ("grid_dim", ctypes.c_uint32 * 3), ("launches", ctypes.c_uint32), ("bound_tests", ctypes.c_uint64),
                ("points_slow", ctypes.c_uint64), ("lane_ops_per_bound", ctypes.c_uint32), ("reserved", ctypes.c_uint32)]

    def as_dict(self) -> dict:
        d = {name: getattr(self, name) for name, _ in self._fields_ if name != "grid_dim"}
        d["grid_dim"] = list(self.grid_dim)
        return d


# every symbol include/treemorph_nn.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "tm_version": (ctypes.c_int, []),
    "tm_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(c_vp)]),
    "tm_destroy": (ctypes.c_int, [c_vp]),
    "tm_last_error": (ctypes.c_char_p, [c_vp]),
    "tm_status_string": (ctypes.c_char_p, [ctypes.c_int]),
    "tm_prepare_cylinders": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_i64, c_f32, c_i32,
                                            c_vp, c_vp, c_vp]),
    "tm_set_cylinders": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64,
                                        c_vp, c_i64, c_i64, c_vp]),
    "tm_cylinder_count": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i64)]),
    "tm_comm_unique_id": (ctypes.c_int, [c_vp]),
    "tm_comm_init_rank": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i32]),
    "tm_comm_init_all": (ctypes.c_int, [ctypes.POINTER(c_vp), c_i32]),
    "tm_comm_destroy": (ctypes.c_int, [c_vp]),
    "tm_comm_info": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "tm_broadcast_cylinders": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64,
                                              c_vp, c_i64, c_i64, c_i32, c_vp]),
    "tm_broadcast_cylinders_all": (ctypes.c_int, [ctypes.POINTER(c_vp), c_i32, c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64,
                                                  c_vp, c_i64, c_vp, c_i64, c_i64, c_i32]),
    "tm_label_points": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.POINTER(TmParams), c_vp, c_vp, c_vp, c_vp,
                                       c_vp, c_vp]),
    "tm_assemble_records": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
    "tm_label_cloud_host": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i64, c_i64, ctypes.POINTER(TmParams), c_vp, c_vp]),
    "tm_label_cloud_host_wide": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i64, c_i64, ctypes.POINTER(TmParams), c_vp, c_i32,
                                                ctypes.POINTER(ctypes.c_double), c_vp]),
    "tm_cloud_upload_host": (ctypes.c_int, [c_vp, c_vp, c_i32, c_i64, c_i64]),
    "tm_proximity_flags_host": (ctypes.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, ctypes.POINTER(TmParams), c_f32,
                                               c_f32, c_vp, c_vp, c_vp]),
    "tm_knn_covariance": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, c_i32, c_vp, c_vp, c_vp]),
    "tm_radius_count": (ctypes.c_int, [c_vp, c_vp, c_i64, c_i64, ctypes.c_double, c_vp, c_vp]),
    "tm_noise_cloud": (ctypes.c_int, [c_vp, c_vp, c_vp, c_i64, c_i64, c_i64, ctypes.c_uint64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
    "tm_host_pipeline_info": (ctypes.c_int, [c_vp, ctypes.POINTER(c_i32), ctypes.POINTER(c_i32)]),
    "tm_measure_host_bandwidth": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(c_i32)]),
    "tm_get_stats": (ctypes.c_int, [c_vp, ctypes.POINTER(TmStats)]),
    "tm_set_profiling": (ctypes.c_int, [c_vp, ctypes.c_int]),
    "tm_get_phase_ms": (ctypes.c_int, [c_vp, ctypes.POINTER(c_f32)]),
    "tm_measure_fp32_peak": (ctypes.c_int, [c_vp, ctypes.POINTER(ctypes.c_double)]),
    "tm_selftest_arithmetic": (ctypes.c_int, [c_vp, ctypes.c_uint64, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint64)]),
}

_lib = None


class TreemorphError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"treemorph_nn status {status}: {message}")
        self.status = status


def library_path() -> str:
    return _build.LIB_PATH


def load(build_if_missing: bool = True):
    """dlopen the C-ABI library and attach the prototypes.  Fails loudly if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if build_if_missing and _build.find_nvcc() is not None and _build.is_stale():
        _build.build()
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing and cannot be built here (no nvcc). "
                           "There is no CPU fallback for the nearest-cylinder path.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError here == a symbol the header promises is not exported
        fn.restype = res
        fn.argtypes = args
    if lib.tm_version() != ABI_VERSION:
        raise RuntimeError(f"libtreemorph_nn ABI {lib.tm_version()} != binding ABI {ABI_VERSION}")
    _lib = lib
    return lib


def check(lib, handle, status: int) -> None:
    if status == TM_OK:
        return
    msg = lib.tm_last_error(handle) if handle else lib.tm_status_string(status)
    text = msg.decode(errors="replace") if msg else ""
    if not text:
        text = lib.tm_status_string(status).decode()
    if status == TM_ERR_NO_CYLINDERS:
        raise IndexError(text)          # the reference raises IndexError from argmin on an empty dim
    if status == TM_ERR_NOMEM:
        raise MemoryError(text)
    if status == TM_ERR_INVALID:
        raise ValueError(text)
    raise TreemorphError(status, text)
