"""ctypes loader for libcuvite_b200.so (the C ABI of include/cuvite_b200.h).

There is no fallback: if the CUDA library is missing or will not load, importing
the compute entry points raises.  The library is built in-tree by
``cuvite_b200.build`` (nvcc, sm_100a) and is never replaced by a CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcuvite_b200.so")

_vp, _i64, _i32, _f64 = C.c_void_p, C.c_int64, C.c_int, C.c_double


class VitePart(C.Structure):
    _fields_ = [("nv_global", _i64), ("ne_global", _i64), ("base", _i64), ("nv_local", _i64),
                ("first_entry", _i64), ("ne_local", _i64)]


class GraphInfo(C.Structure):
    _fields_ = [("nv_local", _i64), ("base", _i64), ("nv_global", _i64), ("ne_local", _i64),
                ("weight_mode", C.c_int32), ("num_classes", C.c_int32),
                ("class_count", _i64 * 24), ("class_entries", _i64 * 24),
                ("max_degree", _i64), ("local_weight", _f64)]


# name -> (restype, argtypes); also the list the "not gpu" test checks against include/cuvite_b200.h
SIGNATURES = {
    "cvb_last_error": (C.c_char_p, []),
    "cvb_version": (C.c_char_p, []),
    "cvb_device_count": (_i32, []),
    "cvb_set_device": (_i32, [_i32]),
    "cvb_create": (_vp, [_vp]),
    "cvb_destroy": (None, [_vp]),
    "cvb_graph_upload_host": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "cvb_graph_upload_device": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp]),
    "cvb_graph_get_info": (_i32, [_vp, C.POINTER(GraphInfo)]),
    "cvb_phase_begin": (_i32, [_vp, _f64]),
    "cvb_sweep": (_i32, [_vp]),
    "cvb_apply_and_reduce": (_i32, [_vp, C.POINTER(_f64)]),
    "cvb_rebuild_and_reduce": (_i32, [_vp, C.POINTER(_f64)]),
    "cvb_rebuild_and_reduce_async": (_i32, [_vp]),
    "cvb_modularity_from_sums": (_f64, [_vp, C.POINTER(_f64)]),
    "cvb_rotate_labels": (_i32, [_vp]),
    "cvb_phase_run": (_i32, [_vp, _f64, _f64, _i32, C.POINTER(_f64), C.POINTER(_i32)]),
    "cvb_get_labels": (_i32, [_vp, _i32, _vp]),
    "cvb_get_cluster_weight": (_i32, [_vp, _vp]),
    "cvb_get_cinfo": (_i32, [_vp, _vp, _vp]),
    "cvb_phase_reset": (_i32, [_vp]),
    "cvb_buffer": (_vp, [_vp, _i32]),
    "cvb_stream": (_vp, [_vp]),
    "cvb_sync": (_i32, [_vp]),
    "cvb_louvain_iteration": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _vp,
                                     _i64, _vp, _vp, _vp, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "cvb_reduce_mod": (_i32, [_vp, _i64, _vp, _vp, C.POINTER(_f64)]),
    "cvb_coarsen": (_i32, [_vp, C.POINTER(_i64)]),
    "cvb_renumber": (_i32, [_vp, _vp, _i64, _vp, C.POINTER(_i64)]),
    "cvb_contract_local": (_i32, [_vp, _vp, _i64, C.POINTER(_i64)]),
    "cvb_graph_from_pairs": (_i32, [_vp, _i64, _i64, _i64, _vp, _vp, _i64]),
    "cvb_graph_download": (_i32, [_vp, _vp, _vp, _vp]),
    "cvb_vite_partition": (_i32, [C.c_char_p, _i32, _i32, _i32, C.POINTER(VitePart), _vp]),
    "cvb_vite_read_rows": (_i32, [C.c_char_p, C.POINTER(VitePart), _vp, _vp]),
    "cvb_graph_load_vite": (_i32, [_vp, C.c_char_p, _i32, _i32, _i32, C.POINTER(VitePart)]),
    "cvb_louvain_run": (_i32, [_vp, _i32, _vp, _vp, _vp, _vp]),
    "cvb_louvain_run_ex": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cvb_phase_run_et": (_i32, [_vp, _f64, _f64, _i32, _f64, C.POINTER(_f64), C.POINTER(_i32)]),
    "cvb_louvain_run_et": (_i32, [_vp, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp]),
    "cvb_phase_run_color": (_i32, [_vp, _f64, _f64, _i32, C.POINTER(_f64), C.POINTER(_i32), C.POINTER(_i32)]),
    "cvb_louvain_run_color": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "cvb_phase_run_color_et": (_i32, [_vp, _f64, _f64, _i32, _i32, _f64, C.POINTER(_f64), C.POINTER(_i32),
                                      C.POINTER(_i32)]),
    "cvb_louvain_run_color_et": (_i32, [_vp, _i32, _i32, _i32, _i32, _f64, _vp, _vp, _vp, _vp]),
    "cvb_color_graph": (_i32, [_vp, _i32, C.POINTER(_i32)]),
    "cvb_get_colors": (_i32, [_vp, _vp]),
    "cvb_dist_unique_id": (_i32, [_vp, _i32]),
    "cvb_dist_init": (_i32, [_vp, _vp, _i32, _i32, _i32]),
    "cvb_dist_finalize": (_i32, [_vp]),
    "cvb_dist_set_parts": (_i32, [_vp, _vp]),
    "cvb_dist_agree": (_i32, [_vp]),
    "cvb_dist_phase_begin": (_i32, [_vp]),
    "cvb_dist_sweep": (_i32, [_vp, C.POINTER(_f64)]),
    "cvb_dist_advance": (_i32, [_vp]),
    "cvb_dist_phase_run": (_i32, [_vp, _f64, _f64, _i32, C.POINTER(_f64), C.POINTER(_i32)]),
    "cvb_dist_stage_ms": (_i32, [_vp, C.POINTER(_f64)]),
    "cvb_launch_count": (_i64, [_vp]),
    "cvb_last_sweep_ms": (C.c_float, [_vp]),
}

_lib = None


def load() -> C.CDLL:
    """Load the CUDA library, binding every symbol of the C ABI.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run `python -m cuvite_b200.build`); "
            "cuvite_b200 has no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class CuviteError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().cvb_last_error().decode(errors="replace")
        raise CuviteError(f"{what or 'cuvite_b200'} failed ({rc}): {msg}")
