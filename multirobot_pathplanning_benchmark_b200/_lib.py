"""ctypes binding of libmrb200.so (the C ABI declared in include/mrb200.h).

There is no CPU fallback: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# (MRB200_LIB: an experiment build of the same library, scripts/build_variant.sh; never a different implementation)
LIB_PATH = os.environ.get("MRB200_LIB") or os.path.join(HERE, "libmrb200.so")

_lib = None

c_f32p, c_f64p = C.POINTER(C.c_float), C.POINTER(C.c_double)
c_u8p, c_i32p, c_vp = C.POINTER(C.c_uint8), C.POINTER(C.c_int32), C.c_void_p

# name -> (restype, argtypes); mirrors include/mrb200.h one to one
SIGNATURES = {
    "mrb200_version": (C.c_int, []),
    "mrb200_last_error": (C.c_char_p, []),
    "mrb200_launch_count": (C.c_int64, []),
    "mrb200_fp32_probe": (C.c_int, [C.c_int, c_vp, c_i32p, c_vp]),
    "mrb200_abstract_create": (C.c_int, [C.c_int, C.c_int, c_f64p, C.c_int, c_f64p, C.c_int, c_f64p, C.POINTER(c_vp)]),
    "mrb200_abstract_destroy": (C.c_int, [c_vp]),
    "mrb200_abstract_check_configs": (C.c_int, [c_vp, c_vp, C.c_int64, c_vp, c_vp]),
    "mrb200_abstract_check_edges": (C.c_int, [c_vp, c_vp, c_vp, C.c_int64, C.c_double, c_vp, C.c_int32, C.c_int32,
                                              C.c_int, c_vp, c_vp, c_vp]),
    "mrb200_abstract_query_configs_host": (C.c_int, [c_vp, c_vp, C.c_int64, c_vp, c_vp]),
    "mrb200_abstract_query_edges_host": (C.c_int, [c_vp, c_vp, c_vp, C.c_int64, C.c_double, c_vp, C.c_int32, C.c_int32, C.c_int,
                                                   c_vp, c_vp, c_vp]),
    "mrb200_scene_create": (C.c_int, [C.c_int, C.POINTER(c_vp)]),
    "mrb200_scene_destroy": (C.c_int, [c_vp]),
    "mrb200_scene_set_mode": (C.c_int, [c_vp, C.c_int, c_vp, C.c_size_t, c_vp]),
    "mrb200_check_configs": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, C.c_float, c_vp, c_vp, C.c_int, c_vp]),
    "mrb200_check_configs_for_robot": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, C.c_float, c_u8p, c_u8p, C.c_int,
                                                 c_vp, c_vp]),
    "mrb200_check_edges": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, C.c_int64, C.c_double, c_vp, C.c_int32, C.c_int32,
                                     C.c_int, C.c_float, c_vp, c_vp, c_vp]),
    "mrb200_query_configs_host": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, C.c_float, c_vp, c_vp, C.c_int, c_vp, c_vp]),
    "mrb200_query_edges_host": (C.c_int, [c_vp, C.c_int, c_vp, c_vp, C.c_int64, C.c_double, c_vp, C.c_int32, C.c_int32, C.c_int,
                                          C.c_float, c_vp, c_vp, c_vp]),
    "mrb200_check_configs_host": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, C.c_float, c_vp, C.c_int64, c_vp]),
    "mrb200_submit_edges_host": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int, c_vp, C.c_int64, C.c_double, c_vp, C.c_int, C.c_float,
                                           C.POINTER(C.c_int64), c_vp]),
    "mrb200_collect_edges_host": (C.c_int, [c_vp, C.c_int64, C.c_int64, c_vp, c_vp]),
    "mrb200_scene_info": (C.c_int, [c_vp, C.c_int, c_i32p]),
    "mrb200_scene_set_two_phase": (C.c_int, [c_vp, C.c_int, C.c_int]),
    "mrb200_scene_get_two_phase": (C.c_int, [c_vp, C.c_int, c_i32p]),
    "mrb200_batch_dist": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, c_vp, c_vp]),
    "mrb200_batch_cost": (C.c_int, [c_vp, C.c_int, c_vp, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, C.c_double, c_vp,
                                    c_vp]),
    "mrb200_minplus_cost": (C.c_int, [c_vp, C.c_int64, c_vp, c_vp, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, C.c_double, c_vp,
                                      c_vp, c_vp]),
    "mrb200_knn_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "mrb200_knn_stats_offset": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "mrb200_knn": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp,
                             C.c_size_t, C.c_int, c_vp]),
    "mrb200_radius_tc_workspace_bytes": (C.c_size_t, [C.c_int64, C.c_int64, C.c_int, C.c_int]),
    "mrb200_radius_tc_count": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, c_vp, C.c_double, C.c_int,
                                         C.c_int, c_vp, C.c_size_t, c_vp, c_vp]),
    "mrb200_radius_tc_fill": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, C.c_int, c_vp, C.c_size_t,
                                        c_vp, c_vp, c_vp, c_vp]),
    "mrb200_radius_splits": (C.c_int, [C.c_int64, C.c_int64]),
    "mrb200_radius_count": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, c_vp, C.c_double,
                                      C.c_int, C.c_int, c_vp, c_vp]),
    "mrb200_radius_fill": (C.c_int, [c_vp, c_vp, C.c_int64, C.c_int64, C.c_int, c_i32p, C.c_int, C.c_int, c_vp, C.c_double,
                                     C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp]),
}


class Mrb200Error(RuntimeError):
    pass


def load():
    """Load libmrb200.so; raises if it has not been built (python -m ...build or
    __graft_entry__.build())."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise Mrb200Error(
                f"{LIB_PATH} is missing: build the CUDA extension first "
                "(python multirobot_pathplanning_benchmark_b200/build.py). There is no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().mrb200_last_error().decode(errors="replace")
        raise Mrb200Error(f"{what or 'mrb200'} failed ({rc}): {msg}")
