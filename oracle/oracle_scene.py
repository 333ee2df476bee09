"""ctypes front-end of the fp64 CPU oracle (oracle/oracle_scene.c).

TEST INFRASTRUCTURE -- only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this.  See the header of oracle_scene.c for what it
restates and its parity status (UNPINNED against rai; primitives pinned numerically).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle_scene.so")
    src = os.path.join(_HERE, "oracle_scene.c")
    hdr = os.path.join(_HERE, "..", "multirobot_pathplanning_benchmark_b200", "csrc", "scene_blob.h")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u64p, f64p, u8p, i32p = (C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint8),
                                 C.POINTER(C.c_int32))
        L.orc_check_configs.argtypes = [u64p, f64p, C.c_int64, C.c_double, u8p, u8p, u8p, f64p, f64p, C.c_int]
        L.orc_check_configs.restype = C.c_int
        L.orc_check_edges.argtypes = [u64p, f64p, f64p, C.c_int64, C.c_double, i32p, C.c_int, C.c_int, C.c_int,
                                      C.c_double, u8p, i32p, i32p, C.c_int]
        L.orc_check_edges.restype = C.c_int
        L.orc_static_penetration.argtypes = [u64p]
        L.orc_static_penetration.restype = C.c_double
        L.orc_world_shapes.argtypes = [u64p, f64p, f64p]
        L.orc_pair_distance.argtypes = [C.c_int, f64p, f64p, C.c_double]
        L.orc_pair_distance.restype = C.c_double
        L.orc_segbox_dist2_local.argtypes = [f64p, f64p, f64p]
        L.orc_segbox_dist2_local.restype = C.c_double
        L.orc_box_box_sat.argtypes = [f64p, f64p]
        L.orc_box_box_sat.restype = C.c_double
        L.orc_box_box_exact_dist.argtypes = [f64p, f64p]
        L.orc_box_box_exact_dist.restype = C.c_double
        L.orc_binary_indices.argtypes = [C.c_int, i32p]
        L.orc_max_threads.restype = C.c_int
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(a):
    return np.ascontiguousarray(a, np.float64)


def max_threads() -> int:
    """host cores this process may use.  NOT omp_get_max_threads(): launchers such as torch.distributed.run export
    OMP_NUM_THREADS=1, which silently turned the "all host threads" CPU baseline into a single-thread one; every
    oracle entry point passes its thread count explicitly (`num_threads` clause), so the environment cannot override it."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def check_configs(blob64: np.ndarray, q, tol: float = -1.0, rel=None, oth=None, nthreads: int = 1):
    """-> (free[B] bool, total_penetration[B], min_pair_distance[B])"""
    q = _f64(q)
    B = len(q)
    flags = np.zeros(B, np.uint8)
    pen = np.zeros(B)
    mind = np.zeros(B)
    relp = othp = None
    if rel is not None:
        rel = np.ascontiguousarray(rel, np.uint8)
        oth = np.ascontiguousarray(oth, np.uint8)
        relp, othp = _p(rel, C.c_uint8), _p(oth, C.c_uint8)
    rc = lib().orc_check_configs(_p(blob64, C.c_uint64), _p(q, C.c_double), B, tol, relp, othp,
                                 _p(flags, C.c_uint8), _p(pen, C.c_double), _p(mind, C.c_double), nthreads)
    if rc:
        raise RuntimeError(f"orc_check_configs -> {rc}")
    return flags.astype(bool), pen, mind


def margin(pen: np.ndarray, mind: np.ndarray, tol: float) -> np.ndarray:
    """Signed clearance of the flag decision: tol - sum(pen) when something penetrates, else
    tol + min pair distance.  Flags must agree wherever |margin| > 1e-5 (BASELINE.json)."""
    return np.where(pen > 0, tol - pen, tol + np.maximum(mind, 0.0))


def check_edges(blob64, q1, q2, resolution: float, Ns=None, n_start=0, n_max=-1, include_endpoints=False,
                tol: float = -1.0, nthreads: int = 1):
    """-> (free[E] bool, first_colliding_position[E] int32 (-1 = none), checks[E])"""
    q1, q2 = _f64(q1), _f64(q2)
    E = len(q1)
    flags = np.zeros(E, np.uint8)
    first = np.zeros(E, np.int32)
    checks = np.zeros(E, np.int32)
    nsp = None
    if Ns is not None:
        Ns = np.ascontiguousarray(Ns, np.int32)
        nsp = _p(Ns, C.c_int32)
    rc = lib().orc_check_edges(_p(blob64, C.c_uint64), _p(q1, C.c_double), _p(q2, C.c_double), E, resolution, nsp,
                               n_start, n_max, int(include_endpoints), tol, _p(flags, C.c_uint8),
                               _p(first, C.c_int32), _p(checks, C.c_int32), nthreads)
    if rc:
        raise RuntimeError(f"orc_check_edges -> {rc}")
    return flags.astype(bool), first, checks


def world_shapes(blob64, q, n_shapes: int) -> np.ndarray:
    W = np.zeros((n_shapes, 16))
    q = _f64(q)
    lib().orc_world_shapes(_p(blob64, C.c_uint64), _p(q, C.c_double), _p(W, C.c_double))
    return W


def pair_distance(ptype: int, wa, wb, rsum: float) -> float:
    wa, wb = _f64(wa), _f64(wb)
    return float(lib().orc_pair_distance(ptype, _p(wa, C.c_double), _p(wb, C.c_double), rsum))


def static_penetration(blob64) -> float:
    return float(lib().orc_static_penetration(_p(blob64, C.c_uint64)))


def binary_indices(N: int) -> np.ndarray:
    seq = np.zeros(N, np.int32)
    lib().orc_binary_indices(N, _p(seq, C.c_int32))
    return seq
