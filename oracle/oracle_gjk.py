"""ctypes front-end of the SECOND CPU oracle (oracle/oracle_gjk.c): generic convex GJK + EPA narrowphase on the same
kinematics and pair lists as oracle_scene.c, with rai's cylinders as true convex cylinders.

TEST INFRASTRUCTURE -- only tests/ may import this.  Parity status: see the header of oracle_gjk.c (an independent
implementation of what rai's narrowphase does; still not rai's own answers)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle_gjk.so")
        srcs = [os.path.join(_HERE, f) for f in ("oracle_gjk.c", "oracle_scene.c")]
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(s) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-s"])
        L = C.CDLL(so)
        u64p, f64p, u8p = C.POINTER(C.c_uint64), C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.orc_gjk_check_configs.argtypes = [u64p, f64p, C.c_int64, C.c_double, u8p, u8p, f64p, f64p, C.c_int]
        L.orc_gjk_check_configs.restype = C.c_int
        L.orc_gjk_pair.argtypes = [C.c_int, f64p, C.c_double, C.c_int, f64p, C.c_double]
        L.orc_gjk_pair.restype = C.c_double
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def cylinder_flags(scene, cs) -> np.ndarray:
    """1 for every collision shape that is a rai `cylinder` (the first oracle and the device model those in general
    position as capsules of the same radius and length)"""
    return np.array([1 if scene.frames[n].shape is not None and scene.frames[n].shape.kind == "cylinder" else 0
                     for n in cs.shape_names], np.uint8)


def check_configs(blob64: np.ndarray, q, is_cyl=None, tol: float = -1.0, nthreads: int = 1):
    """-> (free[B] bool, total_penetration[B], min_pair_distance[B])"""
    q = np.ascontiguousarray(q, np.float64)
    B = len(q)
    flags, pen, mind = np.zeros(B, np.uint8), np.zeros(B), np.zeros(B)
    cyl = None if is_cyl is None else np.ascontiguousarray(is_cyl, np.uint8)
    rc = lib().orc_gjk_check_configs(_p(blob64, C.c_uint64), _p(q, C.c_double), B, tol, None if cyl is None else _p(cyl, C.c_uint8),
                                     _p(flags, C.c_uint8), _p(pen, C.c_double), _p(mind, C.c_double), nthreads)
    if rc:
        raise RuntimeError(f"orc_gjk_check_configs -> {rc}")
    return flags.astype(bool), pen, mind


def pair(core_a: int, wa, core_b: int, wb, cyl_ra: float = 0.0, cyl_rb: float = 0.0) -> float:
    wa, wb = np.ascontiguousarray(wa, np.float64), np.ascontiguousarray(wb, np.float64)
    return float(lib().orc_gjk_pair(core_a, _p(wa, C.c_double), cyl_ra, core_b, _p(wb, C.c_double), cyl_rb))
