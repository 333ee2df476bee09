"""Tensor-in / tensor-out batch API over the C ABI (include/mrb200.h).

PyTorch is plumbing only: device memory, streams, torch.distributed.  Every method takes
CUDA tensors, passes raw pointers + the current stream to libmrb200.so and returns CUDA
tensors; nothing here computes collisions on the host.

Reference calls these replace, batched (P/ = src/multi_robot_multi_goal_planning/ in the
reference): is_collision_free (P/problems/rai_base_env.py:442-513, abstract_env.py:255-276),
is_collision_free_for_robot (rai_base_env.py:515-615), is_edge_collision_free
(rai_base_env.py:618-676, abstract_env.py:301-354).
"""
from __future__ import annotations

import contextlib
import ctypes as C
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .scene import CompiledScene


def _require_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise ValueError(f"{name} must be a CUDA tensor (there is no CPU path)")
    if t.dtype != dtype:
        raise ValueError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class AbstractBackend:
    """Sphere agents vs sphere / box obstacles (AbstractEnvironment), fp64, bit-exact."""

    def __init__(self, n_agents: int, dim: int, radii: Sequence[float],
                 spheres: Sequence[Tuple[Sequence[float], float]] = (),
                 rects_minmax: Sequence[Tuple[Sequence[float], Sequence[float]]] = (), device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.n_agents, self.dim = n_agents, dim
        self.D = n_agents * dim
        radii = np.ascontiguousarray(radii, np.float64)
        sph = np.ascontiguousarray([list(c) + [r] for c, r in spheres], np.float64).reshape(-1)
        rect = np.ascontiguousarray([list(lo) + list(hi) for lo, hi in rects_minmax], np.float64).reshape(-1)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_abstract_create(
                n_agents, dim, radii.ctypes.data_as(_lib.c_f64p), len(spheres),
                sph.ctypes.data_as(_lib.c_f64p) if len(spheres) else None, len(rects_minmax),
                rect.ctypes.data_as(_lib.c_f64p) if len(rects_minmax) else None, C.byref(h)), "abstract_create")
        self.handle = h

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mrb200_abstract_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def check_configs(self, q: torch.Tensor) -> torch.Tensor:
        """q [B, D] float64 cuda -> bool [B], True = collision free."""
        q = _require_cuda(q, torch.float64, "q")
        if q.dim() != 2 or q.shape[1] != self.D:
            raise ValueError(f"q must be [B, {self.D}]")
        out = torch.empty(q.shape[0], dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            _lib.check(self.lib.mrb200_abstract_check_configs(self.handle, q.data_ptr(), q.shape[0], out.data_ptr(),
                                                              _stream(q.device)), "abstract_check_configs")
        return out.view(torch.bool)

    def check_edges(self, q1: torch.Tensor, q2: torch.Tensor, resolution: float, N: Optional[torch.Tensor] = None,
                    n_start: int = 0, n_max: Optional[int] = None, include_endpoints: bool = False):
        """-> (free bool [E], first colliding position int32 [E], -1 if none)."""
        q1 = _require_cuda(q1, torch.float64, "q1")
        q2 = _require_cuda(q2, torch.float64, "q2")
        if q1.shape != q2.shape or q1.dim() != 2 or q1.shape[1] != self.D:
            raise ValueError(f"q1, q2 must both be [E, {self.D}]")
        if N is not None:
            N = _require_cuda(N, torch.int32, "N")
        E = q1.shape[0]
        free = torch.empty(E, dtype=torch.uint8, device=q1.device)
        first = torch.empty(E, dtype=torch.int32, device=q1.device)
        with torch.cuda.device(q1.device):
            _lib.check(self.lib.mrb200_abstract_check_edges(
                self.handle, q1.data_ptr(), q2.data_ptr(), E, float(resolution), N.data_ptr() if N is not None else None,
                int(n_start), -1 if n_max is None else int(n_max), int(include_endpoints), free.data_ptr(),
                first.data_ptr(), _stream(q1.device)), "abstract_check_edges")
        return free.view(torch.bool), first


    # ---- single-query seam: numpy fp64 in, numpy out, one library call ----
    def query_configs_host(self, q: np.ndarray) -> np.ndarray:
        q = np.ascontiguousarray(q, np.float64)
        if q.ndim != 2 or q.shape[1] != self.D:
            raise ValueError(f"q must be [B, {self.D}]")
        out = np.empty(q.shape[0], np.uint8)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_abstract_query_configs_host(self.handle, q.ctypes.data, q.shape[0], out.ctypes.data,
                                                                   _stream(self.device)), "abstract_query_configs_host")
        return out.view(np.bool_)

    def query_edges_host(self, q1: np.ndarray, q2: np.ndarray, resolution: float, N: Optional[np.ndarray] = None, n_start: int = 0,
                         n_max: Optional[int] = None, include_endpoints: bool = False):
        q1, q2 = np.ascontiguousarray(q1, np.float64), np.ascontiguousarray(q2, np.float64)
        if q1.shape != q2.shape or q1.ndim != 2 or q1.shape[1] != self.D:
            raise ValueError(f"q1, q2 must both be [E, {self.D}]")
        E = q1.shape[0]
        Nh = None if N is None else np.ascontiguousarray(N, np.int32)
        free, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_abstract_query_edges_host(
                self.handle, q1.ctypes.data, q2.ctypes.data, E, float(resolution), Nh.ctypes.data if Nh is not None else None,
                int(n_start), -1 if n_max is None else int(n_max), int(include_endpoints), free.ctypes.data, first.ctypes.data,
                _stream(self.device)), "abstract_query_edges_host")
        return free.view(np.bool_), first


class SceneBackend:
    """Primitive scene with per-mode slots (one compiled blob per mode)."""

    def __init__(self, max_modes: int = 64, device=None):
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_scene_create(int(max_modes), C.byref(h)), "scene_create")
        self.handle = h
        self.max_modes = max_modes
        self.compiled = {}

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.mrb200_scene_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    def set_mode(self, slot: int, cs: CompiledScene) -> None:
        blob = np.ascontiguousarray(cs.blob32)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_scene_set_mode(self.handle, int(slot), blob.ctypes.data, blob.nbytes,
                                                      _stream(self.device)), "scene_set_mode")
        self.compiled[slot] = cs

    def info(self, slot: int):
        out = (C.c_int32 * 4)()
        _lib.check(self.lib.mrb200_scene_info(self.handle, int(slot), out), "scene_info")
        return {"D": out[0], "n_shapes": out[1], "n_pairs": out[2], "smem_bytes": out[3]}

    TWO_PHASE_POLICIES = {"auto": 0, "always": 1, "never": 2}

    def set_two_phase(self, slot: int, policy: str = "auto") -> None:
        """Kernel choice for large plain batches on this slot: "auto" (measure what the table / floor bound decides on
        the first large batches, then settle), "always" or "never" (two-phase tiles).  Flags do not depend on it."""
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mrb200_scene_set_two_phase(self.handle, int(slot), self.TWO_PHASE_POLICIES[policy]),
                       "scene_set_two_phase")

    def two_phase_info(self, slot: int):
        out = (C.c_int32 * 3)()
        _lib.check(self.lib.mrb200_scene_get_two_phase(self.handle, int(slot), out), "scene_get_two_phase")
        return {"state": ("measuring", "two_phase", "single_pass")[out[0]], "seen": out[1], "decided_by_bound": out[2]}

    def _q(self, slot, q, name="q"):
        q = _require_cuda(q, torch.float32, name)
        D = self.compiled[slot].dof
        if q.dim() != 2 or q.shape[1] != D:
            raise ValueError(f"{name} must be [B, {D}]")
        return q

    def check_configs(self, slot: int, q: torch.Tensor, tol: Optional[float] = None, return_penetration: bool = False,
                      full_eval: bool = False, out: Optional[torch.Tensor] = None):
        """q [B, D] float32 cuda -> bool [B] (True = free) [, total penetration float32 [B]]."""
        q = self._q(slot, q)
        B = q.shape[0]
        flags = out if out is not None else torch.empty(B, dtype=torch.uint8, device=q.device)
        pen = torch.empty(B, dtype=torch.float32, device=q.device) if return_penetration else None
        with torch.cuda.device(q.device):
            _lib.check(self.lib.mrb200_check_configs(
                self.handle, int(slot), q.data_ptr(), B, -1.0 if tol is None else float(tol), flags.data_ptr(),
                pen.data_ptr() if pen is not None else None, int(full_eval or return_penetration),
                _stream(q.device)), "check_configs")
        f = flags.view(torch.bool)
        return (f, pen) if return_penetration else f

    def check_configs_for_robot(self, slot: int, q: torch.Tensor, relevant: np.ndarray, other: np.ndarray,
                                tol: Optional[float] = None) -> torch.Tensor:
        q = self._q(slot, q)
        relevant = np.ascontiguousarray(relevant, np.uint8)
        other = np.ascontiguousarray(other, np.uint8)
        flags = torch.empty(q.shape[0], dtype=torch.uint8, device=q.device)
        with torch.cuda.device(q.device):
            _lib.check(self.lib.mrb200_check_configs_for_robot(
                self.handle, int(slot), q.data_ptr(), q.shape[0], -1.0 if tol is None else float(tol),
                relevant.ctypes.data_as(_lib.c_u8p), other.ctypes.data_as(_lib.c_u8p), len(relevant),
                flags.data_ptr(), _stream(q.device)), "check_configs_for_robot")
        return flags.view(torch.bool)

    def check_edges(self, slot: int, q1: torch.Tensor, q2: torch.Tensor, resolution: float,
                    N: Optional[torch.Tensor] = None, n_start: int = 0, n_max: Optional[int] = None,
                    include_endpoints: bool = False, tol: Optional[float] = None):
        """-> (free bool [E], first colliding position in binary order int32 [E], -1 if none)."""
        q1 = self._q(slot, q1, "q1")
        q2 = self._q(slot, q2, "q2")
        if q1.shape != q2.shape:
            raise ValueError("q1 and q2 must have the same shape")
        if N is not None:
            N = _require_cuda(N, torch.int32, "N")
        E = q1.shape[0]
        free = torch.empty(E, dtype=torch.uint8, device=q1.device)
        first = torch.empty(E, dtype=torch.int32, device=q1.device)
        with torch.cuda.device(q1.device):
            _lib.check(self.lib.mrb200_check_edges(
                self.handle, int(slot), q1.data_ptr(), q2.data_ptr(), E, float(resolution),
                N.data_ptr() if N is not None else None, int(n_start), -1 if n_max is None else int(n_max),
                int(include_endpoints), -1.0 if tol is None else float(tol), free.data_ptr(), first.data_ptr(),
                _stream(q1.device)), "check_edges")
        return free.view(torch.bool), first


    # ---- the planners' one-query-at-a-time seam: numpy in, numpy out, one library call (copies + launch + sync) ----
    def query_configs_host(self, slot: int, q: np.ndarray, tol: Optional[float] = None, relevant: Optional[np.ndarray] = None,
                           other: Optional[np.ndarray] = None) -> np.ndarray:
        q = np.ascontiguousarray(q, np.float32)
        if q.ndim != 2 or q.shape[1] != self.compiled[slot].dof:
            raise ValueError(f"q must be [B, {self.compiled[slot].dof}]")
        out = np.empty(q.shape[0], np.uint8)
        rel = oth = None
        n = 0
        if relevant is not None:
            rel, oth = np.ascontiguousarray(relevant, np.uint8), np.ascontiguousarray(other, np.uint8)
            n = len(rel)
        with self._on_device():
            rc = self.lib.mrb200_query_configs_host(
                self.handle, slot, q.ctypes.data, q.shape[0], -1.0 if tol is None else float(tol),
                rel.ctypes.data if rel is not None else None, oth.ctypes.data if oth is not None else None, n,
                out.ctypes.data, _stream(self.device))
        if rc:
            _lib.check(rc, "query_configs_host")
        return out

    def query_edges_host(self, slot: int, q1: np.ndarray, q2: np.ndarray, resolution: float, N: Optional[np.ndarray] = None,
                         n_start: int = 0, n_max: Optional[int] = None, include_endpoints: bool = False,
                         tol: Optional[float] = None):
        q1, q2 = np.ascontiguousarray(q1, np.float32), np.ascontiguousarray(q2, np.float32)
        if q1.shape != q2.shape or q1.ndim != 2 or q1.shape[1] != self.compiled[slot].dof:
            raise ValueError(f"q1, q2 must both be [E, {self.compiled[slot].dof}]")
        E = q1.shape[0]
        Nh = None if N is None else np.ascontiguousarray(N, np.int32)
        free, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        with self._on_device():
            rc = self.lib.mrb200_query_edges_host(
                self.handle, slot, q1.ctypes.data, q2.ctypes.data, E, float(resolution), Nh.ctypes.data if Nh is not None else None,
                int(n_start), -1 if n_max is None else int(n_max), int(include_endpoints), -1.0 if tol is None else float(tol),
                free.ctypes.data, first.ctypes.data, _stream(self.device))
        if rc:
            _lib.check(rc, "query_edges_host")
        return free, first

    # ---- asynchronous host-buffer edge batches (candidate-edge speculation of env.py) ----
    def submit_edges_host(self, slot: int, q1: np.ndarray, q2: np.ndarray, resolution: float, N: Optional[np.ndarray] = None,
                          include_endpoints: bool = False, tol: Optional[float] = None) -> int:
        """q1 [D] or [E, D], q2 [E, D] (numpy); queues copies + kernel on the current stream and returns a ticket"""
        q2 = np.ascontiguousarray(q2, np.float32)
        q1 = np.ascontiguousarray(q1, np.float32).reshape(-1, q2.shape[1])
        if q2.ndim != 2 or q2.shape[1] != self.compiled[slot].dof or q1.shape[0] not in (1, q2.shape[0]):
            raise ValueError("q1 must be [D] or [E, D] and q2 [E, D]")
        Nh = None if N is None else np.ascontiguousarray(N, np.int32)
        ticket = C.c_int64(-1)
        with self._on_device():
            rc = self.lib.mrb200_submit_edges_host(
                self.handle, slot, q1.ctypes.data, q1.shape[0], q2.ctypes.data, q2.shape[0], float(resolution),
                Nh.ctypes.data if Nh is not None else None, int(include_endpoints), -1.0 if tol is None else float(tol),
                C.byref(ticket), _stream(self.device))
        if rc:
            _lib.check(rc, "submit_edges_host")
        return int(ticket.value)

    def collect_edges_host(self, ticket: int, E: int):
        free, first = np.empty(E, np.uint8), np.empty(E, np.int32)
        rc = self.lib.mrb200_collect_edges_host(self.handle, int(ticket), int(E), free.ctypes.data, first.ctypes.data)
        if rc:
            _lib.check(rc, "collect_edges_host")
        return free, first

    def _on_device(self):
        """device guard that costs nothing when this backend's GPU already is the current one (the latency-critical
        single-query path)"""
        if torch.cuda.current_device() == self.device.index:
            return contextlib.nullcontext()
        return torch.cuda.device(self.device)


def check_configs_host(be: SceneBackend, slot: int, q_host: torch.Tensor, out_host: torch.Tensor,
                       chunk: int = 1 << 18, state: Optional[dict] = None, tol: Optional[float] = None) -> None:
    """Host-buffer entry point for whole sample batches (what a CPU-side caller of the reference API uses): q_host
    [B, D] float32 and out_host [B] uint8, contiguous host tensors (pinned ones are copied from / to directly).  One
    call of `mrb200_check_configs_host`: the library overlaps the H2D copy, the kernel and the read-back of consecutive
    chunks on its own side streams and returns after every flag has landed in out_host.  (`state` is accepted for
    callers of the round-2 Python loop and ignored.)"""
    if q_host.is_cuda or out_host.is_cuda:
        raise ValueError("host tensors expected")
    if q_host.dtype != torch.float32 or out_host.dtype != torch.uint8 or not q_host.is_contiguous() or not out_host.is_contiguous():
        raise ValueError("q_host must be contiguous float32 [B, D], out_host contiguous uint8 [B]")
    B, D = q_host.shape
    if D != be.compiled[slot].dof or out_host.numel() != B:
        raise ValueError(f"q_host must be [B, {be.compiled[slot].dof}] and out_host [B]")
    with torch.cuda.device(be.device):
        _lib.check(be.lib.mrb200_check_configs_host(be.handle, slot, q_host.data_ptr(), B, -1.0 if tol is None else float(tol),
                                                    out_host.data_ptr(), int(chunk), _stream(be.device)), "check_configs_host")


def fp32_fma_peak_tflops(device=None, iters: int = 1 << 15, reps: int = 5) -> float:
    """Measured FP32 FMA throughput of this GPU (16 independent chains per thread, `iters` a multiple of 16)."""
    lib = _lib.load()
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = C.c_int32()
    with torch.cuda.device(device):
        _lib.check(lib.mrb200_fp32_probe(iters, None, C.byref(n), None), "fp32_probe")
        out = torch.empty(n.value, dtype=torch.float32, device=device)
        best = 0.0
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _lib.check(lib.mrb200_fp32_probe(iters, out.data_ptr(), C.byref(n), _stream(device)), "fp32_probe")
            e1.record()
            e1.synchronize()
            best = max(best, n.value * iters * 32.0 / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    return best


def launch_count() -> int:
    return int(_lib.load().mrb200_launch_count())
