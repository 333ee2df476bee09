"""Distance metrics and neighbour search on the device (tensor in / tensor out).

Reference calls replaced (P/ = src/multi_robot_multi_goal_planning/ in the reference):
  batch_config_dist                 P/problems/core/configuration.py:342-349 (impl :303-329)
  PRM get_neighbors k-NN / r-disc   P/planners/prm/prm_graph.py:389-549
  RRT* nearest / near               P/planners/rrtstar_base.py:439-453, 1228-1337
  IT* get_neighbors                 P/planners/itstar_base.py:1388-1527
Coordinates are fp64 like the reference's arrays; returned neighbour indices are int32.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _lib

METRICS = {"euclidean": 0, "sum_euclidean": 1, "max_euclidean": 2, "max": 3}


def _slices(slices) -> Tuple[Optional[np.ndarray], int]:
    if slices is None:
        return None, 0
    s = np.ascontiguousarray(np.asarray(slices, np.int32).reshape(-1, 2))
    return s, len(s)


def _f64(t: torch.Tensor, name: str) -> torch.Tensor:
    if not t.is_cuda or t.dtype != torch.float64:
        raise ValueError(f"{name} must be a float64 CUDA tensor (there is no CPU path)")
    return t.contiguous()


def _stream(dev) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def batch_config_dist(q: torch.Tensor, pts: torch.Tensor, slices=None, metric: str = "max") -> torch.Tensor:
    """One-to-many distance, [N] float64 (configuration.py:303-329)."""
    lib = _lib.load()
    q, pts = _f64(q.reshape(-1), "q"), _f64(pts, "pts")
    N, D = pts.shape
    if q.numel() != D:
        raise ValueError("dimension mismatch")
    s, R = _slices(slices)
    out = torch.empty(N, dtype=torch.float64, device=pts.device)
    with torch.cuda.device(pts.device):
        _lib.check(lib.mrb200_batch_dist(q.data_ptr(), pts.data_ptr(), N, D, s.ctypes.data_as(_lib.c_i32p) if s is not None else None,
                                         R, METRICS[metric], out.data_ptr(), _stream(pts.device)), "batch_dist")
    return out


def batch_config_cost(a: torch.Tensor, b: torch.Tensor, slices, metric: str = "euclidean", reduction: str = "max",
                      w: float = 0.01) -> torch.Tensor:
    """Cost between configuration(s) a ([D] or [N, D]) and rows of b [N, D] (configuration.py:437-510):
    per-robot `metric` in {"euclidean", "max"}, reduced by "max" (= max + w * sum) or "sum"."""
    lib = _lib.load()
    b = _f64(b, "b")
    a = _f64(a, "a")
    N, D = b.shape
    single = a.dim() == 1
    if (single and a.numel() != D) or (not single and a.shape != b.shape):
        raise ValueError("shape mismatch")
    s, R = _slices(slices)
    out = torch.empty(N, dtype=torch.float64, device=b.device)
    with torch.cuda.device(b.device):
        _lib.check(lib.mrb200_batch_cost(a.data_ptr(), int(single), b.data_ptr(), N, D, s.ctypes.data_as(_lib.c_i32p), R,
                                         int(metric != "euclidean"), int(reduction == "sum"), float(w), out.data_ptr(),
                                         _stream(b.device)), "batch_cost")
    return out


def minplus_cost(a: torch.Tensor, b: torch.Tensor, lb_b: torch.Tensor, slices, metric: str = "euclidean", reduction: str = "max",
                 w: float = 0.01, return_arg: bool = False):
    """out[i] = min_j (cost(a_i, b_j) + lb_b[j]) -- one relaxation step of the goal lower bound (prm_graph.py:143-220)"""
    lib = _lib.load()
    a, b, lb_b = _f64(a, "a"), _f64(b, "b"), _f64(lb_b.reshape(-1), "lb_b")
    T1, D = a.shape
    T2 = b.shape[0]
    if b.shape[1] != D or lb_b.numel() != T2:
        raise ValueError("shape mismatch")
    s, R = _slices(slices)
    out = torch.empty(T1, dtype=torch.float64, device=a.device)
    arg = torch.empty(T1, dtype=torch.int32, device=a.device) if return_arg else None
    with torch.cuda.device(a.device):
        _lib.check(lib.mrb200_minplus_cost(a.data_ptr(), T1, b.data_ptr(), lb_b.data_ptr(), T2, D, s.ctypes.data_as(_lib.c_i32p), R,
                                           int(metric != "euclidean"), int(reduction == "sum"), float(w), out.data_ptr(),
                                           arg.data_ptr() if arg is not None else None, _stream(a.device)), "minplus_cost")
    return (out, arg) if return_arg else out


def lower_bound_to_goal_layers(layers, goal_lb, slices, metric: str = "euclidean", reduction: str = "max", w: float = 0.01):
    """Goal lower bounds of the exit (transition) configurations of a SEQUENCE of modes, as whole-layer relaxations:
    `layers[m]` = [T_m, D] exit configurations of mode m (the last layer = the goal configurations with bounds `goal_lb`).
    The reference's compute_lower_bound_to_goal (prm_graph.py:143-220) reaches the same numbers by a Dijkstra over single
    nodes, one batch_config_cost call per popped node; for a sequence of modes every path visits the layers in order, so
    lb_m[i] = min_j (cost(layer_m[i], layer_{m+1}[j]) + lb_{m+1}[j]).  -> list of [T_m] tensors"""
    lbs = [None] * len(layers)
    lbs[-1] = _f64(goal_lb.reshape(-1), "goal_lb")
    for m in range(len(layers) - 2, -1, -1):
        lbs[m] = minplus_cost(layers[m], layers[m + 1], lbs[m + 1], slices, metric, reduction, w)
    return lbs


def prm_k_star(N: int, D: int) -> int:
    """k* = int(e (1 + 1/D) ln N) + 1, clipped to N (prm_graph.py:440-445)."""
    return min(int(math.e * (1 + 1 / D) * math.log(N)) + 1, N) if N > 0 else 0


def prm_r_star(N: int, D: int, informed_measure: float = 1.0) -> float:
    """PRM* connection radius (prm_graph.py:479-498)."""
    if N <= 1:
        return 1e6
    unit_ball = (math.pi ** 0.5) ** D / math.gamma(D / 2 + 1)
    return 1.001 * 2 * (informed_measure / unit_ball * (math.log(N) / N) * (1 + 1 / D)) ** (1 / D)


LAST_STATS = {}   # filled by batch_knn(..., stats=True): {"uncertified_rows": rows the tensor path handed to the exact kernel}


def batch_knn(queries: torch.Tensor, corpus: torch.Tensor, slices=None, metric: str = "max_euclidean", k: int = 1,
              mode: str = "auto", return_dist: bool = True, stats: bool = False):
    """k nearest corpus rows of every query row, ascending (distance, index).
    -> (idx [Q, k] int32, dist [Q, k] float64); missing neighbours are -1 / inf."""
    lib = _lib.load()
    queries, corpus = _f64(queries, "queries"), _f64(corpus, "corpus")
    Q, D = queries.shape
    N = corpus.shape[0]
    if corpus.shape[1] != D:
        raise ValueError("dimension mismatch")
    s, R = _slices(slices)
    dev = queries.device
    idx = torch.empty(Q, k, dtype=torch.int32, device=dev)
    dist = torch.empty(Q, k, dtype=torch.float64, device=dev) if return_dist else None
    with torch.cuda.device(dev):
        ws_bytes = int(lib.mrb200_knn_workspace_bytes(Q, N, D, k))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        _lib.check(lib.mrb200_knn(queries.data_ptr(), corpus.data_ptr(), Q, N, D,
                                  s.ctypes.data_as(_lib.c_i32p) if s is not None else None, R, METRICS[metric], int(k),
                                  idx.data_ptr(), dist.data_ptr() if dist is not None else None, ws.data_ptr(), ws_bytes,
                                  {"auto": 0, "exact": 1, "tensor": 2}[mode], _stream(dev)), "knn")
        if stats:   # (synchronises)
            off = int(lib.mrb200_knn_stats_offset(Q, N, D, k))
            st = ws[off:off + 8].view(torch.int32).cpu().numpy()
            LAST_STATS.clear()
            LAST_STATS.update({"uncertified_rows": int(st[1]), "max_sq_slice_norm": float(np.int32(st[0]).view(np.float32))})
    return (idx, dist) if return_dist else idx


def _radius_exact(queries, corpus, radius, radii, slices, metric, inclusive, return_dist):
    lib = _lib.load()
    Q, D = queries.shape
    N = corpus.shape[0]
    s, R = _slices(slices)
    sp = s.ctypes.data_as(_lib.c_i32p) if s is not None else None
    dev = queries.device
    with torch.cuda.device(dev):
        splits = int(lib.mrb200_radius_splits(Q, N))
        counts = torch.zeros(Q * splits, dtype=torch.int64, device=dev)
        args = (queries.data_ptr(), corpus.data_ptr(), Q, N, D, sp, R, METRICS[metric],
                radii.data_ptr() if radii is not None else None, radius, int(inclusive), splits)
        _lib.check(lib.mrb200_radius_count(*args, counts.data_ptr(), _stream(dev)), "radius_count")
        ends = torch.cumsum(counts, 0)           # plumbing: the scan between the two launches
        offs = ends - counts
        total = int(ends[-1].item()) if Q else 0
        idx = torch.empty(total, dtype=torch.int32, device=dev)
        dist = torch.empty(total, dtype=torch.float64, device=dev) if return_dist else None
        if total:
            _lib.check(lib.mrb200_radius_fill(*args, offs.data_ptr(), idx.data_ptr(), dist.data_ptr() if dist is not None else None,
                                              _stream(dev)), "radius_fill")
        row_off = torch.cat([offs.view(Q, splits)[:, 0], ends[-1:]]) if Q else torch.zeros(1, dtype=torch.int64, device=dev)
    return row_off, idx, dist


def batch_radius(queries: torch.Tensor, corpus: torch.Tensor, radius: Union[float, torch.Tensor], slices=None,
                 metric: str = "max_euclidean", inclusive: bool = False, return_dist: bool = False, mode: str = "auto",
                 cap: int = 512):
    """Radius neighbours in CSR form, indices ascending per row (the reference's np.where order).
    inclusive=False: d < r (PRM, prm_graph.py:500); True: d <= r + 1e-10 (RRT*/IT*).
    mode: "exact" = fp64 CUDA-core kernels; "tensor" = tcgen05 candidate generator + exact fp64 filter (euclidean /
    max_euclidean, at most `cap` candidates per row, rows beyond that are answered by the exact kernels); "auto" picks the
    tensor path for large selective searches.  Both return identical results.
    -> (offsets [Q+1] int64, indices int32 [, dists float64])"""
    lib = _lib.load()
    queries, corpus = _f64(queries, "queries"), _f64(corpus, "corpus")
    Q, D = queries.shape
    N = corpus.shape[0]
    dev = queries.device
    radii = None
    r = 0.0
    if isinstance(radius, torch.Tensor):
        radii = _f64(radius.reshape(-1), "radius")
        if radii.numel() != Q:
            raise ValueError("one radius per query expected")
    else:
        r = float(radius)
    tensor_ok = metric in ("euclidean", "max_euclidean") and N >= 1024 and Q >= 1
    if mode == "tensor" and not tensor_ok:
        raise ValueError("the tensor-core radius path needs metric euclidean / max_euclidean and N >= 1024")
    use_tensor = mode == "tensor" or (mode == "auto" and tensor_ok and Q >= 256 and N >= 4096)
    if not use_tensor:
        off, idx, dist = _radius_exact(queries, corpus, r, radii, slices, metric, inclusive, return_dist)
        return (off, idx, dist) if return_dist else (off, idx)
    s, R = _slices(slices)
    sp = s.ctypes.data_as(_lib.c_i32p) if s is not None else None
    with torch.cuda.device(dev):
        ws_bytes = int(lib.mrb200_radius_tc_workspace_bytes(Q, N, D, cap))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        counts = torch.empty(Q, dtype=torch.int64, device=dev)
        _lib.check(lib.mrb200_radius_tc_count(queries.data_ptr(), corpus.data_ptr(), Q, N, D, sp, R, METRICS[metric],
                                              radii.data_ptr() if radii is not None else None, r, int(inclusive), int(cap),
                                              ws.data_ptr(), ws_bytes, counts.data_ptr(), _stream(dev)), "radius_tc_count")
        over = torch.nonzero(counts < 0).flatten()        # rows with more than `cap` candidates
        n_over = int(over.numel())
        LAST_STATS["radius_overflow_rows"] = n_over
        if mode == "auto" and n_over > Q // 4:             # a radius this wide is not a candidate-list problem
            off, idx, dist = _radius_exact(queries, corpus, r, radii, slices, metric, inclusive, return_dist)
            return (off, idx, dist) if return_dist else (off, idx)
        ex = None
        if n_over:   # plumbing: the exact kernels answer the overflowed rows, their results are spliced in below
            ex = _radius_exact(queries[over].contiguous(), corpus, r, None if radii is None else radii[over].contiguous(), slices, metric,
                               inclusive, return_dist)
            counts[over] = ex[0][1:] - ex[0][:-1]
        ends = torch.cumsum(counts, 0)
        offs = (ends - counts).contiguous()
        total = int(ends[-1].item())
        idx = torch.empty(total, dtype=torch.int32, device=dev)
        dist = torch.empty(total, dtype=torch.float64, device=dev) if return_dist else None
        if total:
            _lib.check(lib.mrb200_radius_tc_fill(queries.data_ptr(), corpus.data_ptr(), Q, N, D, sp, R, METRICS[metric], int(cap),
                                                 ws.data_ptr(), ws_bytes, offs.data_ptr(), idx.data_ptr(),
                                                 dist.data_ptr() if dist is not None else None, _stream(dev)), "radius_tc_fill")
        if n_over:
            lens = ex[0][1:] - ex[0][:-1]
            dst = torch.repeat_interleave(offs[over], lens) + (torch.arange(int(ex[0][-1].item()), device=dev) -
                                                              torch.repeat_interleave(ex[0][:-1], lens))
            idx[dst] = ex[1]
            if dist is not None:
                dist[dst] = ex[2]
        row_off = torch.cat([offs, ends[-1:]])
    return (row_off, idx, dist) if return_dist else (row_off, idx)
