"""CPU ORACLE (test infrastructure, NOT product code) -- numpy restatement of the
reference's importable hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product package never does.

Every function cites the reference file:line it restates (paths relative to
/root/reference, P/ = src/multi_robot_multi_goal_planning/).

Parity status: PINNED.  tests/golden/abstract_golden.npz holds outputs of the
*unmodified reference* (imported in the build container by
scripts/make_golden_abstract.py) and tests/test_oracle_abstract.py checks this
file against them bit-for-bit (flags, indices) / to 1 ulp (fp64 distances; the
reference's numba kernels use fastmath).
"""
from __future__ import annotations

from collections import deque
from functools import lru_cache
from typing import Optional, Sequence, Tuple

import numpy as np


# --------------------------------------------------------------------------------------
# A8: edge discretisation order.  P/problems/planning_env.py:34-51
# --------------------------------------------------------------------------------------
@lru_cache(maxsize=None)
def binary_search_indices(N: int) -> Tuple[int, ...]:
    """Breadth-first midpoint order of range(N) (planning_env.py:34-51)."""
    seq = [0] * N
    queue = deque([(0, N - 1)])
    k = 0
    while queue:
        s, e = queue.popleft()
        mid = (s + e) // 2
        seq[k] = mid
        k += 1
        if s <= mid - 1:
            queue.append((s, mid - 1))
        if mid + 1 <= e:
            queue.append((mid + 1, e))
    return tuple(seq)


def binary_index_closed_form(N: int, p: int) -> int:
    """p-th element of binary_search_indices(N) without building the table.

    This is the *device* algorithm (csrc/binary_order.cuh) restated on the host so the
    CPU suite can check it against the BFS above for every N.  Levels 0..H-1 of the
    midpoint tree are full (H = floor(log2 N)); the last level holds the remainder and
    is enumerated left to right, skipping empty intervals.
    """
    H = N.bit_length() - 1
    t = min((p + 1).bit_length() - 1, H)
    r = p - ((1 << t) - 1)
    s, n, h = 0, N, t

    def cnt(n_, h_):  # nodes at depth h_ in a balanced subtree holding n_ elements
        return max(0, min(n_ - ((1 << h_) - 1), 1 << h_))

    while True:
        nl = (n - 1) // 2
        mid = s + nl
        if h == 0:
            return mid
        cl = cnt(nl, h - 1)
        if r < cl:
            n = nl
        else:
            r -= cl
            s = mid + 1
            n = n - 1 - nl
        h -= 1


def edge_num_points(q1: np.ndarray, q2: np.ndarray, resolution: float) -> int:
    """N = max(2, int(|q2-q1|_inf / resolution) + 1)  (abstract_env.py:321-323,
    rai_base_env.py:638-640; config_dist(...,"max") == compute_abs_max_reduction,
    configuration.py:180-197, 328-329)."""
    d = float(np.max(np.abs(np.asarray(q1, np.float64) - np.asarray(q2, np.float64))))
    return max(2, int(d / resolution) + 1)


def edge_positions(N: int, N_start: int = 0, N_max: Optional[int] = None,
                   include_endpoints: bool = False):
    """Interpolation indices the reference visits, in order (abstract_env.py:328-345).
    Yields (position_in_binary_order, i)."""
    if N_max is None:
        N_max = N
    N_max = min(N, N_max)
    idx = binary_search_indices(N)
    for p in range(N_start, N_max):
        i = idx[p]
        if not include_endpoints and (i == 0 or i == N - 1):
            continue
        yield p, i


def edge_interpolate(q1: np.ndarray, q2: np.ndarray, N: int, i: int) -> np.ndarray:
    """q = q1 + ((q2-q1)/(N-1)) * i, fp64, this operation order (abstract_env.py:341,348)."""
    d = (q2 - q1) / (N - 1)
    return q1 + d * i


# --------------------------------------------------------------------------------------
# A3/A4: abstract sphere-agent environment.  P/problems/abstract_env.py:41-109, 255-299
# --------------------------------------------------------------------------------------
class AbstractScene:
    """Geometry of an AbstractEnvironment: n_agents spheres of dimension `dim`, sphere and
    axis-aligned box obstacles.  abstract.test = make_middle_obstacle_n_dim_env
    (abstract_env.py:363-378, 386-390)."""

    def __init__(self, n_agents: int, dim: int, radii: Sequence[float],
                 spheres: Sequence[Tuple[Sequence[float], float]] = (),
                 rects: Sequence[Tuple[Sequence[float], Sequence[float]]] = ()):
        self.n_agents = n_agents
        self.dim = dim
        self.radii = np.asarray(radii, np.float64)
        self.sph_c = np.asarray([c for c, _ in spheres], np.float64).reshape(-1, dim)
        self.sph_r = np.asarray([r for _, r in spheres], np.float64)
        # Rectangle.__init__ (abstract_env.py:62-67): min/max = center -/+ bounds/2
        self.rect_min = np.asarray([np.asarray(c, np.float64) - np.asarray(b, np.float64) / 2
                                    for c, b in rects], np.float64).reshape(-1, dim)
        self.rect_max = np.asarray([np.asarray(c, np.float64) + np.asarray(b, np.float64) / 2
                                    for c, b in rects], np.float64).reshape(-1, dim)

    @staticmethod
    def abstract_test() -> "AbstractScene":
        return AbstractScene(2, 2, [0.1, 0.1], spheres=[([0.0, 0.0], 0.2)],
                             rects=[([0.0, 0.4], [0.5, 0.5])])

    # ---- scalar restatement (exactly the reference's control flow) ----
    def is_collision_free(self, q: np.ndarray) -> bool:
        """abstract_env.py:255-276.  robot-robot `<`, sphere obstacle `<` (:47),
        rectangle `<=` on squared distance (:79-84)."""
        q = np.asarray(q, np.float64).reshape(self.n_agents, self.dim)
        for i in range(self.n_agents):
            for j in range(i + 1, self.n_agents):
                if np.linalg.norm(q[i] - q[j]) < self.radii[i] + self.radii[j]:
                    return False
        for i in range(self.n_agents):
            for c, r in zip(self.sph_c, self.sph_r):
                if np.linalg.norm(c - q[i]) < r + self.radii[i]:
                    return False
            for lo, hi in zip(self.rect_min, self.rect_max):
                cp = np.clip(q[i], lo, hi)
                if np.sum((cp - q[i]) ** 2) <= self.radii[i] ** 2:
                    return False
        return True

    # ---- vectorised restatement (same comparisons, whole batch) ----
    def batch_flags(self, qs: np.ndarray) -> np.ndarray:
        """Per-item flags (True = free) for qs[B, n_agents*dim]; the reference's own batch
        variant (abstract_env.py:278-299) only returns all(flags).

        np.linalg.norm goes through BLAS ddot, which on this image's OpenBLAS is a sequential
        FMA chain for short vectors; plain numpy arithmetic differs from it by <= 1 ulp.  Rows
        whose decision is within a few ulp of a threshold are therefore re-evaluated with the
        scalar restatement above (which calls np.linalg.norm exactly like the reference)."""
        flat = np.asarray(qs, np.float64)
        qs = flat.reshape(len(flat), self.n_agents, self.dim)
        free = np.ones(len(qs), bool)
        near = np.zeros(len(qs), bool)
        eps = 8 * np.finfo(np.float64).eps

        def test(val, thr, inclusive):
            nonlocal free, near
            free &= ~((val <= thr) if inclusive else (val < thr))
            near |= np.abs(val - thr) <= eps * np.maximum(np.abs(thr), 1e-300)

        for i in range(self.n_agents):
            for j in range(i + 1, self.n_agents):
                d = qs[:, i] - qs[:, j]
                test(np.sqrt(np.sum(d * d, axis=1)), self.radii[i] + self.radii[j], False)
        for i in range(self.n_agents):
            for c, r in zip(self.sph_c, self.sph_r):
                d = c - qs[:, i]
                test(np.sqrt(np.sum(d * d, axis=1)), r + self.radii[i], False)
            for lo, hi in zip(self.rect_min, self.rect_max):
                cp = np.clip(qs[:, i], lo, hi)
                test(np.sum((cp - qs[:, i]) ** 2, axis=1), self.radii[i] ** 2, True)
        for k in np.nonzero(near)[0]:
            free[k] = self.is_collision_free(flat[k])
        return free

    def is_edge_collision_free(self, q1, q2, resolution=0.01, include_endpoints=False,
                               N_start=0, N_max=None, N=None):
        """abstract_env.py:301-354.  Returns (flag, first_colliding_position or -1,
        number_of_config_checks_made)."""
        q1 = np.asarray(q1, np.float64)
        q2 = np.asarray(q2, np.float64)
        if N is None:
            N = edge_num_points(q1, q2, resolution)
        checks = 0
        for p, i in edge_positions(N, N_start, N_max, include_endpoints):
            checks += 1
            if not self.is_collision_free(edge_interpolate(q1, q2, N, i)):
                return False, p, checks
        return True, -1, checks

    def batch_edge_flags(self, q1s, q2s, resolution=0.01, include_endpoints=False,
                         N_start=0, N_max=None, Ns=None):
        """Vectorised-per-edge version of the above (all positions of one edge at once,
        then first colliding position = min).  Same results, much faster."""
        q1s = np.asarray(q1s, np.float64)
        q2s = np.asarray(q2s, np.float64)
        E = len(q1s)
        flags = np.ones(E, bool)
        first = np.full(E, -1, np.int32)
        for e in range(E):
            N = int(Ns[e]) if Ns is not None else edge_num_points(q1s[e], q2s[e], resolution)
            pos = list(edge_positions(N, N_start, N_max, include_endpoints))
            if not pos:
                continue
            ps = np.array([p for p, _ in pos])
            iis = np.array([i for _, i in pos], np.float64)
            d = (q2s[e] - q1s[e]) / (N - 1)
            qs = q1s[e][None, :] + d[None, :] * iis[:, None]
            f = self.batch_flags(qs)
            if not f.all():
                flags[e] = False
                first[e] = ps[np.argmin(f)]
        return flags, first


# --------------------------------------------------------------------------------------
# A10/A11: distance metrics and costs.  P/problems/core/configuration.py:101-219, 303-329, 437-510
# --------------------------------------------------------------------------------------
def batch_config_dist(q: np.ndarray, pts: np.ndarray, slices: np.ndarray,
                      metric: str = "max") -> np.ndarray:
    """One-to-many distance (configuration.py:305-329).  Sequential left-to-right sums of
    squares like compute_sliced_euclidean_dists (:107-126); the reference compiles that
    loop with fastmath so its own results are only reproducible to ~1 ulp."""
    diff = np.asarray(q, np.float64)[None, :] - np.asarray(pts, np.float64)
    if metric == "euclidean":
        return _sliced_norm(diff, 0, diff.shape[1])
    if metric in ("sum_euclidean", "max_euclidean"):
        per = np.stack([_sliced_norm(diff, s, e) for s, e in slices])
        return per.sum(axis=0) if metric == "sum_euclidean" else per.max(axis=0)
    return np.max(np.abs(diff), axis=1)


def _sliced_norm(diff, s, e):
    acc = np.zeros(diff.shape[0])
    for k in range(s, e):
        acc = acc + diff[:, k] * diff[:, k]
    return np.sqrt(acc)


def batch_config_cost(diff: np.ndarray, slices: np.ndarray, metric: str = "max",
                      reduction: str = "max", w: float = 0.01) -> np.ndarray:
    """_batch_config_cost_impl (configuration.py:491-510): per-robot euclidean or max-abs,
    reduced by sum or by max + w*sum (compute_max_sum_reduction :156-171)."""
    diff = np.asarray(diff, np.float64)
    if metric == "euclidean":
        per = np.stack([_sliced_norm(diff, s, e) for s, e in slices])
    else:
        per = np.stack([np.max(np.abs(diff[:, s:e]), axis=1) for s, e in slices])
    if reduction == "max":
        sm = per[0].copy()
        for i in range(1, len(per)):
            sm = sm + per[i]
        return per.max(axis=0) + w * sm
    if reduction == "sum":
        sm = per[0].copy()
        for i in range(1, len(per)):
            sm = sm + per[i]
        return sm
    raise ValueError


# --------------------------------------------------------------------------------------
# A12-A14: neighbour selection.  P/planners/prm/prm_graph.py:389-549,
# P/planners/rrtstar_base.py:439-453, P/planners/itstar_base.py:1345-1527
# --------------------------------------------------------------------------------------
def prm_k_star(N: int, D: int) -> int:
    """k* = int(e*(1+1/D)*ln N)+1 clipped to N (prm_graph.py:440-445)."""
    k = int(np.e * (1 + 1 / D) * np.log(N)) + 1
    return min(k, N)


def knn_indices(dists: np.ndarray, k: int) -> np.ndarray:
    """prm_graph.py:446-447: argpartition then argsort of the k kept -> ascending distance.
    Ties are unspecified in the reference; here they break by index (stable), which is
    what the device re-rank reproduces."""
    k = min(k, len(dists))
    order = np.lexsort((np.arange(len(dists)), dists))
    return order[:k]


def radius_indices(dists: np.ndarray, r: float, inclusive_eps: Optional[float] = None) -> np.ndarray:
    """PRM r-disc: dists < r in index order (prm_graph.py:500); RRT*/IT*:
    dists <= r + 1e-10 (rrtstar_base.py:439-453, itstar_base.py:1388-1527)."""
    if inclusive_eps is None:
        return np.nonzero(dists < r)[0]
    return np.nonzero(dists <= r + inclusive_eps)[0]
