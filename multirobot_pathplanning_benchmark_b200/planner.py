"""A whole-batch multi-modal PRM built only from the backend's batch calls.

The reference's planners issue one collision / neighbour query per Python call
(P/planners/composite_prm_planner.py:583-894, P/planners/prm/prm_graph.py:389-747,
P/ = src/multi_robot_multi_goal_planning/); SURVEY.md 8(f)1 names speculative batching as what
actually moves time-to-first-solution.  This module is the batch-native form of that loop for
problems whose tasks form a fixed sequence (SequenceMixin problems such as rai.2d_handover or
rai.box_stacking): per mode it validates whole sample batches, builds the k-NN graph in one call,
validates all candidate edges in one call, links consecutive modes through transition
configurations valid in both, and searches the layered graph.  It is used by bench.py to measure
time-to-first-solution with the B200 backend and, for comparison, with a CPU backend that answers
the very same calls.

Costs follow the reference: per-robot euclidean distance reduced by max + 0.01 * sum
(P/problems/core/configuration.py:156-171, 437-510).
"""
from __future__ import annotations

import time
from dataclasses import dataclass, field
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.sparse import coo_matrix
from scipy.sparse.csgraph import dijkstra


@dataclass
class SeqTask:
    """One step of the task sequence: `robots` must reach `goal` (their stacked joint values); afterwards
    `relink` = (parent frame, child frame) re-parents an object (pick / handover / place), or None (goto)."""
    robots: List[str]
    goal: np.ndarray
    relink: Optional[Tuple[str, str]] = None


@dataclass
class PRMResult:
    path: Optional[List[Tuple[int, np.ndarray]]]   # (mode index, configuration)
    cost: float
    time_s: float
    stats: dict = field(default_factory=dict)


def config_cost(a: np.ndarray, b: np.ndarray, slices: Sequence[Tuple[int, int]], w: float = 0.01) -> np.ndarray:
    """batch_config_cost(metric='euclidean', reduction='max') of the reference, rows of a vs rows of b."""
    d = np.stack([np.linalg.norm(a[:, s:e] - b[:, s:e], axis=1) for s, e in slices])
    return d.max(axis=0) + w * d.sum(axis=0)


class BatchedPRM:
    """model: env.SceneModel (or anything with slot_for / check_configs / check_edges / base scene).
    knn: callable(queries[Q,D] f64, corpus[N,D] f64, slices, metric, k) -> idx[Q,k] (numpy or tensor)."""

    def __init__(self, model, tasks: List[SeqTask], start: np.ndarray, knn: Callable, seed: int = 0,
                 samples_per_mode: int = 2000, transitions_per_mode: int = 200, k: Optional[int] = None,
                 metric: str = "max_euclidean", max_rounds: int = 6):
        self.model = model
        self.scene = model.base
        self.tasks = tasks
        self.start = np.asarray(start, np.float64)
        self.knn = knn
        self.rng = np.random.RandomState(seed)
        self.n0, self.t0, self.k, self.metric, self.max_rounds = samples_per_mode, transitions_per_mode, k, metric, max_rounds
        sl = self.scene.robot_slices()
        self.slices = [sl[r] for r in self.scene.robots]
        self.lim = self.scene.limits()
        self.D = self.scene.dof
        self.stats = {"config_checks": 0, "edge_checks": 0, "knn_queries": 0, "rounds": 0}

    # ---- helpers ------------------------------------------------------------------------------
    @staticmethod
    def _np(x):
        return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)

    def _free(self, slot, q):
        self.stats["config_checks"] += len(q)
        return self._np(self.model.check_configs(slot, q.astype(np.float32))).astype(bool)

    def _edges_free(self, slot, q1, q2):
        self.stats["edge_checks"] += len(q1)
        return self._np(self.model.check_edges(slot, q1.astype(np.float32), q2.astype(np.float32))[0]).astype(bool)

    def _mode_slots(self):
        """device slot of every mode: mode i = the first i tasks done (their relinks applied at the goal)."""
        slots, relinks, q = [], [], self.start.copy()
        sl = self.scene.robot_slices()
        slots.append(self.model.slot_for(()))
        for t in self.tasks:
            off = 0
            for r in t.robots:
                s, e = sl[r]
                q[s:e] = t.goal[off:off + e - s]
                off += e - s
            if t.relink is not None:
                relinks.append((t.relink[0], t.relink[1], q.copy()))
            key = tuple((p, c, np.round(qq, 9).tobytes()) for p, c, qq in relinks)
            slots.append(self.model.slot_for(key, list(relinks)))
        return slots

    def _uniform(self, n):
        return self.rng.uniform(self.lim[0], self.lim[1], (n, self.D)).astype(np.float32).astype(np.float64)

    def _transition_samples(self, task: SeqTask, n):
        q = self._uniform(n)
        sl = self.scene.robot_slices()
        off = 0
        for r in task.robots:
            s, e = sl[r]
            q[:, s:e] = task.goal[off:off + e - s]
            off += e - s
        return q

    # ---- main loop ----------------------------------------------------------------------------
    def plan(self, max_time: float = 300.0) -> PRMResult:
        t_start = time.perf_counter()
        slots = self._mode_slots()
        M = len(self.tasks)            # modes 0..M-1 are planned in; "mode M" is the terminal state
        nodes: List[np.ndarray] = [np.zeros((0, self.D)) for _ in range(M)]      # valid samples per mode
        trans: List[np.ndarray] = [np.zeros((0, self.D)) for _ in range(M)]      # valid transition configs out of mode i
        for rnd in range(self.max_rounds):
            self.stats["rounds"] = rnd + 1
            n_new, t_new = self.n0 * 2 ** rnd, self.t0 * 2 ** rnd
            for i, task in enumerate(self.tasks):
                q = self._uniform(n_new)
                nodes[i] = np.vstack([nodes[i], q[self._free(slots[i], q)]])
                qt = self._transition_samples(task, t_new)
                ok = self._free(slots[i], qt)
                if i + 1 < M:
                    ok &= self._free(slots[i + 1], qt)   # a transition must be valid before and after the relink
                trans[i] = np.vstack([trans[i], qt[ok]])
            res = self._search(slots, nodes, trans)
            if res is not None:
                path, cost = res
                return PRMResult(path, cost, time.perf_counter() - t_start, dict(self.stats))
            if time.perf_counter() - t_start > max_time:
                break
        return PRMResult(None, float("inf"), time.perf_counter() - t_start, dict(self.stats))

    def _search(self, slots, nodes, trans):
        M = len(self.tasks)
        # vertex layout per mode i: [entry points (start or transitions of mode i-1)] [samples] [transitions out]
        verts, offs, entry_n = [], [], []
        for i in range(M):
            entry = self.start[None] if i == 0 else trans[i - 1]
            if len(trans[i]) == 0 or len(entry) == 0:
                return None
            verts.append(np.vstack([entry, nodes[i], trans[i]]))
            entry_n.append(len(entry))
        base = np.cumsum([0] + [len(v) for v in verts])
        rows, cols, vals = [], [], []
        for i in range(M):
            V = verts[i]
            k = self.k or min(len(V) - 1, int(np.e * (1 + 1 / self.D) * np.log(len(V))) + 1)
            if k < 1:
                return None
            idx = self._np(self.knn(V, V, self.slices, self.metric, k + 1))[:, 1:]   # drop self
            self.stats["knn_queries"] += len(V)
            a = np.repeat(np.arange(len(V)), idx.shape[1])
            b = idx.reshape(-1)
            # undirected: every unordered pair {a, b} that appears in either endpoint's list is checked once
            valid = (b >= 0) & (a != b)
            lo, hi = np.minimum(a[valid], b[valid]).astype(np.int64), np.maximum(a[valid], b[valid]).astype(np.int64)
            key = np.unique(lo * len(V) + hi)               # sorted by (lo, hi)
            e = np.stack([key // len(V), key % len(V)], axis=1)
            if len(e) == 0:
                return None
            ok = self._edges_free(slots[i], V[e[:, 0]], V[e[:, 1]])
            e = e[ok]
            c = config_cost(V[e[:, 0]], V[e[:, 1]], self.slices)
            rows += [base[i] + e[:, 0], base[i] + e[:, 1]]
            cols += [base[i] + e[:, 1], base[i] + e[:, 0]]
            vals += [c, c]
            if i + 1 < M:   # transition config j of mode i == entry point j of mode i+1 (zero cost, tiny epsilon for csgraph)
                tj = np.arange(len(trans[i]))
                src = base[i] + entry_n[i] + len(nodes[i]) + tj
                dst = base[i + 1] + tj
                rows += [src]
                cols += [dst]
                vals += [np.full(len(tj), 1e-9)]
        n = base[-1]
        G = coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n)).tocsr()
        dist, pred = dijkstra(G, directed=True, indices=0, return_predecessors=True)
        last = M - 1
        goal_ids = base[last] + entry_n[last] + len(nodes[last]) + np.arange(len(trans[last]))
        best = goal_ids[np.argmin(dist[goal_ids])]
        if not np.isfinite(dist[best]):
            return None
        path, v = [], int(best)
        while v >= 0:
            i = int(np.searchsorted(base, v, side="right") - 1)
            path.append((i, verts[i][v - base[i]]))
            v = int(pred[v]) if pred[v] >= 0 else -1
        return path[::-1], float(dist[best])
