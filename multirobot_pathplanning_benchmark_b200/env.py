"""`B200Env`: the reference's environment interface on top of the CUDA backend.

The reference's planners talk to an environment only through `BaseProblem`
(P/problems/planning_env.py:1531-1994, P/ = src/multi_robot_multi_goal_planning/); backends are
classes registered with `@register("name")` (P/problems/core/registry.py:8-25).  This module
defines such classes for the B200 backend.  When the reference package is importable they derive
from its `BaseProblem` and mode-logic mixins (exactly like `rai_envs.py:573` composes
`SequenceMixin` with `rai_env`) and are registered as `b200.*`; without it the geometry / batch
API of `SceneModel` is still usable on its own (that is what the GPU tests and bench.py use).

Collision queries never run on the host: `SceneModel` owns a device (by default the CUDA
`SceneBackend`; it raises if libmrb200.so or a GPU is missing).  Tests may inject another device
object with the same four methods to exercise the host logic on a CPU-only machine.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from .scene import CompiledScene, Scene, compile_blob
from .scenes import SCENES

from .refimport import ensure_reference

try:  # the reference package (optional at import time: baseline/_ref, /root/reference/src or a user's install)
    if not ensure_reference():
        raise ImportError("reference package not found")
    from multi_robot_multi_goal_planning.problems.planning_env import (  # type: ignore
        BaseModeLogic, BaseProblem, DependencyGraphMixin, Mode, SequenceMixin, State, Task, ProblemSpec, AgentType, ConstraintType,
        ManipulationType, DependencyType, DynamicsType, GoalType, SafePoseType, generate_binary_search_indices)
    from multi_robot_multi_goal_planning.problems.core.configuration import (  # type: ignore
        NpConfiguration, batch_config_cost, config_cost, config_dist)
    from multi_robot_multi_goal_planning.problems.core.dependency_graph import DependencyGraph  # type: ignore
    from multi_robot_multi_goal_planning.problems.core.goals import GoalSet, SingleGoal  # type: ignore
    from multi_robot_multi_goal_planning.problems.core.registry import register  # type: ignore
    HAVE_REFERENCE = True
except Exception:  # pragma: no cover - exercised on machines without the reference
    HAVE_REFERENCE = False
    BaseProblem = object  # type: ignore

    def register(_names):  # type: ignore
        return lambda cls: cls


# ----------------------------------------------------------------------------------------------
# devices
# ----------------------------------------------------------------------------------------------
class CudaDevice:
    """numpy-in / numpy-out adapter over backend.SceneBackend for the single-query API; the batch
    API hands CUDA tensors straight through."""

    def __init__(self, max_modes: int = 256, device=None):
        import torch
        from .backend import SceneBackend
        self.torch = torch
        self.be = SceneBackend(max_modes=max_modes, device=device)
        self.dev = self.be.device

    def set_mode(self, slot: int, cs: CompiledScene) -> None:
        self.be.set_mode(slot, cs)

    def _t(self, a, dtype=None):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a
        return t.from_numpy(np.ascontiguousarray(a, np.float32)).to(self.dev)

    def check_configs(self, slot, q, tol=None):
        return self.be.check_configs(slot, self._t(q), tol=tol)

    def check_configs_for_robot(self, slot, q, rel, oth, tol=None):
        return self.be.check_configs_for_robot(slot, self._t(q), rel, oth, tol=tol)

    def check_edges(self, slot, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False, tol=None):
        t = self.torch
        if N is not None and not isinstance(N, t.Tensor):
            N = t.from_numpy(np.ascontiguousarray(N, np.int32)).to(self.dev)
        return self.be.check_edges(slot, self._t(q1), self._t(q2), resolution, N=N, n_start=n_start, n_max=n_max,
                                   include_endpoints=include_endpoints, tol=tol)

    # single-query seam (numpy in, numpy out): one library call does the copies, the launch and the synchronisation
    def query_configs(self, slot, q, tol=None, rel=None, oth=None):
        return self.be.query_configs_host(slot, q, tol, rel, oth)

    def query_edges(self, slot, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False, tol=None):
        return self.be.query_edges_host(slot, q1, q2, resolution, N=N, n_start=n_start, n_max=n_max,
                                        include_endpoints=include_endpoints, tol=tol)

    # asynchronous whole-edge batches (candidate-edge speculation): -> ticket, later (free, first colliding position)
    def prefetch_edges(self, slot, q1, q2, resolution, N=None, include_endpoints=False, tol=None):
        return self.be.submit_edges_host(slot, q1, q2, resolution, N=N, include_endpoints=include_endpoints, tol=tol)

    def collect_edges(self, ticket, E):
        return self.be.collect_edges_host(ticket, E)

    @staticmethod
    def to_numpy(x):
        return x.cpu().numpy() if hasattr(x, "cpu") else np.asarray(x)

    def __deepcopy__(self, memo):
        return self  # a handle to device-resident, append-only state: copies of an environment share it


# ----------------------------------------------------------------------------------------------
# speculative batching behind the single-query API
# ----------------------------------------------------------------------------------------------
class SpeculativeCache:
    """Turns the planners' one-query-at-a-time calls into whole-batch device launches without touching
    planner code, and without changing a single answer:

      * uniform samples: the reference draws them in blocks of 1000 (`_make_uniform_sampler`,
        P/problems/planning_env.py:1697-1708) and the rejection samplers then ask `is_collision_free` row
        by row (P/planners/collision_free_sampler.py:106-112).  The first such query validates the whole
        block in the query's mode with one launch; the other 999 answers come from the cached flags.
      * edges: planners revisit an edge with growing windows `N_start / N_max` over the reference's binary
        order (P/problems/rai_base_env.py:618-676; IT* sparse-then-dense checks, the interleaved path check
        P/problems/planning_env.py:1827-1874).  The first query scans the whole edge once and stores the first
        colliding position p0; a window [a, b) is free iff p0 is none or p0 >= b, collides iff a <= p0 < b,
        and only a window that starts after a known collision goes back to the device.

      * pinned samples: transition sampling overwrites some robots' coordinates of a block row with their goal values
        before asking (`_apply_pinned`, P/planners/collision_free_sampler.py:99-110; composite_prm_planner.py:300-303),
        so the exact-row lookup misses.  The sampler hands rows out in order, so the query is compared with the row
        handed out last: if it differs only in a set of coordinates, the remaining rows of the block are validated with
        the same coordinates overwritten, in one launch -- the following transition samples of that mode hit.
      * candidate edges: the PRM search expands a node by asking for its neighbours and their edge costs
        (`env.batch_config_cost(node.q, neighbour_array)`, P/planners/prm/prm_graph.py:706-712) and afterwards checks
        the candidate edges lazily, one `is_edge_collision_free` call per queue pop (:666-677).  The cost call names
        every candidate edge of the node, so all of them go to the device in ONE asynchronous launch
        (mrb200_submit_edges_host) while the planner pushes them onto its queue; the later single queries are answered
        from the batch's first colliding positions.  A batch issued for the wrong kinematic tree (the node was a
        transition twin) is re-issued once for the right one.

    Keys are the exact fp64 bytes of the configurations, so a configuration that was edited after sampling
    (pinned robots) simply misses the cache and takes the normal path."""

    class Candidates:
        """one node's candidate edges: start q1, ends q2 [E, D], per-edge N, the launch parameters and (once collected)
        the first colliding positions"""
        __slots__ = ("slot", "q2", "N", "params", "ticket", "first", "index")

        def __init__(self, slot, q2, N, params, ticket):
            self.slot, self.q2, self.N, self.params, self.ticket = slot, q2, N, params, ticket
            self.first = None
            self.index = None

        def row_of(self, q2_bytes: bytes) -> Optional[int]:
            if self.index is None:   # built on the first lookup, one vectorised pass
                rows = np.ascontiguousarray(self.q2).view(np.dtype((np.void, self.q2.shape[1] * 8))).ravel().tolist()
                self.index = {}
                for i, r in enumerate(rows):
                    self.index.setdefault(r, i)
            return self.index.get(q2_bytes)

    def __init__(self, max_edges: int = 1 << 18, max_candidate_batches: int = 256):
        self.block: Optional[np.ndarray] = None
        self.block_index: Optional[Dict[bytes, int]] = None
        self.block_flags: Dict[int, np.ndarray] = {}
        self.edges: Dict[tuple, int] = {}
        self.max_edges = max_edges
        # start configuration -> its latest batches (a transition configuration is expanded twice: as the node of the
        # mode it ends and as its twin in the mode it starts, with different neighbour sets)
        self.candidates: Dict[bytes, List["SpeculativeCache.Candidates"]] = {}
        self.max_candidate_batches = max_candidate_batches
        self.cursor = 0                       # index of the block row handed out last
        self.pinned_flags: Dict[tuple, tuple] = {}   # (slot, changed coordinates, their values) -> (first row, flags)
        self.stats = {"config_launches": 0, "config_hits": 0, "pinned_launches": 0, "pinned_hits": 0, "edge_launches": 0, "edge_hits": 0,
                      "candidate_batches": 0, "candidate_edges": 0, "candidate_hits": 0, "candidate_reissues": 0}

    def new_block(self, batch: np.ndarray) -> None:
        self.block = batch
        self.block_index = None     # row lookup table, built on the first query of the block (one vectorised pass)
        self.block_flags = {}
        self.pinned_flags = {}
        self.cursor = 0

    def _block_row(self, q_state: np.ndarray) -> Optional[int]:
        if self.block is None:
            return None
        if self.block_index is None:
            b = np.ascontiguousarray(self.block)
            rows = b.view(np.dtype((np.void, b.shape[1] * 8))).ravel().tolist()
            self.block_index = dict(zip(rows, range(len(rows))))
        return self.block_index.get(np.ascontiguousarray(q_state, np.float64).tobytes())

    def add_candidates(self, q1_bytes: bytes, cand: "SpeculativeCache.Candidates") -> None:
        if len(self.candidates) >= self.max_candidate_batches:
            for k in list(self.candidates)[: self.max_candidate_batches // 2]:   # oldest half (insertion order)
                del self.candidates[k]
        lst = self.candidates.pop(q1_bytes, [])
        self.candidates[q1_bytes] = ([cand] + lst)[:4]
        self.stats["candidate_batches"] += 1
        self.stats["candidate_edges"] += len(cand.N)

    def config_flag(self, slot: int, q_state: np.ndarray, check_batch) -> Optional[bool]:
        """flag of a configuration of the current sample block in mode slot `slot`, or None if it is not one"""
        i = self._block_row(q_state)
        if i is None:
            return self._pinned_flag(slot, q_state, check_batch)
        flags = self.block_flags.get(slot)
        if flags is None:
            flags = np.asarray(check_batch(self.block)).astype(bool)
            self.block_flags[slot] = flags
            self.stats["config_launches"] += 1
        else:
            self.stats["config_hits"] += 1
        return bool(flags[i])

    def _pinned_flag(self, slot: int, q_state: np.ndarray, check_batch) -> Optional[bool]:
        """the row handed out last with some coordinates overwritten (pinned robots)?  Then validate the rest of the block
        with the same overwrite."""
        if self.block is None or not (0 <= self.cursor < len(self.block)):
            return None
        q = np.asarray(q_state, np.float64)
        row = self.block[self.cursor]
        if q.shape != row.shape:
            return None
        diff = q != row
        n = int(diff.sum())
        if n == 0 or n == len(q):
            return None
        key = (slot, diff.tobytes(), q[diff].tobytes())
        hit = self.pinned_flags.get(key)
        if hit is None:
            batch = self.block[self.cursor:].copy()
            batch[:, diff] = q[diff]
            hit = (self.cursor, np.asarray(check_batch(batch)).astype(bool))
            if len(self.pinned_flags) >= 64:
                self.pinned_flags.clear()
            self.pinned_flags[key] = hit
            self.stats["pinned_launches"] += 1
        else:
            self.stats["pinned_hits"] += 1
        first, flags = hit
        return bool(flags[self.cursor - first])

    def edge_window(self, key: tuple, n_start: int, n_max: Optional[int], N: int, full_scan) -> Optional[bool]:
        """answer for window [n_start, n_max) of the edge `key`, or None if only the device can tell"""
        first = self.edges.get(key)
        if first is None:
            first = int(full_scan())     # (the caller counts the launch: a candidate batch may answer without one)
            if len(self.edges) >= self.max_edges:
                self.edges.clear()
            self.edges[key] = first
        else:
            self.stats["edge_hits"] += 1
        hi = N if n_max is None else min(n_max, N)
        if first < 0 or first >= hi:
            return True
        if first >= n_start:
            return False
        return None


# ----------------------------------------------------------------------------------------------
# geometry + modes, independent of the reference package
# ----------------------------------------------------------------------------------------------
class SceneModel:
    """A primitive scene, its per-mode kinematic trees and the device that checks them.

    A mode's tree is the base scene plus a chain of relinks (parent, child, configuration at
    which the child was attached) -- what rai_env.set_to_mode replays
    (P/problems/rai_base_env.py:704-836).  Each distinct tree is compiled once into a blob and
    uploaded into a device slot."""

    def __init__(self, scene: Scene, tol: float, resolution: float, device=None, max_modes: int = 256):
        self.base = scene
        self.tol = float(tol)
        self.resolution = float(resolution)
        self.max_modes = max_modes
        self._device = device
        self._slots: Dict[tuple, int] = {}
        self._by_signature: Dict[tuple, int] = {}
        self._chains: Dict[tuple, Scene] = {}
        self._compiled: Dict[int, CompiledScene] = {}
        self._scenes: Dict[int, Scene] = {}

    @property
    def device(self):
        if self._device is None:
            self._device = CudaDevice(self.max_modes)  # raises without libmrb200.so / a GPU: no CPU fallback
        return self._device

    @staticmethod
    def _chain_key(relinks) -> tuple:
        return tuple((p, c, np.asarray(q, np.float64).tobytes()) for p, c, q in relinks)

    def relinked(self, relinks: Sequence[Tuple[str, str, np.ndarray]]) -> Scene:
        """base scene + the chain of re-parentings; chains are built from their longest known prefix (a mode's chain
        extends its predecessor's by one link), so a new mode costs one frame-table copy and one attach"""
        relinks = list(relinks)
        k = len(relinks)
        while k > 0 and self._chain_key(relinks[:k]) not in self._chains:
            k -= 1
        sc = self._chains[self._chain_key(relinks[:k])] if k else self.base
        for i in range(k, len(relinks)):
            parent, child, q = relinks[i]
            sc = sc.copy()
            sc.attach(parent, child, np.asarray(q, np.float64))
            self._chains[self._chain_key(relinks[:i + 1])] = sc
        return sc

    @staticmethod
    def _signature(sc: Scene, base: Scene) -> tuple:
        """what a chain of re-parentings changed: (object, parent, rounded relative pose, contact) of every frame that
        differs from the base scene.  Two chains with the same signature are the same kinematic tree (the pose of a
        held object depends on the holder's joints only, not on where the other robots stood)."""
        out = []
        for n, f in sc.frames.items():
            b = base.frames[n]
            if f.parent != b.parent or f.rel is not b.rel:
                out.append((n, f.parent, np.round(np.concatenate([f.rel.t, f.rel.R.ravel()]), 7).tobytes(), f.contact))
        return tuple(sorted(out))

    def slot_for(self, key: tuple, relinks: Sequence[Tuple[str, str, np.ndarray]] = ()) -> int:
        if key not in self._slots:
            sc = self.relinked(relinks)
            sig = self._signature(sc, self.base)
            if sig in self._by_signature:         # an equivalent tree is already on the device
                self._slots[key] = self._by_signature[sig]
                return self._slots[key]
            if len(self._compiled) >= self.max_modes:
                raise RuntimeError(f"more than {self.max_modes} distinct kinematic trees; raise max_modes")
            slot = len(self._compiled)
            cs = compile_blob(sc, self.tol)
            self.device.set_mode(slot, cs)
            self._slots[key] = slot
            self._by_signature[sig] = slot
            self._compiled[slot] = cs
            self._scenes[slot] = sc
        return self._slots[key]

    def __deepcopy__(self, memo):
        """The reference deep-copies its environments, one per planner run (P/scripts/run_experiment.py:276,420;
        custom copies in P/problems/rai_base_env.py:337-369).  A SceneModel is a cache of compiled kinematic trees
        keyed by their relink chain -- a pure function of the key, living in slots of one device handle -- so
        copies share it; per-environment state (current mode, speculation cache, sampler) lives in B200Env."""
        return self

    def compiled(self, slot: int) -> CompiledScene:
        return self._compiled[slot]

    def scene(self, slot: int) -> Scene:
        return self._scenes[slot]

    def sample_informed(self, slot: int, n: int, focal_points: np.ndarray, cost_bound: float, metric: str = "euclidean",
                        reduction: str = "max", rng: Optional[np.random.RandomState] = None, max_rounds: int = 64):
        """The planners' informed rejection loop in whole batches (P/planners/composite_prm_planner.py:180-227: a
        uniform sample is kept only if cost(start focus, q) + cost(q, goal focus) <= the incumbent cost, and only
        then collision checked).  Both costs are evaluated on the device for the whole batch
        (mrb200_batch_cost: the reference's per-robot metric and reduction), the ellipse test compacts the batch, and
        only the survivors reach the FK + narrowphase kernel.
        -> (collision-free configurations inside the ellipse [<= n, D] fp64, samples drawn, samples pruned by the ellipse)"""
        import torch
        from . import knn as K
        rng = rng or np.random
        dev = self.device.dev
        gen = torch.Generator(device=dev).manual_seed(int(rng.randint(0, 2 ** 31 - 1)))  # samples are drawn on the device
        lim = self.base.limits()
        lo = torch.from_numpy(lim[0].astype(np.float32)).to(dev)
        width = torch.from_numpy((lim[1] - lim[0]).astype(np.float32)).to(dev)
        rs = self.base.robot_slices()
        sl = [list(rs[r]) for r in self.base.robots]
        f0 = torch.from_numpy(np.asarray(focal_points[0], np.float64)).to(dev)
        f1 = torch.from_numpy(np.asarray(focal_points[1], np.float64)).to(dev)
        out, have, drawn, pruned = [], 0, 0, 0
        batch = max(1024, 4 * n)
        for _ in range(max_rounds):
            qd = lo + width * torch.rand((batch, lim.shape[1]), generator=gen, device=dev, dtype=torch.float32)
            q64 = qd.double()
            cost = K.batch_config_cost(f0, q64, sl, metric, reduction) + K.batch_config_cost(f1, q64, sl, metric, reduction)
            keep = cost <= cost_bound
            drawn += batch
            pruned += int(batch - keep.sum().item())
            cand = qd[keep].contiguous()
            if cand.shape[0]:
                ok = self.device.check_configs(slot, cand).bool()
                good = cand[ok].double().cpu().numpy()
                out.append(good)
                have += len(good)
            if have >= n:
                break
            batch = min(2 * batch, 1 << 22)
        res = np.concatenate(out)[:n] if out else np.zeros((0, lim.shape[1]))
        return res, drawn, pruned

    # batch API (arrays or CUDA tensors in, device results out)
    def check_configs(self, slot, qs, tol=None):
        return self.device.check_configs(slot, qs, tol)

    def check_edges(self, slot, q1s, q2s, resolution=None, **kw):
        return self.device.check_edges(slot, q1s, q2s, self.resolution if resolution is None else resolution, **kw)


if HAVE_REFERENCE:

    class B200Env(BaseProblem):
        """Primitive-scene environment answering the reference's collision API from the GPU.

        Same contract as rai_env (P/problems/rai_base_env.py:258-836): `is_collision_free`
        raises ValueError on q=None like abstract_env.py:256-257; a failing device call raises
        (never reports "free")."""

        def __init__(self, scene: Scene, tol: float, resolution: float, device=None, speculate: bool = True):
            self.model = SceneModel(scene, tol, resolution, device=device)
            self.spec_cache = SpeculativeCache() if speculate else None
            self.scene = scene
            self.robots = list(scene.robots)
            sl = scene.robot_slices()
            self.robot_idx = {r: list(range(sl[r][0], sl[r][1])) for r in self.robots}
            self.robot_dims = {r: sl[r][1] - sl[r][0] for r in self.robots}
            home = scene.home()
            self.start_pos = NpConfiguration.from_list([home[sl[r][0]:sl[r][1]] for r in self.robots])
            self.limits = scene.limits()
            self.collision_tolerance = tol
            self.collision_resolution = resolution
            self.cost_metric = "euclidean"
            self.cost_reduction = "max"
            self.manipulating_env = False
            self.prev_mode = None
            self._slot = None
            self._mask_cache: Dict[tuple, tuple] = {}
            self._slot_by_mode: Dict[Optional[int], int] = {}
            self._last_edge_end: Optional[bytes] = None      # end configuration of the last edge found free (candidate speculation)
            self.max_candidate_edges = 2048
            self.query_stats = {"configs": 0, "robot": 0, "edges": 0, "paths": 0}   # what the planner asked, cached or not
            super().__init__()
            self.spec = ProblemSpec(agent_type=AgentType.MULTI_AGENT, constraints=ConstraintType.UNCONSTRAINED,
                                    manipulation=ManipulationType.MANIPULATION, dependency=DependencyType.FULLY_ORDERED,
                                    dynamics=DynamicsType.GEOMETRIC, goals=GoalType.MULTI_GOAL,
                                    home_pose=SafePoseType.HAS_NO_SAFE_HOME_POSE)

        # ---- visualisation: no-ops -------------------------------------------------------
        def show(self, blocking: bool = False) -> None:
            pass

        def show_config(self, q, blocking: bool = True) -> None:
            pass

        def display_path(self, *a, **k) -> None:
            pass

        # ---- costs: the reference's own kernels (small batches, host) ----------------------
        def config_cost(self, start, end) -> float:
            return config_cost(start, end, self.cost_metric, self.cost_reduction)

        DEVICE_COST_MIN_ROWS = 32768   # below this the reference's own numba kernel beats a device round trip

        def batch_config_cost(self, starts, ends, tmp_agent_slice=None):
            if (tmp_agent_slice is None and isinstance(ends, np.ndarray) and ends.ndim == 2 and len(ends) >= self.DEVICE_COST_MIN_ROWS
                    and hasattr(starts, "state") and hasattr(self.model.device, "dev")):
                # one-to-many cost over a large array (informed-set tests over whole sample batches): mrb200_batch_cost
                import torch
                from . import knn as K
                dev = self.model.device.dev
                sl = [[self.robot_idx[r][0], self.robot_idx[r][-1] + 1] for r in self.robots]
                a = torch.from_numpy(np.ascontiguousarray(starts.state(), np.float64)).to(dev)
                b = torch.from_numpy(np.ascontiguousarray(ends, np.float64)).to(dev)
                return K.batch_config_cost(a, b, sl, self.cost_metric, self.cost_reduction).cpu().numpy()
            costs = batch_config_cost(starts, ends, self.cost_metric, self.cost_reduction, tmp_agent_slice=tmp_agent_slice)
            # candidate-edge speculation (SpeculativeCache): the PRM search asks for the edge costs from the node it has
            # just reached to all of its neighbours (prm_graph.py:706-712) -- those are the edges it may check next
            if (self.spec_cache is not None and self._last_edge_end is not None and tmp_agent_slice is None
                    and isinstance(ends, np.ndarray) and ends.ndim == 2 and len(ends) >= 2 and hasattr(starts, "state")):
                q1 = starts.state()
                if q1.tobytes() == self._last_edge_end:
                    self._prefetch_candidates(q1, ends, costs)
            return costs

        def _prefetch_candidates(self, q1: np.ndarray, ends: np.ndarray, costs: np.ndarray) -> None:
            dev = self.model.device
            submit = getattr(dev, "prefetch_edges", None)
            if submit is None or self._slot is None:
                return
            key = q1.tobytes()
            for old in self.spec_cache.candidates.get(key, ()):
                if old.slot == self._slot and old.q2.shape == ends.shape and np.array_equal(old.q2, ends):
                    return                               # same node expanded again in the same search state
            if len(ends) > self.max_candidate_edges:     # cheapest edges first
                keep = np.argpartition(costs, self.max_candidate_edges)[: self.max_candidate_edges]
                ends = ends[keep]
            q2 = np.ascontiguousarray(ends, np.float64)
            # N exactly like is_edge_collision_free: max(2, int(|dq|_inf / resolution) + 1) in fp64
            N = np.maximum((np.max(np.abs(q2 - q1[None, :]), axis=1) / self.collision_resolution).astype(np.int64) + 1, 2).astype(np.int32)
            params = (float(self.collision_resolution), False, None)
            ticket = submit(self._slot, q1.astype(np.float32), q2.astype(np.float32), self.collision_resolution, N=N)
            self.spec_cache.add_candidates(key, SpeculativeCache.Candidates(self._slot, q2, N, params, ticket))

        def _candidate_first(self, q1_bytes: bytes, q2: np.ndarray, N: int, params: tuple) -> Optional[int]:
            """first colliding position of edge (q1, q2) if it is one of q1's speculated candidate edges, else None"""
            cand, i = None, None
            q2_bytes = q2.tobytes()
            for c in self.spec_cache.candidates.get(q1_bytes, ()):
                if c.params != params:
                    continue
                j = c.row_of(q2_bytes)
                if j is not None and int(c.N[j]) == N:
                    if cand is None or (c.slot == self._slot and cand.slot != self._slot):
                        cand, i = c, j
            if cand is None:
                return None
            dev = self.model.device
            if cand.slot != self._slot:      # speculated for another kinematic tree (transition twin): redo the batch here
                q1 = np.frombuffer(q1_bytes, np.float64)
                free, first = dev.check_edges(self._slot, np.repeat(q1[None].astype(np.float32), len(cand.N), 0),
                                              cand.q2.astype(np.float32), params[0], N=cand.N)
                cand.first = np.array(CudaDevice.to_numpy(first), np.int32)
                cand.slot, cand.ticket = self._slot, None
                self.spec_cache.stats["candidate_reissues"] += 1
            elif cand.first is None:
                try:
                    _, first = dev.collect_edges(cand.ticket, len(cand.N))
                except Exception:      # the ticket has expired (many newer batches since): run the batch again, synchronously
                    q1 = np.frombuffer(q1_bytes, np.float64)
                    _, first = dev.check_edges(self._slot, np.repeat(q1[None].astype(np.float32), len(cand.N), 0),
                                               cand.q2.astype(np.float32), params[0], N=cand.N)
                    first = CudaDevice.to_numpy(first)
                    self.spec_cache.stats["candidate_reissues"] += 1
                cand.first = np.array(first, np.int32)
            self.spec_cache.stats["candidate_hits"] += 1
            return int(cand.first[i])

        # ---- modes -------------------------------------------------------------------------
        def _relinks_for_mode(self, m: "Mode") -> List[Tuple[str, str, np.ndarray]]:
            """Replay of the mode chain, like rai_env.set_to_mode (rai_base_env.py:742-810): at every
            past transition the switching robots sit at the next mode's entry configuration and the
            finished task's frames[1] is attached to frames[0] (unless the task is a plain goto)."""
            chain = []
            cur = m
            while cur is not None:
                chain.append(cur)
                cur = cur.prev_mode
            chain = chain[::-1]
            q = self.scene.home().copy()
            relinks = []
            for mode, nxt in zip(chain[:-1], chain[1:]):
                task = self.get_active_task(mode, nxt.task_ids)
                for r in task.robots:
                    i = self.robots.index(r)
                    q[self.robot_idx[r]] = nxt.entry_configuration[i]
                done_task = self.tasks[mode.task_ids[self.robots.index(task.robots[0])]]
                if done_task.type is not None and done_task.type != "goto":
                    relinks.append((done_task.frames[0], done_task.frames[1], q.copy()))
            return relinks

        def _slot_for_mode(self, m: Optional["Mode"]) -> int:
            if not self.manipulating_env or m is None:
                return self.model.slot_for(())
            relinks = self._relinks_for_mode(m)
            key = tuple((p, c, np.round(q, 9).tobytes()) for p, c, q in relinks)
            return self.model.slot_for(key, relinks)

        def set_to_mode(self, m: "Mode", config=None, use_cached: bool = True, place_in_cache: bool = True):
            if m is self.prev_mode and self._slot is not None:
                return
            # Mode.id is unique per mode object (planning_env.py:139-140): the replay of the mode chain runs once per
            # mode, like the per-mode cache of rai_env.set_to_mode (rai_base_env.py:816-828)
            mid = None if m is None else m.id
            slot = self._slot_by_mode.get(mid)
            if slot is None:
                slot = self._slot_by_mode[mid] = self._slot_for_mode(m)
            self._slot = slot
            self.prev_mode = m

        def get_scenegraph_info_for_mode(self, mode: "Mode", is_start_mode: bool = False):
            """{object: (parent name, rounded relative pose bytes)} (rai_base_env.py:678-702)."""
            if not self.manipulating_env:
                return {}
            sc = self.model.relinked(self._relinks_for_mode(mode))
            from .scene import mat_to_quat
            sg = {}
            for f in sc.frames.values():
                if "obj" in f.name:
                    pose = np.round(np.concatenate([f.rel.t, mat_to_quat(f.rel.R)]), 3)
                    pose = np.where(np.abs(pose) < 1e-6, 0.0, pose)
                    sg[f.name] = (f.parent, pose.tobytes())
            mode._cached_hash = None
            return sg

        # ---- sampling: the reference's block sampler, with the block handed to the speculation cache ----
        def _make_uniform_sampler(self, batch_size=1000):
            """Same random stream and same yielded configurations as BaseProblem._make_uniform_sampler
            (planning_env.py:1697-1708)."""
            while True:
                batch = np.random.uniform(low=self.limits[0, :], high=self.limits[1, :], size=(batch_size, self.limits.shape[1]))
                if self.spec_cache is not None:
                    self.spec_cache.new_block(batch)
                for i in range(batch_size):
                    if self.spec_cache is not None:
                        self.spec_cache.cursor = i
                    yield self.start_pos.from_flat(batch[i])

        # ---- collision queries -------------------------------------------------------------
        def is_collision_free(self, q, m, collision_tolerance: Optional[float] = None) -> bool:
            if q is None:
                raise ValueError
            self.query_stats["configs"] += 1
            self.set_to_mode(m)
            if self.spec_cache is not None and collision_tolerance is None:
                slot = self._slot
                hit = self.spec_cache.config_flag(slot, q.state(), lambda b: CudaDevice.to_numpy(
                    self.model.device.check_configs(slot, b.astype(np.float32))))
                if hit is not None:
                    return hit
            return self._one_config(np.asarray(q.state(), np.float32)[None], collision_tolerance)

        def _one_config(self, q, tol=None, rel=None, oth=None) -> bool:
            """one configuration through the device's host-buffer entry (mrb200_query_configs_host) when it has one"""
            dev = self.model.device
            fast = getattr(dev, "query_configs", None)
            if fast is not None:
                return bool(fast(self._slot, q, tol, rel, oth)[0])
            out = dev.check_configs(self._slot, q, tol) if rel is None else dev.check_configs_for_robot(self._slot, q, rel, oth, tol)
            return bool(CudaDevice.to_numpy(out)[0])

        def _edges(self, a, b, resolution, **kw):
            """(free, first colliding position) of edges as numpy arrays, through mrb200_query_edges_host when available"""
            dev = self.model.device
            fast = getattr(dev, "query_edges", None)
            free, first = (fast or dev.check_edges)(self._slot, a, b, resolution, **kw)
            return CudaDevice.to_numpy(free), CudaDevice.to_numpy(first)

        def is_collision_free_np(self, q, m, collision_tolerance=None, set_mode: bool = True) -> bool:
            self.query_stats["configs"] += 1
            if set_mode:
                self.set_to_mode(m)
            return self._one_config(np.asarray(q, np.float32)[None], collision_tolerance)

        def _robot_masks(self, robots: List[str], m: "Mode"):
            frames = set()
            for r in robots:
                t = self.tasks[m.task_ids[self.robots.index(r)]]
                if t.frames is not None:
                    frames.update(t.frames)
            key = (self._slot, tuple(robots), tuple(sorted(frames)))
            hit = self._mask_cache.get(key)
            if hit is None:   # (the substring tests over all shape names are the same for every query of this kind)
                cs = self.model.compiled(self._slot)
                others = [r for r in self.robots if r not in robots]
                rel = np.array([any(r in n for r in robots) or n in frames for n in cs.shape_names], np.uint8)
                oth = np.array([any(o in n for o in others) for n in cs.shape_names], np.uint8)
                hit = self._mask_cache[key] = (rel, oth)
            return hit

        def is_collision_free_for_robot(self, r, q, m=None, collision_tolerance=None, set_mode: bool = True) -> bool:
            """rai_base_env.py:515-615: free unless the total penetration exceeds the tolerance AND some
            penetrating pair involves robot r (or a frame of its task) and no other robot."""
            if isinstance(r, str):
                r = [r]
            self.query_stats["robot"] += 1
            if set_mode:
                self.set_to_mode(m)
            rel, oth = self._robot_masks(list(r), m)
            return self._one_config(np.asarray(q, np.float32)[None], collision_tolerance, rel, oth)

        def is_edge_collision_free(self, q1, q2, m, resolution=None, tolerance=None, include_endpoints: bool = False,
                                   N_start: int = 0, N_max: Optional[int] = None, N: Optional[int] = None) -> bool:
            """rai_base_env.py:618-676.  N is computed here in fp64 from the fp64 endpoints exactly like the
            reference, then handed to the kernel."""
            if resolution is None:
                resolution = self.collision_resolution
            if N is None:
                N = max(2, int(config_dist(q1, q2, "max") / resolution) + 1)
            if N_start > N:
                assert False
            self.query_stats["edges"] += 1
            s1, s2 = q1.state(), q2.state()
            whole = N_start == 0 and (N_max is None or N_max >= N)
            if N <= 2 and not include_endpoints:   # no interior sample (rai_base_env.py:655-674 loops over nothing): free
                if whole:
                    self._last_edge_end = s2.tobytes()
                return True
            self.set_to_mode(m)
            a, b = np.asarray(s1, np.float32)[None], np.asarray(s2, np.float32)[None]
            Ns = np.array([N], np.int32)
            if self.spec_cache is not None:
                slot = self._slot
                tol_key = None if tolerance is None or tolerance == self.collision_tolerance else float(tolerance)
                b1 = s1.tobytes()
                key = (slot, b1, s2.tobytes(), int(N), float(resolution), bool(include_endpoints), tol_key)

                def full_scan():
                    p0 = self._candidate_first(b1, s2, int(N), (float(resolution), bool(include_endpoints), tol_key))
                    if p0 is not None:
                        return p0
                    self.spec_cache.stats["edge_launches"] += 1
                    return self._edges(a, b, resolution, N=Ns, include_endpoints=include_endpoints, tol=tolerance)[1][0]

                hit = self.spec_cache.edge_window(key, N_start, N_max, int(N), full_scan)
                if hit is not None:
                    if hit and whole:
                        self._last_edge_end = s2.tobytes()
                    return hit
            free, _ = self._edges(a, b, resolution, N=Ns, n_start=N_start, n_max=N_max, include_endpoints=include_endpoints,
                                  tol=tolerance)
            if free[0] and whole:
                self._last_edge_end = s2.tobytes()
            return bool(free[0])

        def is_path_collision_free(self, path, binary_order: bool = True, resolution=None, tolerance=None,
                                   check_edges_in_order: bool = False, check_start_and_end: bool = True) -> bool:
            """Same answer as BaseProblem.is_path_collision_free (planning_env.py:1765-1881) -- every vertex and
            the interior of every edge collision free -- from one vertex batch and one edge batch per mode."""
            if resolution is None:
                resolution = self.collision_resolution
            self.query_stats["paths"] += 1
            L = len(path)
            by_mode: Dict[int, list] = {}
            for i, st in enumerate(path):
                by_mode.setdefault(id(st.mode), [st.mode, [], []])
                # vertices the reference visits (planning_env.py:1791-1824 in-order: every vertex but the first when
                # check_start_and_end is off, the last one always; :1827-1878 interleaved: the last one only with
                # check_start_and_end)
                if i == 0:
                    vertex = check_start_and_end
                elif i == L - 1:
                    vertex = check_edges_in_order or check_start_and_end
                else:
                    vertex = True
                if vertex:
                    by_mode[id(st.mode)][1].append(i)
                if i + 1 < L:
                    by_mode[id(st.mode)][2].append(i)
            for mode, verts, edges in by_mode.values():
                self.set_to_mode(mode)
                if verts:   # vertices are checked at the environment's own tolerance (is_collision_free(q, mode))
                    q = np.stack([np.asarray(path[i].q.state(), np.float32) for i in verts])
                    if not bool(CudaDevice.to_numpy(self.model.device.check_configs(self._slot, q, None)).all()):
                        return False
                if edges:
                    q1 = np.stack([np.asarray(path[i].q.state(), np.float32) for i in edges])
                    q2 = np.stack([np.asarray(path[i + 1].q.state(), np.float32) for i in edges])
                    N = np.array([max(2, int(config_dist(path[i].q, path[i + 1].q, "max") / resolution) + 1) for i in edges], np.int32)
                    free, _ = self.model.device.check_edges(self._slot, q1, q2, resolution, N=N, tol=tolerance)
                    if not bool(CudaDevice.to_numpy(free).all()):
                        return False
            return True

        def sample_valid_uniform_batch(self, mode, n: int, rng: Optional[np.random.RandomState] = None, max_rounds: int = 64):
            """Rejection sampling of n collision-free configurations in `mode`, whole batches at a time
            (the batch form of JointRejectionSampler, P/planners/collision_free_sampler.py:106-112)."""
            rng = rng or np.random
            self.set_to_mode(mode)
            out, have = [], 0
            batch = max(256, 2 * n)
            for _ in range(max_rounds):
                q = rng.uniform(self.limits[0], self.limits[1], (batch, self.limits.shape[1])).astype(np.float32)
                ok = CudaDevice.to_numpy(self.model.device.check_configs(self._slot, q)).astype(bool)
                out.append(q[ok].astype(np.float64))
                have += int(ok.sum())
                if have >= n:
                    break
                batch *= 2
            res = np.concatenate(out)[:n]
            if len(res) < n:
                raise RuntimeError("could not find enough collision-free configurations")
            return res

        def _robot_mask_arrays(self, robot: str, mode):
            self.set_to_mode(mode)
            return self._robot_masks([robot], mode)

        def sample_valid_per_robot_batch(self, mode, n: int, pinned: Optional[Dict[str, np.ndarray]] = None,
                                         rng: Optional[np.random.RandomState] = None, max_per_robot_attempts: int = 100,
                                         oversample: float = 1.5):
            """Batch form of PerRobotRejectionSampler (P/planners/collision_free_sampler.py:115-143): every candidate row
            starts as a uniform sample; robot by robot, the rows whose robot is in collision by the per-robot rule
            (is_collision_free_for_robot, rai_base_env.py:515-615) get that robot's coordinates redrawn -- one device launch
            per robot and attempt over all rows still failing -- then one full check of the surviving rows.
            -> (collision-free configurations [<= n, D] fp64, per-robot checks made)"""
            rng = rng or np.random
            pinned = pinned or {}
            self.set_to_mode(mode)
            dev = self.model.device
            B = max(32, int(n * oversample))
            lim = self.limits
            q = rng.uniform(lim[0], lim[1], (B, lim.shape[1]))
            for r, v in pinned.items():
                q[:, self.robot_idx[r]] = np.asarray(v, np.float64)
            alive = np.ones(B, bool)
            checks = 0
            for r in self.robots:
                if r in pinned:
                    continue
                rel, oth = self._robot_masks([r], mode)
                idx = self.robot_idx[r]
                todo = np.nonzero(alive)[0]
                for _ in range(max_per_robot_attempts):
                    if not len(todo):
                        break
                    ok = CudaDevice.to_numpy(dev.check_configs_for_robot(self._slot, q[todo].astype(np.float32), rel, oth)).astype(bool)
                    checks += len(todo)
                    todo = todo[~ok]
                    q[np.ix_(todo, idx)] = rng.uniform(lim[0, idx], lim[1, idx], (len(todo), len(idx)))
                alive[todo] = False                      # attempts exhausted (the reference gives such a sample up)
            rows = np.nonzero(alive)[0]
            if len(rows):
                ok = CudaDevice.to_numpy(dev.check_configs(self._slot, q[rows].astype(np.float32))).astype(bool)
                rows = rows[ok]
            return q[rows][:n].astype(np.float32).astype(np.float64), checks

        def sample_valid_gibbs_batch(self, mode, n: int, seed_q: Optional[np.ndarray] = None, pinned: Optional[Dict[str, np.ndarray]] = None,
                                     rng: Optional[np.random.RandomState] = None, sweeps: int = 1, max_per_robot_attempts: int = 100,
                                     oversample: float = 1.5):
            """Batch form of GibbsSampler (collision_free_sampler.py:146-195): all rows start at the seed configuration
            (default: the start pose); robot by robot, each row proposes new coordinates for that robot until the WHOLE
            configuration is collision free (or the attempts run out) -- one launch per robot and attempt over the rows
            still proposing; a final full check validates the rows.  -> (configurations [<= n, D] fp64, checks made)"""
            rng = rng or np.random
            pinned = pinned or {}
            self.set_to_mode(mode)
            dev = self.model.device
            B = max(32, int(n * oversample))
            lim = self.limits
            q = np.tile(np.asarray(self.start_pos.state() if seed_q is None else seed_q, np.float64), (B, 1))
            for r, v in pinned.items():
                q[:, self.robot_idx[r]] = np.asarray(v, np.float64)
            checks = 0
            for _ in range(sweeps):
                for r in self.robots:
                    if r in pinned:
                        continue
                    idx = self.robot_idx[r]
                    todo = np.arange(B)
                    for _ in range(max_per_robot_attempts):
                        if not len(todo):
                            break
                        q[np.ix_(todo, idx)] = rng.uniform(lim[0, idx], lim[1, idx], (len(todo), len(idx)))
                        ok = CudaDevice.to_numpy(dev.check_configs(self._slot, q[todo].astype(np.float32))).astype(bool)
                        checks += len(todo)
                        todo = todo[~ok]
            ok = CudaDevice.to_numpy(dev.check_configs(self._slot, q.astype(np.float32))).astype(bool)
            return q[ok][:n].astype(np.float32).astype(np.float64), checks + B

        def sample_valid_informed_batch(self, mode, n: int, focal_points: np.ndarray, cost_bound: float,
                                        rng: Optional[np.random.RandomState] = None):
            """informed rejection sampling in whole batches, see SceneModel.sample_informed"""
            self.set_to_mode(mode)
            return self.model.sample_informed(self._slot, n, focal_points, cost_bound, self.cost_metric, self.cost_reduction, rng)

        # ---- distances / neighbours (numpy or CUDA tensors in; see neighbours.py for the planners' own call) ----
        def _slices(self):
            return [[self.robot_idx[r][0], self.robot_idx[r][-1] + 1] for r in self.robots]

        def _f64(self, a):
            import torch
            if isinstance(a, torch.Tensor):
                return a
            return torch.from_numpy(np.ascontiguousarray(a, np.float64)).to(self.model.device.dev)

        def batch_config_dist(self, q, pts, metric: str = "max"):
            """one-to-many distance on the device (P/problems/core/configuration.py:303-349), numpy [N] float64"""
            from . import knn as K
            qs = q.state() if hasattr(q, "state") else q
            return K.batch_config_dist(self._f64(qs), self._f64(pts), self._slices(), metric).cpu().numpy()

        def batch_knn(self, queries, corpus, k: int, metric: str = "max_euclidean", mode: str = "auto"):
            """k nearest corpus rows of every query row -> (idx [Q, k] int32, dist [Q, k] float64) device tensors;
            the reference's selection (P/planners/prm/prm_graph.py:440-447) for whole batches"""
            from . import knn as K
            return K.batch_knn(self._f64(queries), self._f64(corpus), self._slices(), metric, k, mode=mode)

        def batch_radius(self, queries, corpus, radius, metric: str = "max_euclidean", inclusive: bool = False, mode: str = "auto"):
            """r-disc neighbours in CSR form (offsets [Q + 1], indices), ascending indices per row
            (prm_graph.py:479-538: d < r; rrtstar_base.py / itstar_base.py: d <= r + 1e-10 with inclusive=True)"""
            from . import knn as K
            return K.batch_radius(self._f64(queries), self._f64(corpus), radius, self._slices(), metric, inclusive=inclusive, mode=mode)

        # ---- additive batch variants (arrays or CUDA tensors in, device tensors out) -----------
        def batch_is_collision_free(self, qs, mode):
            self.set_to_mode(mode)
            return self.model.check_configs(self._slot, qs)

        def batch_is_edge_collision_free(self, q1s, q2s, mode, resolution=None, N_start: int = 0, N_max=None,
                                         include_endpoints: bool = False):
            self.set_to_mode(mode)
            return self.model.check_edges(self._slot, q1s, q2s, resolution, n_start=N_start, n_max=N_max,
                                          include_endpoints=include_endpoints)

    def _two_dim_handover_tasks(env: "B200Env"):
        """Task list of rai.2d_handover (rai_envs.py:502-543) with hand-placed keyframes: the
        reference solves them with rai's KOMO at construction time (rai_config.py:844-968), which is
        not available; these are equivalent in construction (touching pick / handover / place poses,
        all collision free in their modes), not identical to any rai instance."""
        k_pick1 = np.array([0.0, 0.77, 0.0])          # a1 just north of obj1 (0, .4)
        k_hand = np.array([-1.2, 1.37, 0.0, -1.2, 0.58, 0.0])  # a1 carries obj1 to (-1.2, 1.0); a2 waits 2 cm south of it
        k_place = np.array([1.22, 0.4, np.pi / 2])    # a2 (obj1 hanging 0.42 'north' of it) drops obj1 at (0.8, 0.4): goal1
        k_pick2 = np.array([0.5, -1.13, 0.0])         # a1 north of obj2 (.5, -1.5)
        k_place2 = np.array([1.3, 1.57, 0.0])         # a1 carries obj2 to (1.3, 1.2)
        terminal = np.concatenate([env.start_pos.state()])
        return [
            Task("a1_pick_obj1", ["a1"], SingleGoal(k_pick1), type="pick", frames=["a1", "obj1"]),
            Task("handover", ["a1", "a2"], GoalSet([k_hand]), type="hanover", frames=["a2", "obj1"]),
            Task("a2_place", ["a2"], SingleGoal(k_place), type="place", frames=["table", "obj1"]),
            Task("a1_pick_obj2", ["a1"], SingleGoal(k_pick2), type="pick", frames=["a1", "obj2"]),
            Task("a1_place_obj2", ["a1"], SingleGoal(k_place2), type="place", frames=["table", "obj2"]),
            Task("terminal", ["a1", "a2"], GoalSet([terminal])),
        ]

    @register("b200.2d_handover")
    class b200_two_dim_handover(SequenceMixin, B200Env):
        """B200 counterpart of rai.2d_handover (rai_envs.py:462-573)."""

        def __init__(self, device=None, speculate: bool = True):
            mk, kw = SCENES["2d_handover"]
            B200Env.__init__(self, mk(), kw["tol"], kw["resolution"], device=device, speculate=speculate)
            self.manipulating_env = True
            self.tasks = _two_dim_handover_tasks(self)
            self.sequence = self._make_sequence_from_names(
                ["a1_pick_obj1", "handover", "a1_pick_obj2", "a1_place_obj2", "a2_place", "terminal"])
            BaseModeLogic.__init__(self)
            self.prev_mode = None

    def _manipulation_env(scene_name: str, default_moves: int):
        """Pick / place sequence problems on the named arm scenes: B200 counterparts of rai.box_rearrangement
        (rai_envs.py:1454-1594) and rai.box_stacking (rai_envs.py:1916-1966).  Same scene, same kind of task list
        (pick: object re-parented to the tool frame; place: back to the table; final task: all robots home); the
        goal keyframes come from problems.py / keyframes.py instead of rai's KOMO (see there)."""

        class _Env(SequenceMixin, B200Env):
            def __init__(self, device=None, speculate: bool = True, n_moves: int = default_moves, seed: int = 0):
                from .problems import manipulation_tasks
                mk, kw = SCENES[scene_name]
                B200Env.__init__(self, mk(), kw["tol"], kw["resolution"], device=device, speculate=speculate)
                self.manipulating_env = True
                specs = manipulation_tasks(scene_name, self.model, n_moves=n_moves, seed=seed)
                self.tasks = [Task(t.name, list(t.robots), SingleGoal(t.goal), type=t.type,
                                   frames=None if t.frames is None else list(t.frames)) for t in specs]
                self.sequence = self._make_sequence_from_names([t.name for t in specs])
                BaseModeLogic.__init__(self)
                self.prev_mode = None

        _Env.__name__ = f"b200_{scene_name}"
        return _Env

    b200_box_rearrangement = register("b200.box_rearrangement")(_manipulation_env("box_rearrangement", 4))
    b200_box_stacking = register("b200.box_stacking")(_manipulation_env("box_stacking", 4))

    @register("b200.dep_mobile_wall_four")
    class b200_dep_mobile_wall_four(DependencyGraphMixin, B200Env):
        """B200 counterpart of rai.dep_mobile_wall_four (rai_envs.py:2206-2276): every mobile manipulator carries the two
        boxes of its wall column to the goal wall; a robot's tasks are chained, robots are independent of each other,
        the terminal task depends on all of them (same dependency graph as :2224-2258)."""

        def __init__(self, device=None, speculate: bool = True, num_robots: int = 4, seed: int = 0):
            from .problems import manipulation_tasks
            mk, kw = SCENES["mobile_wall_four"]
            B200Env.__init__(self, mk(num_robots) if num_robots != 4 else mk(), kw["tol"], kw["resolution"], device=device,
                             speculate=speculate)
            self.manipulating_env = True
            specs = manipulation_tasks("mobile_wall_four", self.model, n_moves=2 * num_robots, seed=seed)
            self.graph = DependencyGraph()
            self.tasks = []
            prev = {}
            for t in specs[:-1]:
                r = t.robots[0]
                self.tasks.append(Task(t.name, [r], SingleGoal(t.goal), type=t.type, frames=list(t.frames)))
                if r in prev:
                    self.graph.add_dependency(t.name, prev[r])
                prev[r] = t.name
            for r in self.robots:
                self.graph.add_dependency("terminal", prev[r])
            self.tasks.append(Task("terminal", list(self.robots), SingleGoal(self.start_pos.state())))
            BaseModeLogic.__init__(self)
            self.prev_mode = None
            self.spec.dependency = DependencyType.UNORDERED

    def _goto_env(scene_name: str, seed: int):
        """Geometry of a named scene with plain goto tasks: every robot moves to a sampled collision-free
        goal, then all return home.  (The reference's pick / place keyframes come from rai's KOMO.)"""

        class _Env(SequenceMixin, B200Env):
            def __init__(self, device=None, speculate: bool = True):
                mk, kw = SCENES[scene_name]
                B200Env.__init__(self, mk(), kw["tol"], kw["resolution"], device=device, speculate=speculate)
                rng = np.random.RandomState(seed)
                lim = self.limits
                slot = self.model.slot_for(())
                goal = None
                for _ in range(200):
                    cand = rng.uniform(lim[0], lim[1], (256, lim.shape[1]))
                    ok = CudaDevice.to_numpy(self.model.check_configs(slot, cand.astype(np.float32)))
                    if ok.any():
                        goal = cand[int(np.argmax(ok))].astype(np.float32).astype(np.float64)
                        break
                if goal is None:
                    raise RuntimeError("no collision-free goal found")
                self.tasks = [Task(f"{r}goal", [r], SingleGoal(goal[self.robot_idx[r]])) for r in self.robots]
                self.tasks.append(Task("terminal", list(self.robots), SingleGoal(self.start_pos.state())))
                self.sequence = self._make_sequence_from_names([t.name for t in self.tasks])
                BaseModeLogic.__init__(self)
                self.prev_mode = None

        _Env.__name__ = f"b200_{scene_name}_goto"
        return _Env

    b200_box_rearrangement_goto = register("b200.box_rearrangement_goto")(_goto_env("box_rearrangement", 1))
    b200_box_stacking_goto = register("b200.box_stacking_goto")(_goto_env("box_stacking", 2))
    b200_mobile_wall_four_goto = register("b200.mobile_wall_four_goto")(_goto_env("mobile_wall_four", 3))

    # ---- abstract.test answered by the CUDA abstract kernels (bit-exact with the reference) ----
    from multi_robot_multi_goal_planning.problems.abstract_env import (  # type: ignore
        Rectangle, Sphere, abstract_env_two_dim_middle_obs)

    class AbstractCudaDevice:
        def __init__(self, n_agents, dim, radii, spheres, rects_minmax):
            import torch
            from .backend import AbstractBackend
            self.torch = torch
            self.be = AbstractBackend(n_agents, dim, radii, spheres, rects_minmax)

        def __deepcopy__(self, memo):
            return self  # immutable device-side scene: environment copies share it

        SMALL = 64   # queries up to this size take the host-buffer entry points (one library call, no torch tensors)

        def check_configs(self, q):
            t = self.torch
            q = np.ascontiguousarray(q, np.float64)
            if len(q) <= self.SMALL:
                return self.be.query_configs_host(q)
            return self.be.check_configs(t.from_numpy(q).cuda()).cpu().numpy()

        def check_edges(self, q1, q2, resolution, N=None, n_start=0, n_max=None, include_endpoints=False):
            t = self.torch
            if len(q1) <= self.SMALL:
                return self.be.query_edges_host(q1, q2, resolution, N=N, n_start=n_start, n_max=n_max, include_endpoints=include_endpoints)
            Nt = None if N is None else t.from_numpy(np.ascontiguousarray(N, np.int32)).cuda()
            f, p = self.be.check_edges(t.from_numpy(np.ascontiguousarray(q1, np.float64)).cuda(),
                                       t.from_numpy(np.ascontiguousarray(q2, np.float64)).cuda(), resolution, N=Nt, n_start=n_start,
                                       n_max=n_max, include_endpoints=include_endpoints)
            return f.cpu().numpy(), p.cpu().numpy()

    @register("b200.abstract_test")
    class b200_abstract_test(abstract_env_two_dim_middle_obs):
        """abstract.test (abstract_env.py:381-421) with every collision query answered by the fp64 CUDA
        kernels; flags are bit-identical, so planners behave exactly as on the reference's own env."""

        def __init__(self, device=None, speculate: bool = True):
            self.spec_cache = SpeculativeCache() if speculate else None
            super().__init__()
            self._device = device

        def _make_uniform_sampler(self, batch_size=1000):
            """planning_env.py:1697-1708, plus the block hand-off to the speculation cache"""
            while True:
                batch = np.random.uniform(low=self.limits[0, :], high=self.limits[1, :], size=(batch_size, self.limits.shape[1]))
                if self.spec_cache is not None:
                    self.spec_cache.new_block(batch)
                for i in range(batch_size):
                    if self.spec_cache is not None:
                        self.spec_cache.cursor = i
                    yield self.start_pos.from_flat(batch[i])

        @property
        def device(self):
            if self._device is None:
                sph = [(o.pos, o.radius) for o in self.obstacles if isinstance(o, Sphere)]
                rect = [(o.min_bounds, o.max_bounds) for o in self.obstacles if isinstance(o, Rectangle)]
                dim = len(self.start_pos[0])
                self._device = AbstractCudaDevice(self.start_pos.num_agents(), dim, self.agent_radii, sph, rect)
            return self._device

        def is_collision_free(self, q, mode):
            if q is None:
                raise ValueError
            if self.spec_cache is not None:
                hit = self.spec_cache.config_flag(0, q.state(), lambda b: self.device.check_configs(b))
                if hit is not None:
                    return hit
            return bool(self.device.check_configs(q.state()[None])[0])

        def is_edge_collision_free(self, q1, q2, mode, resolution=None, tolerance=None, include_endpoints=False, N_start=0,
                                   N_max=None, N=None):
            if resolution is None:
                resolution = self.collision_resolution
            if N is None:
                N = max(2, int(config_dist(q1, q2, "max") / resolution) + 1)
            if N_start > N:
                assert False
            Ns = np.array([N], np.int32)
            if self.spec_cache is not None:  # (the abstract env ignores `tolerance`, abstract_env.py:301-354)
                key = (0, q1.state().tobytes(), q2.state().tobytes(), int(N), float(resolution), bool(include_endpoints))
                def full_scan():
                    self.spec_cache.stats["edge_launches"] += 1
                    return self.device.check_edges(q1.state()[None], q2.state()[None], resolution, N=Ns,
                                                   include_endpoints=include_endpoints)[1][0]

                hit = self.spec_cache.edge_window(key, N_start, N_max, int(N), full_scan)
                if hit is not None:
                    return hit
            f, _ = self.device.check_edges(q1.state()[None], q2.state()[None], resolution, N=Ns,
                                           n_start=N_start, n_max=N_max, include_endpoints=include_endpoints)
            return bool(f[0])

        def batch_is_collision_free(self, qs, mode=None):
            return self.device.check_configs(np.asarray(qs, np.float64))

        def batch_is_edge_collision_free(self, q1s, q2s, mode=None, resolution=None, **kw):
            return self.device.check_edges(np.asarray(q1s, np.float64), np.asarray(q2s, np.float64),
                                           self.collision_resolution if resolution is None else resolution, **kw)
