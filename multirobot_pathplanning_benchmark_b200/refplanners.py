"""Runs the reference's own, unmodified planners on a `b200.*` environment and reports what BASELINE.json's third
metric asks for: time-to-first-solution (`info["times"][0]`, like P/scripts/run_planner.py:47-346 and the experiment
runner P/scripts/run_experiment.py).

The planner classes come from the reference package (refimport.py: a user's install, the offline install under
baseline/_ref, or the build container's checkout); nothing of them is restated here.  The environment's device is
whatever the caller hands in: the CUDA device (default, product path) or -- only from tests/ and bench.py's CPU arm --
an oracle-backed CPU device with the same four methods.  `DeviceMeter` wraps either and records how many round trips
the planner made and how long the backend was busy, which is what separates backend speed-up from the planner's own
Python time in the reported ratios.
"""
from __future__ import annotations

import contextlib
import io
import random
import time
from typing import Any, Dict, Optional

import numpy as np

from .refimport import ensure_reference

PLANNERS = ("composite_prm", "rrt_star", "birrt_star", "aitstar", "eitstar")


def make_planner(env, name: str, distance_metric: str = "max_euclidean", with_mode_validation: Optional[bool] = None, **options):
    """The planner objects of P/scripts/run_planner.py:160-196 with their default configurations."""
    if not ensure_reference():
        raise RuntimeError("the reference package is not available (see refimport.py)")
    from multi_robot_multi_goal_planning.planners.composite_prm_planner import CompositePRM, CompositePRMConfig
    from multi_robot_multi_goal_planning.planners.itstar_base import BaseITConfig
    from multi_robot_multi_goal_planning.planners.planner_aitstar import AITstar
    from multi_robot_multi_goal_planning.planners.planner_birrtstar import BidirectionalRRTstar
    from multi_robot_multi_goal_planning.planners.planner_eitstar import EITstar
    from multi_robot_multi_goal_planning.planners.planner_rrtstar import RRTstar
    from multi_robot_multi_goal_planning.planners.rrtstar_base import BaseRRTConfig
    if name == "composite_prm":
        cfg = CompositePRMConfig(**options)
        cfg.distance_metric = distance_metric
        if with_mode_validation is not None:
            cfg.with_mode_validation = with_mode_validation
        return CompositePRM(env, cfg)
    if name in ("rrt_star", "birrt_star"):
        cfg = BaseRRTConfig(**options)
        cfg.distance_metric = distance_metric
        if with_mode_validation is not None:
            cfg.with_mode_validation = with_mode_validation
        return (RRTstar if name == "rrt_star" else BidirectionalRRTstar)(env, config=cfg)
    if name in ("aitstar", "eitstar"):
        cfg = BaseITConfig(**options)
        cfg.distance_metric = distance_metric
        if with_mode_validation is not None:
            cfg.with_mode_validation = with_mode_validation
        return (AITstar if name == "aitstar" else EITstar)(env, config=cfg)
    raise ValueError(f"unknown planner {name!r}; one of {PLANNERS}")


class DeviceMeter:
    """Pass-through wrapper of a scene / abstract device: counts calls and items, accumulates wall time spent inside
    the device (launch + synchronisation for CUDA; the computation itself for a CPU device)."""

    METHODS = ("check_configs", "check_configs_for_robot", "check_edges", "query_configs", "query_edges", "prefetch_edges",
               "collect_edges")

    def __init__(self, inner):
        self.inner = inner
        self.calls: Dict[str, int] = {}
        self.items: Dict[str, int] = {}
        self.seconds = 0.0
        for m in self.METHODS:
            if hasattr(inner, m):
                setattr(self, m, self._wrap(m, getattr(inner, m)))

    def _wrap(self, name, fn):
        def call(*a, **k):
            t = time.perf_counter()
            try:
                return fn(*a, **k)
            finally:
                self.seconds += time.perf_counter() - t
                self.calls[name] = self.calls.get(name, 0) + 1
                n = 0
                for x in a[:3]:
                    if hasattr(x, "shape") and len(getattr(x, "shape")) == 2:
                        n = int(x.shape[0])
                        break
                self.items[name] = self.items.get(name, 0) + n
        return call

    def __getattr__(self, name):   # set_mode, dev, to_numpy, be, ...
        return getattr(self.inner, name)

    def __deepcopy__(self, memo):
        return self

    def snapshot(self):
        return {"calls": dict(self.calls), "items": dict(self.items), "seconds": self.seconds}


def run_planner(env, planner_name: str, seed: int, max_time: float, optimize: bool = False, quiet: bool = True,
                **planner_kw) -> Dict[str, Any]:
    """One planner run with the reference's seeding (run_planner.py:145-146, 198-199).  -> dict with the first-solution
    time / cost (`info["times"][0]`, `info["costs"][0]`), the final cost and the wall time of `plan`."""
    from multi_robot_multi_goal_planning.planners.termination_conditions import RuntimeTerminationCondition
    np.random.seed(seed)
    random.seed(seed)
    planner = make_planner(env, planner_name, **planner_kw)
    np.random.seed(seed)
    random.seed(seed)
    sink = io.StringIO()
    t = time.perf_counter()
    with (contextlib.redirect_stdout(sink) if quiet else contextlib.nullcontext()):
        path, info = planner.plan(RuntimeTerminationCondition(max_time), optimize=optimize)
    wall = time.perf_counter() - t
    times, costs = list(info.get("times", [])), list(info.get("costs", []))
    out = {"planner": planner_name, "seed": seed, "solved": path is not None, "wall_s": wall,
           "ttfs_s": float(times[0]) if times else None, "first_cost": float(costs[0]) if costs else None,
           "final_cost": float(costs[-1]) if costs else None, "n_solutions": max(len(costs) - 1, 0) if costs else 0}
    out["_path"] = path
    return out


def metered_env(env_factory, device):
    """environment built on a metered device -> (env, meter)"""
    meter = DeviceMeter(device)
    return env_factory(device=meter), meter
