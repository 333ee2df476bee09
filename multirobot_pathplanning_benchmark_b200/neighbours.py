"""The planners' distance / neighbour calls on the device, behind the reference's own function signature.

The reference's planners never call the environment for distances: they bind the module-level function
`batch_config_dist(pt, batch_pts, metric)` (P/problems/core/configuration.py:342-349) into lambdas
(P/planners/composite_prm_planner.py:606-610, planner_aitstar.py:346-350, planner_eitstar.py:365-369) or call it directly
(P/planners/rrtstar_base.py:1246-1337) with `batch_pts` = the per-mode numpy array of all node configurations
(P/planners/prm/prm_graph.py:399-402 caches it per mode).  `DeviceBatchDist` is a drop-in for that function: numpy in,
numpy out, same metric names; the per-mode arrays are kept RESIDENT on the device (keyed by the identity and shape of the
array the planner passes: it passes the same cached object until it adds nodes), so a query moves one configuration in
and N distances out.  Below `min_rows` the reference's own numba kernel answers: one device round trip (~40 us) loses
against a 5 us host pass over a few thousand rows -- the planners' graphs at time-to-first-solution sizes -- and wins by an
order of magnitude at the 100k rows per mode of BASELINE config 4.

`install()` swaps the function in every planner module of the reference (module attribute patch, nothing on disk is
touched); `uninstall()` restores it.  Whole-batch neighbour search for batch-native callers is knn.batch_knn /
knn.batch_radius (tcgen05 candidate generator); B200Env exposes both with numpy arguments."""
from __future__ import annotations

import importlib
from collections import OrderedDict
from typing import Optional

import numpy as np

PLANNER_MODULES = ("composite_prm_planner", "prm.prm_graph", "prm_static_env", "planner_aitstar", "planner_eitstar", "rrtstar_base",
                   "prioritized_planner", "itstar_base")


class DeviceBatchDist:
    def __init__(self, reference_fn, device=None, min_rows: int = 16384, max_arrays: int = 16):
        self.reference_fn = reference_fn
        self.device = device
        self.min_rows = int(min_rows)
        self.max_arrays = int(max_arrays)
        self.cache: "OrderedDict[tuple, tuple]" = OrderedDict()   # (id, shape, data ptr) -> (device tensor, keep-alive array)
        self.stats = {"device_calls": 0, "host_calls": 0, "uploads": 0}

    def _resident(self, arr: np.ndarray):
        import torch
        key = (id(arr), arr.shape, arr.ctypes.data)
        hit = self.cache.get(key)
        if hit is not None:
            self.cache.move_to_end(key)
            return hit[0]
        dev = self.device or torch.device("cuda", torch.cuda.current_device())
        t = torch.from_numpy(np.ascontiguousarray(arr, np.float64)).to(dev)
        self.cache[key] = (t, arr)          # the array is kept alive: its id cannot be reused while the entry exists
        self.stats["uploads"] += 1
        while len(self.cache) > self.max_arrays:
            self.cache.popitem(last=False)
        return t

    def __call__(self, pt, batch_pts, metric: str = "max"):
        if not isinstance(batch_pts, np.ndarray) or batch_pts.ndim != 2 or len(batch_pts) < self.min_rows or not hasattr(pt, "_array_slice"):
            self.stats["host_calls"] += 1
            return self.reference_fn(pt, batch_pts, metric)
        import torch
        from . import knn as K
        corpus = self._resident(batch_pts)
        q = torch.from_numpy(np.ascontiguousarray(pt.state(), np.float64)).to(corpus.device)
        self.stats["device_calls"] += 1
        return K.batch_config_dist(q, corpus, np.asarray(pt._array_slice, np.int32), metric).cpu().numpy()


_installed: Optional[dict] = None


def install(min_rows: int = 16384, device=None) -> DeviceBatchDist:
    """Route every planner module's `batch_config_dist` through the device for large per-mode arrays."""
    global _installed
    from .refimport import ensure_reference
    if not ensure_reference():
        raise RuntimeError("the reference package is not available")
    from multi_robot_multi_goal_planning.problems.core import configuration as C
    if _installed is not None:
        uninstall()
    fn = DeviceBatchDist(C.batch_config_dist, device=device, min_rows=min_rows)
    patched = {}
    for name in PLANNER_MODULES:
        try:
            mod = importlib.import_module("multi_robot_multi_goal_planning.planners." + name)
        except Exception:   # optional planners with missing third-party imports
            continue
        if getattr(mod, "batch_config_dist", None) is not None:
            patched[mod] = mod.batch_config_dist
            mod.batch_config_dist = fn
    _installed = {"fn": fn, "patched": patched}
    return fn


def uninstall() -> None:
    global _installed
    if _installed is None:
        return
    for mod, orig in _installed["patched"].items():
        mod.batch_config_dist = orig
    _installed = None
