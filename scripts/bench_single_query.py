"""Latency of the single-query seam the reference's planners use (one configuration / one edge per call, numpy in,
bool out) through env.CudaDevice: python scripts/bench_single_query.py"""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200.env import CudaDevice, SceneModel
from multirobot_pathplanning_benchmark_b200.scenes import SCENES

for name in ("2d_handover", "box_rearrangement", "box_stacking"):
    mk, kw = SCENES[name]
    model = SceneModel(mk(), kw["tol"], kw["resolution"])
    slot = model.slot_for(())
    sc = model.base
    lim = sc.limits()
    rng = np.random.RandomState(0)
    qs = rng.uniform(lim[0], lim[1], (2000, sc.dof))
    dev = model.device
    for _ in range(50):
        CudaDevice.to_numpy(dev.check_configs(slot, qs[0].astype(np.float32)[None]))
    t = time.perf_counter()
    n_free = 0
    for q in qs:
        n_free += bool(CudaDevice.to_numpy(dev.check_configs(slot, np.asarray(q, np.float32)[None]))[0])
    tc = (time.perf_counter() - t) / len(qs)
    q2 = qs + rng.uniform(-0.3, 0.3, qs.shape)
    t = time.perf_counter()
    for a, b in zip(qs[:1000], q2[:1000]):
        N = max(2, int(np.max(np.abs(a - b)) / kw["resolution"]) + 1)
        f, _ = dev.check_edges(slot, np.asarray(a, np.float32)[None], np.asarray(b, np.float32)[None], kw["resolution"], N=np.array([N], np.int32))
        bool(CudaDevice.to_numpy(f)[0])
    te = (time.perf_counter() - t) / 1000
    # the same through the host-buffer entry points of the C ABI (what B200Env uses for single queries)
    t = time.perf_counter()
    for q in qs:
        bool(dev.query_configs(slot, np.asarray(q, np.float32)[None])[0])
    tch = (time.perf_counter() - t) / len(qs)
    t = time.perf_counter()
    for a, b in zip(qs[:1000], q2[:1000]):
        N = max(2, int(np.max(np.abs(a - b)) / kw["resolution"]) + 1)
        bool(dev.query_edges(slot, np.asarray(a, np.float32)[None], np.asarray(b, np.float32)[None], kw["resolution"], N=np.array([N], np.int32))[0][0])
    teh = (time.perf_counter() - t) / 1000
    print(f"{name:18s} one configuration: {tc * 1e6:6.1f} us via torch tensors, {tch * 1e6:6.1f} us via mrb200_query_configs_host;  "
          f"one edge: {te * 1e6:6.1f} us / {teh * 1e6:6.1f} us   ({n_free} of {len(qs)} free)")
