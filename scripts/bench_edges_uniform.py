"""uniform-endpoint edges (the bench's edge metric) for the four scenes: python scripts/bench_edges_uniform.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
be = SceneBackend(max_modes=4)
for name in ["box_rearrangement", "box_stacking", "mobile_wall_four", "2d_handover"]:
    mk, kw = SCENES[name]
    sc = mk(); cs = S.compile_blob(sc, kw["tol"]); be.set_mode(0, cs)
    lim = sc.limits(); E = 100_000 if name == "2d_handover" else 16_384
    q1 = torch.from_numpy(np.random.RandomState(8).uniform(lim[0], lim[1], (E, sc.dof)).astype(np.float32)).cuda()
    q2 = torch.from_numpy(np.random.RandomState(9).uniform(lim[0], lim[1], (E, sc.dof)).astype(np.float32)).cuda()
    be.check_edges(0, q1, q2, kw["resolution"]); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): f, _ = be.check_edges(0, q1, q2, kw["resolution"])
    b.record(); b.synchronize()
    print(f"{name:18s} uniform edges: {E / (a.elapsed_time(b) / 10 * 1e-3):.3e} edges/s, free {f.float().mean().item():.4f}")
