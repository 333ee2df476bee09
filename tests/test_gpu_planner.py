"""Batch-native PRM (planner.py) on the GPU: the CUDA backend and the CPU oracle backend answer the same batch
calls, so with the same seed the planner must find the same plan (BASELINE metric: time-to-first-solution)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("scene,n0,t0,n_moves", [("2d_handover", 400, 50, 0), ("box_rearrangement", 1000, 100, 1), ("box_stacking", 800, 80, 1)])
def test_same_seed_same_plan_on_both_backends(cuda_lib, scene, n0, t0, n_moves):
    """pick / place sequences with held objects (problems.py); the keyframes themselves are solved against each
    backend's own collision answers and must coincide too"""
    sys.path.insert(0, os.path.join(ROOT, "scripts"))
    import ttfs
    ttfs.run("2d_handover", "b200", 99, 200, 30, 30)   # warm-up: CUDA context, module load
    gpu = ttfs.run(scene, "b200", 0, n0, t0, 120, n_moves=n_moves)
    cpu = ttfs.run(scene, "cpu", 0, n0, t0, 300, n_moves=n_moves)
    assert gpu["solved"] and cpu["solved"]
    assert abs(gpu["cost"] - cpu["cost"]) < 1e-9
    for k in ("config_checks", "edge_checks", "knn_queries", "rounds"):
        assert gpu[k] == cpu[k], k
    assert gpu["time_s"] < cpu["time_s"]
