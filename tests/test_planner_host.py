"""Host logic of the batch-native PRM (planner.py) with the oracle-backed stand-in device: it must find a
plan through all six modes of the 2-D handover problem, and every state / edge of that plan must be valid
according to the CPU oracle."""
import numpy as np

from multirobot_pathplanning_benchmark_b200.env import SceneModel
from multirobot_pathplanning_benchmark_b200.planner import BatchedPRM, SeqTask
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_abstract as OA
from oracle import oracle_scene as O
from tests.fakes import OracleSceneDevice


def cpu_knn(queries, corpus, slices, metric, k):
    sl = np.asarray(slices)
    out = np.full((len(queries), k), -1, np.int64)
    for i, q in enumerate(queries):
        idx = OA.knn_indices(OA.batch_config_dist(q, corpus, sl, metric), k)
        out[i, :len(idx)] = idx
    return out


def handover_tasks(start):
    return [
        SeqTask(["a1"], np.array([0.0, 0.77, 0.0]), ("a1", "obj1")),
        SeqTask(["a1", "a2"], np.array([-1.2, 1.37, 0.0, -1.2, 0.58, 0.0]), ("a2", "obj1")),
        SeqTask(["a1"], np.array([0.5, -1.13, 0.0]), ("a1", "obj2")),
        SeqTask(["a1"], np.array([1.3, 1.57, 0.0]), ("table", "obj2")),
        SeqTask(["a2"], np.array([1.22, 0.4, np.pi / 2]), ("table", "obj1")),
        SeqTask(["a1", "a2"], start.copy(), None),
    ]


def test_batched_prm_solves_handover_with_valid_plan():
    mk, kw = SCENES["2d_handover"]
    sc = mk()
    model = SceneModel(sc, kw["tol"], kw["resolution"], device=OracleSceneDevice())
    start = sc.home()
    prm = BatchedPRM(model, handover_tasks(start), start, cpu_knn, seed=1, samples_per_mode=300, transitions_per_mode=40)
    res = prm.plan(max_time=120)
    assert res.path is not None and np.isfinite(res.cost)
    modes = [m for m, _ in res.path]
    assert modes[0] == 0 and modes[-1] == 5 and all(b - a in (0, 1) for a, b in zip(modes, modes[1:]))
    assert np.allclose(res.path[0][1], start) and np.allclose(res.path[-1][1], start)
    slots = prm._mode_slots()
    for (m1, q1), (m2, q2) in zip(res.path, res.path[1:]):
        cs = model.compiled(slots[m1])
        assert O.check_configs(cs.blob64, np.float32(q1)[None].astype(np.float64))[0][0]
        if m1 == m2:
            assert O.check_edges(cs.blob64, np.float32(q1)[None].astype(np.float64), np.float32(q2)[None].astype(np.float64), kw["resolution"])[0][0]
        else:
            assert np.array_equal(q1, q2)  # mode switches happen in place
    assert res.stats["config_checks"] > 0 and res.stats["edge_checks"] > 0
