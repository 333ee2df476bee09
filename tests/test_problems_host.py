"""Keyframe IK (keyframes.py) and the pick / place task lists (problems.py) on the host, with the oracle-backed
stand-in device: keyframes reach their targets and are collision free in their modes; the batch planner and the
reference's own PRM solve the resulting multi-mode problems with held objects."""
import random

import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import problems as P
from multirobot_pathplanning_benchmark_b200.env import SceneModel
from multirobot_pathplanning_benchmark_b200.keyframes import pick_residual, solve_ik
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from tests.fakes import OracleSceneDevice


def make_model(name):
    mk, kw = SCENES[name]
    return SceneModel(mk(), kw["tol"], kw["resolution"], device=OracleSceneDevice())


def test_ik_reaches_the_target_within_limits():
    model = make_model("box_rearrangement")
    sc = model.base
    X = sc.fk(sc.home())
    target = X["obj11"].t + np.array([0, 0, 0.1])
    q = solve_ik(sc, sc.home(), "a1_", pick_residual("a1_ur_vacuum", target, [1, 0, 0], [0, 0, -1]))
    assert q is not None
    T = sc.fk(q)["a1_ur_vacuum"]
    assert np.abs(T.t - target).max() < 2e-4 and T.R[2, 0] < -0.9999
    lim = sc.limits()
    assert np.all(q >= lim[0] - 1e-6) and np.all(q <= lim[1] + 1e-6)
    assert np.allclose(q[6:], sc.home()[6:], atol=1e-6)   # the other robot did not move (fp32-rounded)
    assert np.array_equal(q, q.astype(np.float32).astype(np.float64))  # fp32-representable, like planner samples
    # an unreachable target is reported, not approximated
    assert solve_ik(sc, sc.home(), "a1_", pick_residual("a1_ur_vacuum", [5.0, 5.0, 5.0], [1, 0, 0], [0, 0, -1]), restarts=3) is None


@pytest.mark.parametrize("name,n_moves", [("box_rearrangement", 3), ("box_stacking", 4), ("mobile_wall_four", 2)])
def test_task_list_keyframes_are_valid_in_their_modes(name, n_moves):
    model = make_model(name)
    sc = model.base
    tasks = P.manipulation_tasks(name, model, n_moves=n_moves)
    assert [t.type for t in tasks] == ["pick", "place"] * n_moves + [None]
    free = P.model_free_fn(model)
    sl = sc.robot_slices()
    relinks, cur = [], sc.copy()
    for t in tasks[:-1]:
        q = sc.home().copy()
        r = t.robots[0]
        q[sl[r][0]:sl[r][1]] = t.goal
        assert free(q, relinks), t.name                       # valid before the re-parenting ...
        parent, obj = t.frames
        if t.type == "pick":   # the tool point is centred over (or in) the object
            X = cur.fk(q)
            assert np.linalg.norm((X[parent].t - X[obj].t)[:2]) < 1e-3
            assert abs((X[parent].t - X[obj].t)[2]) < 0.15
        relinks = relinks + [(parent, obj, q.copy())]
        cur.attach(parent, obj, q)
        assert free(q, relinks), t.name                       # ... and after it
        assert cur.frames[obj].parent == parent and cur.frames[obj].contact == -1
    # every object that was moved ends on the table at its goal cell, upright
    moves, _ = P.PROBLEMS[name]
    X = cur.fk(sc.home())
    for (_, obj, goal_rel) in moves(n_moves):
        assert np.abs(X["table"].inv().apply(X[obj].t) - np.asarray(goal_rel)).max() < 1e-3
        assert X[obj].R[2, 2] > 0.9999


def test_batch_planner_solves_the_pick_place_sequence():
    from multirobot_pathplanning_benchmark_b200.planner import BatchedPRM, SeqTask
    from oracle import oracle_abstract as OA
    model = make_model("box_stacking")
    sc = model.base
    tasks = [SeqTask(list(t.robots), t.goal, t.frames) for t in P.manipulation_tasks("box_stacking", model, n_moves=1)]

    def knn(q, c, sl, metric, k):
        out = np.full((len(q), k), -1, np.int64)
        for i, row in enumerate(q):
            idx = OA.knn_indices(OA.batch_config_dist(row, c, np.asarray(sl), metric), k)
            out[i, :len(idx)] = idx
        return out
    res = BatchedPRM(model, tasks, sc.home(), knn, seed=0, samples_per_mode=300, transitions_per_mode=40).plan(max_time=120)
    assert res.path is not None and np.isfinite(res.cost)
    modes = [m for m, _ in res.path]
    assert modes == sorted(modes) and set(modes) == {0, 1, 2}     # pick mode, carry mode, return mode
    assert np.allclose(res.path[0][1], sc.home()) and np.allclose(res.path[-1][1], sc.home())


def test_reference_prm_solves_b200_box_rearrangement(reference):
    import importlib
    from multirobot_pathplanning_benchmark_b200 import env as E
    if not E.HAVE_REFERENCE:
        E = importlib.reload(E)
    from multi_robot_multi_goal_planning.planners.composite_prm_planner import CompositePRM, CompositePRMConfig
    from multi_robot_multi_goal_planning.planners.termination_conditions import RuntimeTerminationCondition
    from multi_robot_multi_goal_planning.problems.core.registry import get_all_environments
    assert {"b200.box_rearrangement", "b200.box_stacking"} <= set(get_all_environments())
    env = E.b200_box_rearrangement(device=OracleSceneDevice(), n_moves=2)
    assert [t.type for t in env.tasks] == ["pick", "place", "pick", "place", None]
    np.random.seed(1)
    random.seed(1)
    path, _ = CompositePRM(env, CompositePRMConfig()).plan(RuntimeTerminationCondition(240), optimize=False)
    assert path is not None and env.is_valid_plan(path)
    assert len({tuple(s.mode.task_ids) for s in path}) == 5 and env.is_terminal_mode(path[-1].mode)
    # the held box really rides on the tool: in the carry mode the scene graph lists it under the vacuum frame
    carry = [s.mode for s in path if s.mode.task_ids[0] == 1][0]
    assert env.get_scenegraph_info_for_mode(carry)["obj00"][0] == "a1_ur_vacuum"


def test_reference_prm_solves_b200_dep_mobile_wall(reference):
    """two-robot instance of the dependency-graph problem (rai.dep_mobile_wall_two): robots progress independently"""
    import importlib
    from multirobot_pathplanning_benchmark_b200 import env as E
    if not E.HAVE_REFERENCE:
        E = importlib.reload(E)
    from multi_robot_multi_goal_planning.planners.composite_prm_planner import CompositePRM, CompositePRMConfig
    from multi_robot_multi_goal_planning.planners.termination_conditions import RuntimeTerminationCondition
    env = E.b200_dep_mobile_wall_four(device=OracleSceneDevice(), num_robots=2)
    assert len(env.tasks) == 9 and env.collision_resolution == 0.02 and env.collision_tolerance == 0.005
    np.random.seed(1)
    random.seed(1)
    path, _ = CompositePRM(env, CompositePRMConfig()).plan(RuntimeTerminationCondition(300), optimize=False)
    assert path is not None and env.is_valid_plan(path) and env.is_terminal_mode(path[-1].mode)
    assert len({tuple(s.mode.task_ids) for s in path}) == 9
