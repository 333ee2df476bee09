"""Host logic of env.py against the unmodified reference planners (CPU, build container only):
the B200 environments are driven by the reference's own PRM / RRT* / EIT* through the BaseProblem
interface, with the device replaced by an oracle-backed stand-in (tests/fakes.py)."""
import random

import numpy as np
import pytest

from tests.fakes import OracleAbstractDevice, OracleSceneDevice


@pytest.fixture(scope="module")
def envmod(reference):
    import importlib
    from multirobot_pathplanning_benchmark_b200 import env
    if not env.HAVE_REFERENCE:  # imported by an earlier test before the reference was on sys.path
        env = importlib.reload(env)
    assert env.HAVE_REFERENCE
    return env


def test_b200_envs_are_registered(reference, envmod):
    from multi_robot_multi_goal_planning.problems.core.registry import get_all_environments
    names = [n for n in get_all_environments() if n.startswith("b200.")]
    assert {"b200.2d_handover", "b200.abstract_test", "b200.box_rearrangement_goto"} <= set(names)


def walk_modes(env):
    """modes along the canonical sequence, entering each at its task's goal keyframe"""
    m = env.start_mode
    q = env.start_pos
    out = [(m, q)]
    while not env.is_terminal_mode(m):
        task = env.get_active_task(m, None)
        goal = task.goal.sample(m)
        q_new = env.start_pos.from_flat(q.state().copy())
        off = 0
        for r in task.robots:
            i = env.robots.index(r)
            q_new[i] = goal[off:off + env.robot_dims[r]]
            off += env.robot_dims[r]
        assert env.is_transition(q_new, m)
        m = env.get_next_modes(q_new, m)[0]
        q = q_new
        out.append((m, q))
    return out


def test_handover_keyframes_and_scenegraph(envmod):
    env = envmod.b200_two_dim_handover(device=OracleSceneDevice())
    assert env.is_collision_free(env.start_pos, env.start_mode)
    modes = walk_modes(env)
    assert len(modes) == 6
    parents = []
    for m, q in modes:
        assert env.is_collision_free(q, m), f"entry configuration of mode {m} collides"
        parents.append(m.sg["obj1"][0] + "/" + m.sg["obj2"][0])
    # obj1: table -> a1 -> a2 (handover) ... -> table; obj2: table -> a1 -> table
    assert parents == ["table/table", "a1/table", "a2/table", "a2/a1", "a2/table", "table/table"]
    # the held object moves with its holder and collides like a part of it
    m_hold, q_hold = modes[1]
    q_bad = env.start_pos.from_flat(q_hold.state().copy())
    q_bad[0] = np.array([0.0, 0.55, 0.0])  # a1 pushes obj1 (hanging 0.37 south of it) into obs1
    assert not env.is_collision_free(q_bad, m_hold)
    assert env.is_collision_free(q_bad, env.start_mode) is False  # in the start mode a1 overlaps obj1 itself
    # distinct kinematic trees get distinct device slots, equal ones share
    assert len(env.model._slots) == len({p for p in parents}) or len(env.model._slots) >= 4
    assert hash(modes[0][0]) != hash(modes[1][0])


def test_edge_and_robot_rule_through_env(envmod):
    dev = OracleSceneDevice()
    env = envmod.b200_two_dim_handover(device=dev)
    m = env.start_mode
    q1 = env.start_pos
    q2 = env.start_pos.from_flat(np.array([-0.5, 0.8, 1.0, 0.0, -0.5, 0.3]))
    assert env.is_edge_collision_free(q1, q2, m)
    q3 = env.start_pos.from_flat(np.array([0.5, 0.8, 0.0, 0.0, -0.5, 0.0]))  # a1 crosses obs2
    assert not env.is_edge_collision_free(q1, q3, m)
    assert env.is_edge_collision_free(q1, q3, m, N_start=0, N_max=1) in (True, False)
    # a1 inside obs2, a2 free: collides "for a1", not "for a2"
    q = np.array([0.4, 1.0, 0.0, 0.0, -0.5, 0.0])
    assert not env.is_collision_free_for_robot("a1", q, m)
    assert env.is_collision_free_for_robot("a2", q, m)
    with pytest.raises(ValueError):
        env.is_collision_free(None, m)


def run_planner(env, planner_name, seed, max_time=20):
    from multi_robot_multi_goal_planning.planners.composite_prm_planner import CompositePRM, CompositePRMConfig
    from multi_robot_multi_goal_planning.planners.planner_rrtstar import RRTstar
    from multi_robot_multi_goal_planning.planners.rrtstar_base import BaseRRTConfig
    from multi_robot_multi_goal_planning.planners.termination_conditions import RuntimeTerminationCondition
    np.random.seed(seed)
    random.seed(seed)
    if planner_name == "prm":
        planner = CompositePRM(env, CompositePRMConfig())
    else:
        planner = RRTstar(env, BaseRRTConfig(with_mode_validation=False))
    return planner.plan(RuntimeTerminationCondition(max_time), optimize=False)


def test_reference_prm_solves_b200_handover(envmod):
    env = envmod.b200_two_dim_handover(device=OracleSceneDevice())
    path, info = run_planner(env, "prm", 1, max_time=60)
    assert path is not None and env.is_valid_plan(path)
    assert env.is_terminal_mode(path[-1].mode)


@pytest.mark.parametrize("planner", ["prm", "rrt"])
def test_planners_behave_identically_on_b200_abstract_test(reference, envmod, planner):
    """Same seed, reference env vs b200 env with bit-exact collision answers -> identical plans."""
    ref_env = reference.get_env_by_name("abstract.test")
    p_ref, _ = run_planner(ref_env, planner, 3)
    b_env = envmod.b200_abstract_test(device=OracleAbstractDevice())
    p_b, _ = run_planner(b_env, planner, 3)
    assert p_ref is not None and p_b is not None and len(p_ref) == len(p_b)
    for a, b in zip(p_ref, p_b):
        assert np.array_equal(a.q.state(), b.q.state()) and a.mode.task_ids == b.mode.task_ids


def test_batched_path_check_equals_reference_path_check(reference, envmod):
    from multi_robot_multi_goal_planning.problems.planning_env import BaseProblem, State
    env = envmod.b200_two_dim_handover(device=OracleSceneDevice())
    rng = np.random.RandomState(0)
    modes = walk_modes(env)
    agree = 0
    for trial in range(30):
        m, q0 = modes[trial % len(modes)]
        pts = [q0.state()] + [q0.state() + rng.uniform(-0.25, 0.25, 6) for _ in range(3)]
        path = [State(env.start_pos.from_flat(p), m) for p in pts]
        want = BaseProblem.is_path_collision_free(env, path)   # the reference's own loop over single queries
        assert env.is_path_collision_free(path) == want
        agree += want
    assert 0 < agree < 30
    qs = env.sample_valid_uniform_batch(env.start_mode, 100, np.random.RandomState(1))
    assert qs.shape == (100, 6) and all(env.is_collision_free(env.start_pos.from_flat(q), env.start_mode) for q in qs[:10])
