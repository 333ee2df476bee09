"""Manipulation task lists for the named arm scenes, independent of the reference package.

The reference's `rai.box_rearrangement` / `rai.box_stacking` (P/problems/rai/rai_envs.py:1454-1594, 1916-1966; P/ =
src/multi_robot_multi_goal_planning/) are sequences of pick / place tasks whose goal keyframes rai's KOMO solves
at construction time, with shuffled goals and random restarts (rai_config.py:3074-3316, 3520-3717).  Here the same
kind of sequence is built with the numerical IK of keyframes.py: every robot repeatedly picks a box (tool at the
box, object re-parented to the tool frame, `contact = -1`) and places it at a goal location on the table (object
re-parented to the table), then all robots return home.  The result is a list of `TaskSpec`, consumed by
planner.BatchedPRM directly and by env.py to build reference `Task` objects.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from .keyframes import pick_residual, place_residual, relative_pose, solve_ik
from .scene import Scene


@dataclass
class TaskSpec:
    name: str
    robots: List[str]
    goal: np.ndarray                       # stacked joint values of `robots`
    type: Optional[str] = None             # None / "goto", "pick", "place" (reference Task.type)
    frames: Optional[Tuple[str, str]] = None   # (new parent, object) re-parented when the task completes


# (robot, object, goal position relative to the table) per move
Move = Tuple[str, str, Sequence[float]]


def pick_place_sequence(scene: Scene, moves: Sequence[Move], free: Callable[[np.ndarray, list], bool], tool: str,
                        seed: int = 0, table: str = "table") -> List[TaskSpec]:
    """Builds pick -> place tasks for `moves`, interleaved robot by robot in the given order.

    free(q, relinks) -> bool: is the full configuration q collision free in the mode reached by `relinks`
    (list of (parent, child, q_at_attach)); asked of the same device that will answer the planner.
    tool: "vacuum" (UR10 + vacuum cup: tool point `<robot>ur_vacuum`, cup axis = local x, pointing down) or
    "two_finger" (UR10 + Robotiq: tool point `<robot>ur_gripper_center`, approach axis = local z pointing down,
    jaws close along local y, aligned with a box side) or "mobile" (mobile manipulator: tool point `<robot>gripper`)."""
    rng = np.random.RandomState(seed)
    # keyframes are solved with every other robot at home, like the reference's KOMO problems, which select the
    # acting robot's joints only (rai_config.py:3086-3090): the planner moves the others out of the way
    q = scene.home().copy()
    relinks: list = []
    cur = scene.copy()
    tasks: List[TaskSpec] = []
    sl = scene.robot_slices()
    for (robot, obj, goal_rel) in moves:
        ee = robot + {"vacuum": "ur_vacuum", "two_finger": "ur_gripper_center", "mobile": "gripper"}[tool]
        X = cur.fk(q)
        box = X[obj]
        half_h = float(cur.frames[obj].shape.size[2]) / 2
        soft, extra_ok = None, (lambda qq: True)
        if tool == "mobile":
            # the gripper sphere (r = 30 mm) 2 mm above the top face; "arm pointing straight down" (gripper z = world z)
            # is a soft objective as in the reference (rai_config.py:7786-7806: distance 0, positionDiff 0,
            # scalarProductZZ 1 as sum-of-squares terms): the joint limits do not always allow it exactly
            target = box.t + np.array([0, 0, half_h + 0.032])
            res = lambda X, ee=ee, target=target: X[ee].t - target
            soft = lambda X, ee=ee: 0.05 * (X[ee].R[:, 2] - np.array([0.0, 0.0, 1.0]))
            extra_ok = lambda qq, ee=ee: cur.fk(qq)[ee].R[2, 2] > 0.8     # approach from above
        elif tool == "vacuum":
            # cup 15 mm above the top face (the tool body, a capsule-modelled cylinder, ends 11 mm beyond the cup point)
            res = pick_residual(ee, box.t + np.array([0, 0, half_h + 0.015]), [1, 0, 0], [0, 0, -1])
        else:
            # grasp centre at the box centre, approach from above, jaws across the box's local x axis
            # (10 mm above it: the palm capsule ends 20 mm above the grasp centre, the box top is 25 mm above its centre)
            res = pick_residual(ee, box.t + np.array([0, 0, 0.01]), [0, 0, 1], [0, 0, -1], align=([0, 1, 0], box.R[:, 0]))
        q_pick = solve_ik(cur, q, robot, res, accept=lambda qq: extra_ok(qq) and free(qq, relinks), rng=rng, soft=soft)
        if q_pick is None:
            raise RuntimeError(f"no collision-free pick keyframe for {robot} / {obj}")
        tasks.append(TaskSpec(f"{robot}pick_{obj}", [robot], q_pick[sl[robot][0]:sl[robot][1]].copy(), "pick", (ee, obj)))
        q = q_pick
        rel = relative_pose(cur, q, ee, obj)
        relinks = relinks + [(ee, obj, q.copy())]
        cur.attach(ee, obj, q)
        goal_world = cur.fk(q)[table].apply(np.asarray(goal_rel, np.float64))
        held = list(relinks)

        def ok_place(qq, held=held, obj=obj):
            # valid while holding and right after the object is handed back to the table
            return free(qq, held) and free(qq, held + [(table, obj, qq.copy())])
        q_place = solve_ik(cur, q, robot, place_residual(ee, rel, goal_world, yaw_axis=([1, 0, 0], [1, 0, 0]) if tool != "vacuum" else None),
                           accept=ok_place, rng=rng)
        if q_place is None:
            raise RuntimeError(f"no collision-free place keyframe for {robot} / {obj}")
        tasks.append(TaskSpec(f"{robot}place_{obj}", [robot], q_place[sl[robot][0]:sl[robot][1]].copy(), "place", (table, obj)))
        relinks = relinks + [(table, obj, q_place.copy())]
        cur.attach(table, obj, q_place)
        q = scene.home().copy()
    tasks.append(TaskSpec("terminal", list(scene.robots), scene.home().copy(), None, None))
    return tasks


def box_rearrangement_moves(n_moves: int = 4) -> List[Move]:
    """two UR10 + vacuum (rai_config.py:2947-3024): boxes of the 3 x 3 grid (pitch 0.15, z = 0.085 above the table
    frame) move to free cells of the surrounding 5 x 5 border, like the reference's intermediate goals
    (rai_config.py:3044-3066); a1 stands at x = -0.5, a2 at x = +0.5."""
    cell = lambda j, k: [j * 0.15 - 0.25, k * 0.15 - 0.35, 0.085]
    plan = [("a1_", "obj00", cell(0, 1)), ("a2_", "obj22", cell(4, 3)), ("a1_", "obj01", cell(0, 3)), ("a2_", "obj21", cell(4, 1)),
            ("a1_", "obj02", cell(1, 4)), ("a2_", "obj20", cell(3, 0))]
    return plan[:n_moves]


def box_stacking_moves(n_moves: int = 4) -> List[Move]:
    """four UR10 + Robotiq (rai_config.py:3319-3513): boxes of the 3 x 3 grid without its centre (pitch 0.15, side
    0.05, z = 0.065) are stacked on the free centre cell, one layer per move; robots a1..a4 stand at the corners."""
    top = lambda level: [0.0 * 0.15 + 0.15 - 0.05, 0.15 - 0.2, 0.065 + 0.05 * level + 0.001 * level]
    plan = [("a1_", "obj01", top(0)), ("a2_", "obj21", top(1)), ("a3_", "obj20", top(2)), ("a4_", "obj00", top(3)),
            ("a1_", "obj02", top(4)), ("a2_", "obj22", top(5))]
    return plan[:n_moves]


def model_free_fn(model) -> Callable[[np.ndarray, list], bool]:
    """`free(q, relinks)` answered by an env.SceneModel (its device checks the configuration in the mode's slot)"""
    def free(q, relinks):
        key = tuple((p, c, np.round(qq, 9).tobytes()) for p, c, qq in relinks)
        slot = model.slot_for(key, list(relinks))
        out = model.check_configs(slot, np.asarray(q, np.float32)[None])
        out = out.cpu().numpy() if hasattr(out, "cpu") else np.asarray(out)
        return bool(out[0])
    return free


def mobile_wall_moves(n_moves: int = 8, num_robots: int = 4) -> List[Move]:
    """four mobile manipulators (rai_config.py:7690-7889): robot i carries the two boxes of column i of the wall at
    y = -1 (top one first) to the goal wall at y = +1, where the column is rebuilt upside down (obj_1i -> bottom,
    obj_0i -> top), goal heights as rai_config.py:7755-7761."""
    size = np.array([0.5, 0.25, 0.15])
    w = num_robots
    plan = []
    for i in range(num_robots):
        for j in (1, 0):
            x = i * size[0] * 1.075 - w / 2 * size[0] + size[0] / 2
            plan.append((f"a{i}_", f"obj_{j}{i}", [x, 1.0, (1 - j) * size[2] * 1.01 + 0.05 + 0.1]))
    return plan[:n_moves]


PROBLEMS = {
    # scene name -> (moves, tool)
    "box_rearrangement": (box_rearrangement_moves, "vacuum"),
    "box_stacking": (box_stacking_moves, "two_finger"),
    "mobile_wall_four": (mobile_wall_moves, "mobile"),
}


def manipulation_tasks(scene_name: str, model, n_moves: int = 4, seed: int = 0) -> List[TaskSpec]:
    moves, tool = PROBLEMS[scene_name]
    moves_list = moves(n_moves) if scene_name != "mobile_wall_four" else moves(n_moves, len(model.base.robots))
    return pick_place_sequence(model.base, moves_list, model_free_fn(model), tool, seed=seed)
