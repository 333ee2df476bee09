"""Goal keyframes for pick / place tasks by numerical inverse kinematics on the host scene model.

The reference computes the keyframes of its manipulation problems with rai's KOMO optimiser at construction
time (P/problems/rai/rai_config.py:28-62 `solve_komo_problem`; box_rearrangement :3074-3200, box_stacking
:3520-3700, mobile wall :7770-7889; P/ = src/multi_robot_multi_goal_planning/), with random restarts, so no two
instances of a reference environment share keyframes either.  KOMO lives in the un-vendored `robotic` wheel;
this module is the stand-in: damped least squares (scipy) on `Scene.fk`, restarted until the result is
collision free according to the device that will also answer the planner's queries.

A keyframe only has to satisfy what the task needs -- tool point at the object, tool axis along a world
direction (pick), or the held object at its goal pose (place) -- which is what the residuals below say.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np
from scipy.optimize import least_squares

from .scene import Scene, Tf


def _robot_slice(scene: Scene, robot: str) -> slice:
    s, e = scene.robot_slices()[robot]
    return slice(s, e)


def solve_ik(scene: Scene, q_full: np.ndarray, robot: str, residual: Callable[[dict], np.ndarray],
             accept: Optional[Callable[[np.ndarray], bool]] = None, rng: Optional[np.random.RandomState] = None,
             restarts: int = 60, tol: float = 1e-4, regularise: float = 1e-3,
             soft: Optional[Callable[[dict], np.ndarray]] = None) -> Optional[np.ndarray]:
    """Joint values of `robot` (others stay as in q_full) with |residual| < tol that `accept` (full q) agrees to.
    `residual(X)` gets the frame poses {name: Tf} of the whole scene; `soft(X)` is minimised along with it but not
    required to vanish (KOMO's sum-of-squares objectives).  First start = current values, then uniform restarts in
    the joint limits.  Returns the full configuration or None."""
    rng = rng or np.random.RandomState(0)
    sl = _robot_slice(scene, robot)
    lim = scene.limits()[:, sl]
    q = np.array(q_full, np.float64)
    seed = q[sl].copy()

    anchor = seed.copy()  # the regulariser picks the solution nearest to the start point of the attempt

    def fun(x):
        q[sl] = x
        X = scene.fk(q)
        r = [residual(X), regularise * (x - anchor)]
        if soft is not None:
            r.append(soft(X))
        return np.concatenate(r)

    n_hard = len(residual(scene.fk(q)))
    for t in range(restarts):
        x0 = np.clip(seed, lim[0], lim[1]) if t == 0 else rng.uniform(lim[0], lim[1])
        anchor[:] = x0
        try:
            sol = least_squares(fun, x0, bounds=(lim[0], lim[1]), xtol=1e-10, ftol=1e-10, gtol=1e-10, max_nfev=400)
            # the regulariser (and soft terms) choose among the solutions; a second pass on the task residual
            # alone, started there, removes the bias they leave on it
            sol = least_squares(lambda x: fun(x)[:n_hard], sol.x, bounds=(lim[0], lim[1]), xtol=1e-12, ftol=1e-12, gtol=1e-12,
                                max_nfev=100)
        except ValueError:
            continue
        q[sl] = sol.x
        if np.max(np.abs(residual(scene.fk(q)))) > tol:
            continue
        # planners hand fp32 configurations to the device: the keyframe is the fp32-rounded vector
        out = q.astype(np.float32).astype(np.float64)
        if accept is None or accept(out):
            return out
    return None


def pick_residual(ee: str, target: np.ndarray, tool_axis_local: Sequence[float], tool_axis_world: Sequence[float],
                  align: Optional[Tuple[Sequence[float], Sequence[float]]] = None) -> Callable[[dict], np.ndarray]:
    """tool point `ee` at `target`, the tool's local axis along a world direction (e.g. the vacuum cup pointing
    down); `align` = (local axis, world axis) additionally asks a second tool axis to be parallel (either sign) to a
    world axis, for parallel-jaw grippers closing across a box."""
    target = np.asarray(target, np.float64)
    al, aw = np.asarray(tool_axis_local, np.float64), np.asarray(tool_axis_world, np.float64)

    def f(X):
        T = X[ee]
        r = [T.t - target, T.R @ al - aw]
        if align is not None:
            a = T.R @ np.asarray(align[0], np.float64)
            w = np.asarray(align[1], np.float64)
            r.append(np.cross(a, w))  # zero iff parallel or anti-parallel
        return np.concatenate(r)
    return f


def place_residual(holder: str, rel: Tf, goal_pos: np.ndarray, up_local: Sequence[float] = (0, 0, 1),
                   yaw_axis: Optional[Tuple[Sequence[float], Sequence[float]]] = None) -> Callable[[dict], np.ndarray]:
    """the object hanging on frame `holder` with relative pose `rel` sits at goal_pos, its local up axis along world z;
    `yaw_axis` = (object-local axis, world axis) fixes the rotation about z up to sign."""
    goal_pos = np.asarray(goal_pos, np.float64)
    ul = np.asarray(up_local, np.float64)

    def f(X):
        T = X[holder] @ rel
        r = [T.t - goal_pos, T.R @ ul - np.array([0.0, 0.0, 1.0])]
        if yaw_axis is not None:
            r.append(np.cross(T.R @ np.asarray(yaw_axis[0], np.float64), np.asarray(yaw_axis[1], np.float64)))
        return np.concatenate(r)
    return f


def relative_pose(scene: Scene, q: np.ndarray, parent: str, child: str) -> Tf:
    """pose of `child` in `parent` at configuration q: what rai's C.attach(parent, child) freezes (rai_base_env.py:801)"""
    X = scene.fk(q)
    return X[parent].inv() @ X[child]
