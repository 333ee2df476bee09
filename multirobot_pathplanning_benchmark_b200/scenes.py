"""The named scenes of BASELINE.json, rebuilt from the reference's scene descriptions.

Geometry of record (paths relative to /root/reference,
P/ = src/multi_robot_multi_goal_planning/):
  * UR10 chain + capsules: P/assets/models/rai/ur10/ur10_clean.g:1-29, ur10.g:9-39,
    ur10_vacuum.g:5-24, ur10_two_finger.g:5-14, robotiq/robotiq.g:12-32,
    robotiq/robotiq_clean.g:1-24.
  * mobile manipulator: P/assets/models/rai/mobile-manipulator-restricted.g:1-68.
  * 2d_handover: P/problems/rai/rai_config.py:65-100, 751-839.
  * box_rearrangement: rai_config.py:2947-3064.   * box_stacking: rai_config.py:3319-3513.
  * mobile wall (dep_mobile_wall_four): rai_config.py:7690-7768.
Only frames that carry a joint or a collision primitive (or are needed as attach
targets) are kept: mesh / marker / contact-less frames never collide
(P/problems/rai_base_env.py:234-255).
"""
from __future__ import annotations

import numpy as np

from .scene import Scene, Tf

S2 = 0.707107


UR10_FILES = {"vacuum": "ur10/ur10_vacuum.g", "two_finger": "ur10/ur10_two_finger.g", None: "ur10/ur10.g"}
MOBILE_FILE = "mobile-manipulator-restricted.g"


def add_ur10(sc: Scene, prefix: str, parent: str, base_rel: Tf, tool: str, q0=None, models_dir=None) -> None:
    """UR10 with collision capsules.  `tool` in {"vacuum", "two_finger", None}.  With `models_dir` (the
    reference's P/assets/models/rai directory) the robot is read from the reference's own `.g` file
    (gfile.add_g_model) instead of the transcription below; tests/test_gfile.py checks both give the same scene."""
    if models_dir is not None:
        import os
        from .gfile import add_g_model
        add_g_model(sc, os.path.join(models_dir, UR10_FILES[tool]), prefix, parent, base_rel, robot=prefix)
        return
    p = prefix + "ur_"
    rob = prefix
    lim = {  # ur10.g:34-39
        "shoulder_pan_joint": [-3.28319, 3.28319], "shoulder_lift_joint": [-3.28319, 0.28319],
        "elbow_joint": [-2.68319, 2.68319], "wrist_1_joint": [-3.28319, 3.28319],
        "wrist_2_joint": [-3.28319, 3.28319], "wrist_3_joint": [-3.28319, 3.28319]}
    if q0 is None:  # ur10_vacuum.g:5-10 / ur10_two_finger.g:5-10
        q0 = [0.0, -2.0, 1.0, -1.0, -1.571, 1.0 if tool == "vacuum" else 0.0]
    # addFile(...).setParent(table).setRelativePose(...).setJoint(rigid): the file's root
    # frame becomes a rigid-jointed child of the table (rai_config.py:2967-2971)
    sc.add(p + "base", parent, rel=base_rel, joint="rigid")
    sc.add(p + "base_link", p + "base")
    sc.add(p + "coll0", p + "base_link", rel="t(0 0 .1) d(90 0 0 1)", shape="capsule", size=[.12, .09], contact=-1)
    chain = [  # (origin rel, joint name, joint type)  ur10_clean.g:3-27
        ([0, 0, 0.1273, 1, 0, 0, 0], "shoulder_pan_joint", "hingeZ"),
        ([0, 0.220941, 0, S2, 0, S2, 0], "shoulder_lift_joint", "hingeY"),
        ([0, -0.1719, 0.612, 1, 0, 0, 0], "elbow_joint", "hingeY"),
        ([0, 0, 0.5723, S2, 0, S2, 0], "wrist_1_joint", "hingeY"),
        ([0, 0.1149, 0, 1, 0, 0, 0], "wrist_2_joint", "hingeZ"),
        ([0, 0, 0.1157, 1, 0, 0, 0], "wrist_3_joint", "hingeY"),
    ]
    prev = p + "base_link"
    for i, (rel, jn, jt) in enumerate(chain):
        sc.add(p + jn + "_origin", prev, rel=rel)
        sc.add(p + jn, p + jn + "_origin", joint=jt, limits=lim[jn], q0=[q0[i]], robot=rob)
        prev = p + jn
    # collision capsules ur10.g:9-23 (coll0 above)
    sc.add(p + "coll2", p + "shoulder_lift_joint_origin", rel="t(0 -.12 .01) d(90 1 0 0)", shape="capsule", size=[.17, .09], contact=-1)
    sc.add(p + "coll3", p + "shoulder_lift_joint", rel="t(0 -.04 .3) d(90 0 0 1)", shape="capsule", size=[.5, .065], contact=-1)
    sc.add(p + "coll4", p + "elbow_joint_origin", rel="t(0 .06 .0) d(90 1 0 0)", shape="capsule", size=[.16, .065], contact=-1)
    sc.add(p + "coll5", p + "elbow_joint", rel="t(0 0 .3) d(90 0 0 1)", shape="capsule", size=[.5, .06], contact=-2)
    sc.add(p + "coll6", p + "wrist_1_joint", rel="t(0 .02 0) d(90 1 0 0)", shape="capsule", size=[.1, .05], contact=-2)
    sc.add(p + "coll7", p + "wrist_2_joint", rel="t(0 0 .04) d(90 0 0 1)", shape="capsule", size=[.12, .05], contact=-2)
    sc.add(p + "coll8", p + "wrist_3_joint", rel="t(0 .02 0) d(90 1 0 0)", shape="capsule", size=[.09, .05], contact=-2)
    # ee: ur10_clean.g:27-29
    sc.add(p + "ee_fixed_joint_origin", p + "wrist_3_joint", rel=[0, 0.0922, 0, S2, 0, 0, S2])
    sc.add(p + "ee_fixed_joint", p + "ee_fixed_joint_origin", joint="rigid")
    sc.add(p + "ee_link", p + "ee_fixed_joint")
    if tool == "vacuum":  # ur10_vacuum.g:12-24 (frames carry the file prefix only)
        sc.add(prefix + "gripper_fill", p + "ee_link", rel="d(90 0 1 0) t(0 0 .025)", shape="cylinder", size=[.05, .021], contact=-1)
        sc.add(prefix + "ur_vacuum", p + "ee_link", rel="t(.06 0 0)")  # contact 0: attach target only
    elif tool == "two_finger":  # ur10_two_finger.g:12-14, robotiq.g:12-32, robotiq_clean.g:3-20
        b = p + "robotiq_base"
        sc.add(b, p + "ee_link", rel="d(90 0 1 0) t(0 0 .036)")
        sc.add(p + "gripper", b, rel="t(0 0 .13)")
        sc.add(p + "gripper_center", p + "gripper")
        sc.add(p + "palm", b, rel="d(90 1 0 0) t(0 .07 .0)", shape="capsule", size=[.11, .04], contact=-1)
        # finger frames are joint-less in robotiq_clean.g (finger_joint inactive, robotiq.g:31-32)
        right = Tf.from_pose([0, 0.0306011, 0.054904, 1, 0, 0, 0]) @ Tf.from_pose([0, 0.0376, 0.043, 1, 0, 0, 0])
        left = Tf.from_pose([0, -0.0306011, 0.054904, 6.12323e-17, 0, 0, 1]) @ Tf.from_pose([0, 0.0376, 0.043, 1, 0, 0, 0])
        sc.add(p + "finger1", b, rel=right @ Tf.from_pose([0, -.009, .025]), shape="capsule", size=[.04, .02], contact=-2)
        sc.add(p + "finger2", b, rel=left @ Tf.from_pose([0, -.009, .025]), shape="capsule", size=[.04, .02], contact=-2)


def add_mobile_manipulator(sc: Scene, prefix: str, z: float, q0_base, models_dir=None) -> None:
    """mobile-manipulator-restricted.g:1-68; `world` frame moved to height z
    (rai_config.py:7708).  `models_dir`: read the reference's `.g` file instead (see add_ur10)."""
    p, rob = prefix, prefix
    if models_dir is not None:
        import os
        from .gfile import add_g_model
        add_g_model(sc, os.path.join(models_dir, MOBILE_FILE), prefix, None, Tf(None, [0, 0, z]), robot=rob,
                    q0={"base": q0_base})
        return
    sc.add(p + "world", None, rel=[0, 0, z])
    sc.add(p + "base", p + "world", joint="transXYPhi", limits=[[-2, 4], [-2, 2], [-3.14, 3.14]], q0=q0_base, robot=rob)
    sc.add(p + "base_coll", p + "base", shape="ssBox", size=[.4, .4, .4, .05], contact=1)
    sc.add(p + "arm0", p + "base", rel="t(0 0 .3)")
    sc.add(p + "arm0_coll", p + "arm0", shape="capsule", size=[.2, .1], contact=1)
    sc.add(p + "joint1", p + "arm0", rel="t(0 0 .2)", joint="hingeX", limits=[0, 1.7], q0=[1.0], robot=rob)
    sc.add(p + "arm1", p + "joint1", rel="t(0 0 .4)")
    sc.add(p + "arm1_coll", p + "arm1", shape="capsule", size=[.5, .08], contact=-2)
    sc.add(p + "joint2", p + "arm1", rel="t(0 0 .4)", joint="hingeX", limits=[0, 1.7], q0=[1.0], robot=rob)
    sc.add(p + "arm2", p + "joint2", rel="t(0 0 .2)")
    sc.add(p + "arm2_coll", p + "arm2", shape="capsule", size=[.3, .06], contact=1)
    sc.add(p + "joint2a", p + "arm2", rel="t(0 0 .25)", joint="hingeZ", limits=[-3.14, 3.14], q0=[0.0], robot=rob)
    # `Edit gripper(joint2a) { Q:"t(0 0 .01) d(180 1 0 0)" }` overrides the declared pose (.g:68)
    sc.add(p + "gripper", p + "joint2a", rel="t(0 0 .01) d(180 1 0 0)", shape="sphere", size=[.03], contact=1)


def add_table_with_walls(sc: Scene, width: float, length: float) -> None:
    """rai_config.py:65-100.  ST.box ignores the 4th size entry."""
    sc.add("table", None, rel=[0, 0, 1.0], shape="box", size=[width, length, .06], contact=1)
    sc.add("wall1", "table", rel=[0, width / 2 + .1, .07], joint="rigid", shape="box", size=[width - .001, .2, .06], contact=1)
    sc.add("wall2", "table", rel=[0, -width / 2 - .1, .07], joint="rigid", shape="box", size=[width - .001, .2, .06], contact=1)
    sc.add("wall3", "table", rel=[length / 2 + .1, 0, .07], joint="rigid", shape="box", size=[.2, length + .4 - .001, .06], contact=1)
    sc.add("wall4", "table", rel=[-length / 2 - .1, 0, .07], joint="rigid", shape="box", size=[.2, length + .4 - .001, .06], contact=1)


def make_two_dim_handover() -> Scene:
    """rai.2d_handover scene (rai_config.py:751-839)."""
    sc = Scene()
    add_table_with_walls(sc, 4, 4)
    lim = [[-2, 2], [-2, 2], [-3.14, 3.14]]
    sc.add("pre_agent_1_frame", "table", rel=[0, 0, .07], joint="rigid")
    sc.add("a1", "pre_agent_1_frame", joint="transXYPhi", limits=lim, q0=[-.5, .8, 0], shape="cylinder", size=[.04, .15], contact=1, robot="a1")
    sc.add("pre_agent_2_frame", "table", rel=[0, 0, .07], joint="rigid")
    sc.add("a2", "pre_agent_2_frame", joint="transXYPhi", limits=lim, q0=[0, -.5, 0], shape="cylinder", size=[.04, .2], contact=1, robot="a2")
    sc.add("obj1", "table", rel=[0, .4, .07], joint="rigid", shape="box", size=[.4, .4, .06], contact=1)
    sc.add("obj2", "table", rel=[.5, -1.5, .07], joint="rigid", shape="box", size=[.3, .4, .06], contact=1)
    for n, pos, size in (("obs1", [0, 0, .07], [2.3, .2, .06]), ("obs2", [.4, 1.05, .07], [.2, 1.8, .06]),
                         ("obs3", [-.4, -.6, .07], [.2, .9, .06]), ("obs4", [.8, .8, .07], [.6, .2, .06])):
        sc.add(n, "table", rel=pos, joint="rigid", shape="box", size=size, contact=-2)
    return sc


_UR_QUAT_NEG = [0.7071, 0, 0, -0.7071]
_UR_QUAT_POS = [0.7071, 0, 0, 0.7071]


def make_box_rearrangement(num_robots: int = 2, num_boxes: int = 9, models_dir=None) -> Scene:
    """rai.box_rearrangement scene (rai_config.py:2947-3024): UR10 + vacuum tools."""
    sc = Scene()
    sc.add("table", None, rel=[0, 0, .2], shape="box", size=[2, 3, .06], contact=1)
    bases = [([-.5, .5, 0], _UR_QUAT_NEG), ([.5, .5, 0], _UR_QUAT_NEG),
             ([.5, -.6, 0], _UR_QUAT_POS), ([-.5, -.6, 0], _UR_QUAT_POS)]
    for i in range(num_robots):
        pos, quat = bases[i]
        add_ur10(sc, f"a{i + 1}_", "table", Tf.from_pose(pos + quat), "vacuum", models_dir=models_dir)
    w, d, size = 3, 3, 0.1
    cnt = 0
    for k in range(d):
        for j in range(w):
            if cnt == num_boxes:
                break
            pos = [j * size * 1.5 - w / 2 * size + size / 2, k * size * 1.5 - 0.2, 0.085]
            sc.add(f"obj{j}{k}", "table", rel=pos, joint="rigid", shape="ssBox", size=[size, size, size, .005], contact=1)
            cnt += 1
    return sc


def make_box_stacking(num_robots: int = 4, num_boxes: int = 8, models_dir=None) -> Scene:
    """rai.box_stacking scene (rai_config.py:3319-3494): UR10 + Robotiq two-finger tools."""
    sc = Scene()
    sc.add("table", None, rel=[0, 0, .2 - .03], shape="box", size=[3, 3, .06], contact=1)
    bases = [([-.5, .5, .03], _UR_QUAT_NEG), ([.5, .5, .03], _UR_QUAT_NEG),
             ([.5, -.6, .03], _UR_QUAT_POS), ([-.5, -.6, .03], _UR_QUAT_POS)]
    for i in range(num_robots):
        pos, quat = bases[i]
        add_ur10(sc, f"a{i + 1}_", "table", Tf.from_pose(pos + quat), "two_finger", models_dir=models_dir)
    w, d, size, height = 3, 3, 0.05, 0.065
    cnt = 0
    for k in range(d):
        for j in range(w):
            if (k == 1 and j == 1) or cnt == num_boxes:
                continue
            pos = [j * size * 3 - w / 2 * size + size / 2, k * size * 3 - 0.2, height]
            sc.add(f"obj{j}{k}", "table", rel=pos, joint="rigid", shape="ssBox", size=[size, size, size, .005], contact=1)
            cnt += 1
    return sc


def make_mobile_wall(num_robots: int = 4, models_dir=None) -> Scene:
    """rai.dep_mobile_wall_four scene (rai_config.py:7690-7768)."""
    sc = Scene()
    sc.add("table", None, rel=[0, 0, -.02], shape="box", size=[20, 20, .06], contact=1)
    for i in range(num_robots):
        add_mobile_manipulator(sc, f"a{i}_", 0.25, [2.5, -(num_robots - 1) / 2 + i, -np.pi / 2], models_dir=models_dir)
    w, h = num_robots, 2
    size = np.array([.5, .25, .15])
    for i in range(h):
        for j in range(w):
            pos = [j * size[0] * 1.075 - w / 2 * size[0] + size[0] / 2, -1, i * size[2] * 1.05 + .05 + .1]
            sc.add(f"obj_{i}{j}", "table", rel=pos, joint="rigid", shape="box", size=list(size), contact=1)
    return sc


def make_abstract_like_scene() -> Scene:
    """A 3-D lift of abstract.test (two sphere agents, one sphere and one box obstacle) that
    exercises the point-point / point-box code paths of the generic kernels in tests."""
    sc = Scene()
    sc.add("ground", None)
    lim = [[-2, 2], [-2, 2], [-3.14, 3.14]]
    sc.add("a1", "ground", joint="transXYPhi", limits=lim, q0=[-.8, 0, 0], shape="sphere", size=[.1], contact=1, robot="a1")
    sc.add("a2", "ground", joint="transXYPhi", limits=lim, q0=[.8, 0, 0], shape="sphere", size=[.1], contact=1, robot="a2")
    sc.add("obs_sphere", "ground", joint="rigid", shape="sphere", size=[.2], contact=1)
    # hangs on the sphere with contact -1 so the two (overlapping) obstacles do not collide with each other
    sc.add("obs_rect", "obs_sphere", rel=[0, .4, 0], joint="rigid", shape="box", size=[.5, .5, .5], contact=-1)
    return sc


SCENES = {
    "2d_handover": (make_two_dim_handover, dict(tol=0.01, resolution=0.01)),          # rai_envs.py:545-546
    "box_rearrangement": (make_box_rearrangement, dict(tol=0.01, resolution=0.01)),  # rai_envs.py:1592, rai_base_env.py:300
    "box_stacking": (make_box_stacking, dict(tol=0.0, resolution=0.01)),             # rai_envs.py:1962-1964
    "mobile_wall_four": (make_mobile_wall, dict(tol=0.005, resolution=0.02)),        # rai_envs.py:2272-2273
    "abstract_like": (make_abstract_like_scene, dict(tol=0.0, resolution=0.01)),
}
