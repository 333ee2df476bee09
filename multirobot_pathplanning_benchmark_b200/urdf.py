"""Reader for the primitive-only URDF scene exports the reference ships for its other backends
(P/assets/models/pinocchio/*.urdf, P/ = src/multi_robot_multi_goal_planning/; loaded there by
P/problems/pinocchio_env.py).  It understands what those files contain: links with box / cylinder /
sphere collision geometry and fixed / prismatic / revolute joints with axis-aligned axes, and turns
them into `Scene` frames.  URDF has no rai `contact` flag: every collision
geometry gets `contact`, visual-only links get none."""
from __future__ import annotations

import math
import xml.etree.ElementTree as ET
from typing import Dict, List, Optional

import numpy as np

from .scene import Scene, Tf

_AXIS = {(1, 0, 0): "X", (0, 1, 0): "Y", (0, 0, 1): "Z"}


def _floats(s: Optional[str], n: int) -> List[float]:
    return [float(x) for x in s.split()] if s else [0.0] * n


def _rpy(r: float, p: float, y: float) -> np.ndarray:
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def _origin(el) -> Tf:
    o = el.find("origin") if el is not None else None
    if o is None:
        return Tf()
    return Tf(_rpy(*_floats(o.get("rpy"), 3)), _floats(o.get("xyz"), 3))


def load_urdf(path: str, contact: int = 1, robot_of=lambda joint_name: None) -> Scene:
    """Scene with one frame per URDF link (named like the link) and one frame per collision geometry that has its
    own origin.  `robot_of(joint_name)` names the robot a movable joint belongs to."""
    root = ET.parse(path).getroot()
    links = {l.get("name"): l for l in root.findall("link")}
    joints = root.findall("joint")
    child_of: Dict[str, ET.Element] = {j.find("child").get("link"): j for j in joints}
    sc = Scene()

    def shape_of(link):
        c = link.find("collision")
        g = c.find("geometry") if c is not None else None
        if g is None:
            return None
        if g.find("box") is not None:
            return "box", _floats(g.find("box").get("size"), 3), _origin(c)
        if g.find("cylinder") is not None:
            e = g.find("cylinder")
            return "cylinder", [float(e.get("length")), float(e.get("radius"))], _origin(c)
        if g.find("sphere") is not None:
            return "sphere", [float(g.find("sphere").get("radius"))], _origin(c)
        raise ValueError(f"unsupported collision geometry on link {link.get('name')}")

    def add(name: str):
        if name in sc.frames:
            return
        j = child_of.get(name)
        kw = {}
        parent = None
        if j is not None:
            parent = j.find("parent").get("link")
            add(parent)
            kw["rel"] = _origin(j)
            t = j.get("type")
            if t == "fixed":
                kw["joint"] = "rigid"
            elif t in ("prismatic", "revolute", "continuous"):
                ax = tuple(int(round(v)) for v in _floats(j.find("axis").get("xyz"), 3))
                if ax not in _AXIS:
                    raise ValueError(f"joint {j.get('name')}: only axis-aligned joint axes are supported")
                kw["joint"] = ("trans" if t == "prismatic" else "hinge") + _AXIS[ax]
                lim = j.find("limit")
                kw["limits"] = [float(lim.get("lower")), float(lim.get("upper"))] if lim is not None else [-math.pi, math.pi]
                kw["robot"] = robot_of(j.get("name"))
            else:
                raise ValueError(f"unsupported joint type {t}")
        sh = shape_of(links[name])
        if sh is not None and np.allclose(sh[2].R, np.eye(3)) and np.allclose(sh[2].t, 0):
            sc.add(name, parent, shape=sh[0], size=sh[1], contact=contact, **kw)
        else:
            sc.add(name, parent, **kw)
            if sh is not None:
                sc.add(name + "_collision", name, rel=sh[2], shape=sh[0], size=sh[1], contact=contact)

    for n in links:
        add(n)
    return sc


# ----------------------------------------------------------------------------------------------
# export: a primitive Scene as URDF, for cross-checks with any installed backend
# ----------------------------------------------------------------------------------------------
def _rpy_of(R: np.ndarray):
    """inverse of _rpy (URDF fixed-axis roll-pitch-yaw)"""
    sp = -float(R[2, 0])
    if abs(sp) < 1.0 - 1e-12:
        return math.atan2(R[2, 1], R[2, 2]), math.asin(sp), math.atan2(R[1, 0], R[0, 0])
    # gimbal lock: pitch = +-90 deg, roll and yaw share one axis
    return math.atan2(-R[1, 2] * (1 if sp > 0 else -1), R[1, 1]), math.copysign(math.pi / 2, sp), 0.0


def _origin_xml(tf: Tf) -> str:
    r, p, y = _rpy_of(np.asarray(tf.R, np.float64))
    t = tf.t
    return f'<origin xyz="{t[0]:.17g} {t[1]:.17g} {t[2]:.17g}" rpy="{r:.17g} {p:.17g} {y:.17g}"/>'


def export_urdf(scene: Scene, name: str = "scene") -> str:
    """URDF text of a primitive scene: one link per frame, joints from the frames' relative poses and joint types
    (a planar `transXYPhi` base becomes prismatic x, prismatic y, revolute z through two helper links, in the scene's
    dof order), box / cylinder / sphere collision geometry.  URDF has no capsules and no rounded boxes: a capsule is
    written as a cylinder of the same length and radius and an ssBox as a box of its outer size, each with a comment
    carrying the exact rai shape, so `load_urdf` round-trips the kinematics exactly and those two shape kinds up to that
    documented approximation.  Frames without `contact` are written without collision geometry (visual only in rai)."""
    out = [f'<?xml version="1.0"?>', f'<robot name="{name}">']
    out.append('  <link name="__world__"/>')   # URDF wants ONE root link without a pose: the scene's roots hang under it
    for f in scene.frames.values():
        out.append(f'  <link name="{f.name}">')
        if f.shape is not None and f.contact != 0:
            k, s = f.shape.kind, f.shape.size
            if k == "box":
                geo = f'<box size="{s[0]:.17g} {s[1]:.17g} {s[2]:.17g}"/>'
            elif k == "ssBox":
                geo = f'<box size="{s[0]:.17g} {s[1]:.17g} {s[2]:.17g}"/> <!-- rai ssBox, rounding radius {s[3]:.17g} -->'
            elif k == "sphere":
                geo = f'<sphere radius="{s[0]:.17g}"/>'
            elif k in ("cylinder", "capsule"):
                geo = f'<cylinder length="{s[0]:.17g}" radius="{s[1]:.17g}"/>' + (" <!-- rai capsule -->" if k == "capsule" else "")
            else:
                raise ValueError(f"frame {f.name}: shape {k} has no URDF primitive")
            out.append(f'    <collision><geometry>{geo}</geometry></collision> <!-- contact {f.contact} -->')
        out.append('  </link>')
    AX = {"X": "1 0 0", "Y": "0 1 0", "Z": "0 0 1"}
    for f in scene.frames.values():
        parent = f.parent if f.parent is not None else "__world__"
        lim = f.limits
        if f.joint == "transXYPhi":
            chain = [("prismatic", "X", f.name + "__x"), ("prismatic", "Y", f.name + "__y"), ("revolute", "Z", f.name)]
            prev = parent
            for i, (jt, ax, child) in enumerate(chain):
                if child != f.name:
                    out.append(f'  <link name="{child}"/>')
                lo, hi = (lim[i] if lim is not None else (-math.pi, math.pi))
                out.append(f'  <joint name="{f.name}_j{ax.lower() if jt == "prismatic" else "phi"}" type="{jt}"><parent link="{prev}"/><child link="{child}"/>'
                           f'{_origin_xml(f.rel) if i == 0 else ""}<axis xyz="{AX[ax]}"/><limit lower="{lo:.17g}" upper="{hi:.17g}" effort="0" velocity="0"/></joint>')
                prev = child
        elif f.joint and f.joint != "rigid":
            jt = "prismatic" if f.joint.startswith("trans") else "revolute"
            lo, hi = (lim[0] if lim is not None else (-math.pi, math.pi))
            out.append(f'  <joint name="{f.name}_joint" type="{jt}"><parent link="{parent}"/><child link="{f.name}"/>{_origin_xml(f.rel)}'
                       f'<axis xyz="{AX[f.joint[-1]]}"/><limit lower="{lo:.17g}" upper="{hi:.17g}" effort="0" velocity="0"/></joint>')
        else:
            out.append(f'  <joint name="{f.name}_fixed" type="fixed"><parent link="{parent}"/><child link="{f.name}"/>{_origin_xml(f.rel)}</joint>')
    out.append('</robot>')
    return "\n".join(out) + "\n"
