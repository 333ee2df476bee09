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
