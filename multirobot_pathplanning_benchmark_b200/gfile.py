"""Reader for the reference's on-disk robot models (rai `.g` graph files) -> `Scene` frames.

The reference builds every rai scene from `.g` files under P/assets/models/rai/ (P/ =
src/multi_robot_multi_goal_planning/), loaded with `C.addFile(path, namePrefix=...)`
(P/problems/rai/rai_config.py:2962-2971, 3367-3385, 7703-7712).  This module parses the subset of
the format those models use, so that a scene of the B200 backend can be assembled from the very
files the reference reads instead of from a hand transcription:

    name (parent) : { key: value, key value ... }      frame definition (colon / commas optional)
    Edit name (parent) { ... }                          merge attributes, optionally re-parent
    Include: <relative/path.g>                          textual include, relative to the including file
    Prefix: "ur_"  /  Prefix: false                     prefix for the names defined from here on
    # comment

Frame attributes understood: `rel` / `Q` / `X` / `A` (pose arrays [x y z (qw qx qy qz)] or the
transformation mini-language "t(x y z) d(deg ax ay az) ...", see scene.Tf.parse), `joint`,
`limits`, `q`, `joint_active`, `shape`, `size`, `contact`.  Everything else (colours, meshes,
masses, `logical`, ...) is kept in `GFrame.attrs` but not interpreted.

Pose semantics (rai): a frame's transform from its parent is A * J(q) * Q for a frame that carries
a joint (`A` = fixed pre-transform; `Q` is then the joint's initial transform and is ignored here,
`q` gives the home value) and Q (alias `rel`) for a plain frame; `X` is an absolute pose and is
only honoured on root frames.
"""
from __future__ import annotations

import os
import re
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

from .scene import JOINT_DOF, Scene, Tf

PRIMITIVES = ("sphere", "capsule", "cylinder", "box", "ssBox")

_TOKEN = re.compile(r"""
    \s+ | \#[^\n]*                         # whitespace, comments
  | (?P<str>"[^"]*"|'[^']*')
  | (?P<path><[^>\n]*>)
  | (?P<punct>[{}()\[\]:,])
  | (?P<word>[^\s{}()\[\]:,"'<>\#]+)
""", re.X)


def _tokens(text: str) -> List[str]:
    out, pos = [], 0
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            raise ValueError(f".g syntax error near {text[pos:pos + 40]!r}")
        pos = m.end()
        if m.lastgroup:
            out.append(m.group(m.lastgroup))
    return out


def _scalar(tok: str):
    if tok[0] in "\"'":
        return tok[1:-1]
    if tok[0] == "<":
        return tok
    low = tok.lower()
    if low in ("true", "false"):
        return low == "true"
    try:
        return float(tok) if re.search(r"[.eE]", tok) or not tok.lstrip("+-").isdigit() else int(tok)
    except ValueError:
        return tok


@dataclass
class GFrame:
    name: str
    parent: Optional[str]
    attrs: Dict[str, object] = field(default_factory=dict)


class _Parser:
    def __init__(self):
        self.frames: List[GFrame] = []
        self.prefix = ""

    # ---- lookups ----------------------------------------------------------------------------
    def find(self, name: str, parent: Optional[str] = None) -> Optional[GFrame]:
        hits = [f for f in self.frames if f.name == name]
        if parent is not None:
            same = [f for f in hits if f.parent == parent]
            if same:
                return same[0]
        return hits[0] if hits else None

    # ---- token-level parsing ------------------------------------------------------------------
    def value(self, t: List[str], i: int):
        tok = t[i]
        if tok == "[":
            vals = []
            i += 1
            while t[i] != "]":
                if t[i] != ",":
                    vals.append(_scalar(t[i]))
                i += 1
            return vals, i + 1
        if tok == "{":
            d, i = self.attr_block(t, i)
            return d, i
        return _scalar(tok), i + 1

    def attr_block(self, t: List[str], i: int):
        assert t[i] == "{", t[i]
        i += 1
        d: Dict[str, object] = {}
        while t[i] != "}":
            if t[i] == ",":
                i += 1
                continue
            key = t[i]
            i += 1
            if t[i] == ":":
                d[key], i = self.value(t, i + 1)
            elif t[i] not in ("}", ",") and (t[i][0] in "\"'<[{" or re.match(r"[+-]?[\d.]", t[i])):
                d[key], i = self.value(t, i)  # `key value` without the colon
            else:
                d[key] = True
        return d, i + 1

    def parse_file(self, path: str) -> None:
        with open(path) as f:
            t = _tokens(f.read())
        here = os.path.dirname(os.path.abspath(path))
        i = 0
        while i < len(t):
            tok = t[i]
            if tok == "Include":
                j = i + 2 if t[i + 1] == ":" else i + 1
                inc = t[j].strip("<>'\"")
                self.parse_file(os.path.join(here, inc))
                i = j + 1
                continue
            if tok == "Prefix":
                j = i + 2 if t[i + 1] == ":" else i + 1
                v = _scalar(t[j])
                self.prefix = v if isinstance(v, str) and v not in ("false", "False") else ""
                i = j + 1
                continue
            edit = tok == "Edit"
            if edit:
                i += 1
            name = t[i]
            i += 1
            parents: List[str] = []
            if i < len(t) and t[i] == "(":
                i += 1
                while t[i] != ")":
                    if t[i] != ",":
                        parents.append(t[i])
                    i += 1
                i += 1
            if i < len(t) and t[i] == ":":
                i += 1
            attrs, i = self.attr_block(t, i)
            parent = parents[0] if parents else None
            if edit:
                # names in Edit statements are looked up as written, then with the active prefix
                fr = self.find(name, parent) or self.find(self.prefix + name, None if parent is None else self.prefix + parent)
                if fr is None:
                    raise KeyError(f"Edit of unknown frame {name!r} in {path}")
                if parent is not None:
                    fr.parent = parent if self.find(parent) else self.prefix + parent
                fr.attrs.update(attrs)
            else:
                self.frames.append(GFrame(self.prefix + name, None if parent is None else self.prefix + parent, attrs))


def load_g(path: str) -> List[GFrame]:
    """Frames of a `.g` file (includes resolved, prefixes and edits applied), in file order."""
    p = _Parser()
    p.parse_file(path)
    return p.frames


def _parents_first(frames: List[GFrame]) -> List[GFrame]:
    """stable topological order (edits may re-parent a frame under one defined later)"""
    done, order, pending = set(), [], list(frames)
    all_names = {f.name for f in frames}
    while pending:
        rest = []
        for f in pending:
            if f.parent is None or f.parent in done or f.parent not in all_names:
                order.append(f)
                done.add(f.name)
            else:
                rest.append(f)
        if len(rest) == len(pending):
            raise ValueError("cycle in .g frame tree: " + ", ".join(f.name for f in rest))
        pending = rest
    return order


def _pose(v) -> Tf:
    if isinstance(v, str):
        return Tf.parse(v)
    return Tf.from_pose([float(x) for x in v])


def _limits(v, dof: int) -> Optional[np.ndarray]:
    """rai stores [lo, hi] per dof first (UR files append velocity / effort entries)."""
    if v is None or dof == 0:
        return None
    a = [float(x) for x in v]
    return np.asarray(a[:2 * dof], np.float64).reshape(dof, 2)


def add_g_model(sc: Scene, path: str, prefix: str, parent: Optional[str], root_rel: Optional[Tf], robot: str,
                root_joint: Optional[str] = "rigid", q0: Optional[Dict[str, float]] = None) -> List[str]:
    """`C.addFile(path, namePrefix=prefix)` followed by re-parenting the model's root frame under `parent` with
    relative pose `root_rel` and a rigid joint (rai_config.py:2962-2971).  With parent=None the root keeps its
    own `X` pose unless `root_rel` overrides it (rai_config.py:7708 moves the mobile bases' `world` frames).
    Mesh / marker frames and shapes without a `contact` flag never collide (rai_base_env.py:234-255), so only
    their frames are kept.  Returns the names of the frames added."""
    frames = _parents_first(load_g(path))
    names = {}
    added = []
    roots = [f for f in frames if f.parent is None]
    # a model file has one kinematic root; extra parent-less frames that nothing hangs on are bookkeeping
    # nodes (e.g. `robotiq_base: {}` before it is edited under the arm) and are skipped
    used_as_parent = {f.parent for f in frames}
    for fr in frames:
        a = fr.attrs
        name = prefix + fr.name
        if name in sc.frames:  # a second frame of the same name (marker under `gripper`): never collides, skip
            continue
        jt = a.get("joint")
        if a.get("joint_active") is False:
            jt = None
        if jt is not None and jt not in JOINT_DOF:
            raise ValueError(f"unsupported joint type {jt!r} on {fr.name}")
        dof = JOINT_DOF.get(jt, 0) if jt else 0
        if jt:
            rel = _pose(a["A"]) if "A" in a else Tf()
        else:
            rel = _pose(a["Q"]) if "Q" in a else (_pose(a["rel"]) if "rel" in a else Tf())
        par = None if fr.parent is None else prefix + fr.parent
        if fr.parent is None:
            if fr not in roots[:1] and fr.name not in used_as_parent:
                continue
            if "X" in a:
                rel = _pose(a["X"])
            if fr is roots[0]:
                par = parent
                if root_rel is not None:
                    rel = root_rel
                if parent is not None and jt is None:
                    jt = root_joint
        shape = a.get("shape")
        contact = int(a.get("contact", 0) or 0)
        if shape not in PRIMITIVES or contact == 0:
            shape, contact = None, 0
        home = None
        if dof:
            if q0 and fr.name in q0:
                home = np.atleast_1d(np.asarray(q0[fr.name], np.float64))
            elif "q" in a:
                home = np.atleast_1d(np.asarray(a["q"], np.float64))
        sc.add(name, par, rel=rel, joint=jt, limits=_limits(a.get("limits"), dof), q0=home, shape=shape,
               size=None if shape is None else [float(x) for x in a["size"]], contact=contact,
               robot=robot if dof else None)
        names[fr.name] = name
        added.append(name)
    return added
