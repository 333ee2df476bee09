"""Host-side scene model and scene-blob compiler.

A `Scene` is a tree of frames (rai `.g` semantics: X_frame = X_parent * rel * J(q)),
some of which carry a collision primitive.  `compile_blob` flattens it, for one mode,
into the flat word array the CUDA kernels stage into shared memory (layout in
csrc/scene_blob.h, mirrored here).  Nothing in this file evaluates collisions.

Reference semantics restated (paths relative to /root/reference,
P/ = src/multi_robot_multi_goal_planning/):
  * frame / joint conventions, shapes, `contact` flag: the `.g` models under
    P/assets/models/rai/ (ur10/ur10.g:9-39, mobile-manipulator-restricted.g:1-68) and
    the scene builders P/problems/rai/rai_config.py:65-100, 751-839, 2947-3064,
    3319-3513, 7690-7768.
  * visual-only (mesh / contact-less) frames never collide: P/problems/rai_base_env.py:234-255.
  * mode relinking (`attach`, `setContact(-1)`): P/problems/rai_base_env.py:776-810.
  * the flag rule "sum of penetrations > tolerance": P/problems/rai_base_env.py:460-477.
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

# ----------------------------------------------------------------------------------
# small fp64 transform helpers (rotation matrix + translation)
# ----------------------------------------------------------------------------------


def quat_to_mat(q: Sequence[float]) -> np.ndarray:
    """[w,x,y,z] -> 3x3 (normalised first, as rai does on read)."""
    w, x, y, z = np.asarray(q, np.float64) / np.linalg.norm(q)
    return np.array([
        [1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
        [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
        [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)],
    ])


def mat_to_quat(R: np.ndarray) -> np.ndarray:
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = [0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s]
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(1.0 + R[i, i] - R[j, j] - R[k, k]) * 2
        q = [0.0] * 4
        q[0] = (R[k, j] - R[j, k]) / s
        q[1 + i] = 0.25 * s
        q[1 + j] = (R[j, i] + R[i, j]) / s
        q[1 + k] = (R[k, i] + R[i, k]) / s
    q = np.array(q)
    return q if q[0] >= 0 else -q


def axis_angle_mat(axis: Sequence[float], rad: float) -> np.ndarray:
    a = np.asarray(axis, np.float64)
    a = a / np.linalg.norm(a)
    c, s = math.cos(rad), math.sin(rad)
    K = np.array([[0, -a[2], a[1]], [a[2], 0, -a[0]], [-a[1], a[0], 0]])
    return np.eye(3) + s * K + (1 - c) * (K @ K)


class Tf:
    """Rigid transform (R, t)."""

    __slots__ = ("R", "t")

    def __init__(self, R=None, t=None):
        self.R = np.eye(3) if R is None else np.asarray(R, np.float64)
        self.t = np.zeros(3) if t is None else np.asarray(t, np.float64)

    def __matmul__(self, o: "Tf") -> "Tf":
        return Tf(self.R @ o.R, self.R @ o.t + self.t)

    def apply(self, p) -> np.ndarray:
        return self.R @ np.asarray(p, np.float64) + self.t

    def inv(self) -> "Tf":
        return Tf(self.R.T, -self.R.T @ self.t)

    @staticmethod
    def from_pose(p: Sequence[float]) -> "Tf":
        """[x,y,z] or [x,y,z,qw,qx,qy,qz]"""
        p = list(p)
        if len(p) == 3:
            return Tf(None, p)
        return Tf(quat_to_mat(p[3:7]), p[:3])

    @staticmethod
    def parse(spec: str) -> "Tf":
        """rai transformation mini-language: sequence of t(x y z) / d(deg ax ay az) /
        r(rad ax ay az) / q(w x y z) / T, each applied *relative* to the running frame."""
        import re
        X = Tf()
        for tag, args in re.findall(r"([tdrqET])\s*(?:\(([^)]*)\))?", spec):
            if tag == "T":
                continue
            v = [float(a) for a in args.replace(",", " ").split()]
            if tag == "t":
                X = X @ Tf(None, v)
            elif tag == "d":
                X = X @ Tf(axis_angle_mat(v[1:4], math.radians(v[0])), None)
            elif tag == "r":
                X = X @ Tf(axis_angle_mat(v[1:4], v[0]), None)
            elif tag == "q":
                X = X @ Tf(quat_to_mat(v), None)
            else:
                raise ValueError(tag)
        return X


JOINT_DOF = {"hingeX": 1, "hingeY": 1, "hingeZ": 1, "transXYPhi": 3, "transX": 1, "transY": 1,
             "transZ": 1, "rigid": 0}
JOINT_CODE = {"hingeX": 1, "hingeY": 2, "hingeZ": 3, "transXYPhi": 4, "transX": 5, "transY": 6, "transZ": 7}

CORE_POINT, CORE_SEG, CORE_BOX, CORE_CYLZ = 0, 1, 2, 3


def joint_tf(jtype: str, q: np.ndarray) -> Tf:
    if jtype == "hingeX":
        return Tf(axis_angle_mat([1, 0, 0], q[0]))
    if jtype == "hingeY":
        return Tf(axis_angle_mat([0, 1, 0], q[0]))
    if jtype == "hingeZ":
        return Tf(axis_angle_mat([0, 0, 1], q[0]))
    if jtype == "transXYPhi":
        return Tf(axis_angle_mat([0, 0, 1], q[2]), [q[0], q[1], 0.0])
    if jtype == "transX":
        return Tf(None, [q[0], 0, 0])
    if jtype == "transY":
        return Tf(None, [0, q[0], 0])
    if jtype == "transZ":
        return Tf(None, [0, 0, q[0]])
    return Tf()


@dataclass
class Shape:
    """rai shape conventions: sphere [r]; capsule / cylinder [length, r] along local z;
    box [x,y,z(,ignored)]; ssBox [x,y,z,r] = box of OUTER size x,y,z rounded by r."""
    kind: str
    size: Tuple[float, ...]

    def core(self, planar_z: bool = False):
        """-> (core_type, radius, local data): every primitive is a sphere-swept core, except
        upright cylinders in planar scenes (`planar_z`: the axis stays parallel to world z for
        every configuration), which are exact z-prisms (disc x interval).  Cylinders in general
        position are modelled as capsules with the same radius and a core segment of the
        cylinder's length (a superset of the cylinder; rai itself uses a convex mesh)."""
        k, s = self.kind, self.size
        if k == "sphere":
            return CORE_POINT, float(s[0]), {}
        if k == "cylinder" and planar_z:
            return CORE_CYLZ, 0.0, {"cyl_r": float(s[1]), "cyl_h": float(s[0]) / 2}
        if k in ("capsule", "cylinder"):
            return CORE_SEG, float(s[1]), {"half_len": float(s[0]) / 2}
        if k == "box":
            return CORE_BOX, 0.0, {"half": np.asarray(s[:3], np.float64) / 2}
        if k == "ssBox":
            r = float(s[3])
            return CORE_BOX, r, {"half": np.asarray(s[:3], np.float64) / 2 - r}
        raise ValueError(k)


@dataclass
class Frame:
    name: str
    parent: Optional[str]
    rel: Tf = field(default_factory=Tf)
    joint: Optional[str] = None          # None = plain frame; "rigid" = link boundary w/o dof
    limits: Optional[np.ndarray] = None  # (dof, 2)
    q0: Optional[np.ndarray] = None
    shape: Optional[Shape] = None
    contact: int = 0
    robot: Optional[str] = None          # joint frames: which robot owns these dofs


class Scene:
    """Ordered frame tree (parents are always added before children)."""

    def __init__(self):
        self.frames: Dict[str, Frame] = {}
        self.robots: List[str] = []

    # ---- construction -----------------------------------------------------------
    def add(self, name, parent=None, rel=None, joint=None, limits=None, q0=None, shape=None,
            size=None, contact=0, robot=None) -> Frame:
        assert name not in self.frames, name
        assert parent is None or parent in self.frames, parent
        if isinstance(rel, str):
            rel = Tf.parse(rel)
        elif rel is not None and not isinstance(rel, Tf):
            rel = Tf.from_pose(rel)
        dof = JOINT_DOF.get(joint, 0) if joint else 0
        f = Frame(name, parent, rel or Tf(), joint,
                  None if limits is None else np.asarray(limits, np.float64).reshape(dof, 2),
                  None if q0 is None else np.asarray(q0, np.float64).reshape(dof),
                  None if shape is None else Shape(shape, tuple(size)), contact, robot)
        if dof and f.q0 is None:
            f.q0 = np.zeros(dof)
        self.frames[name] = f
        if robot is not None and dof and robot not in self.robots:
            self.robots.append(robot)
        return f

    def copy(self) -> "Scene":
        """Independent frame table; transforms, limits and shapes are shared (they are never mutated in place:
        attach / remove replace the `rel` object and rebuild the table)."""
        sc = Scene()
        sc.frames = {n: copy.copy(f) for n, f in self.frames.items()}
        sc.robots = list(self.robots)
        return sc

    # ---- joint vector layout ----------------------------------------------------
    def dof_frames(self) -> List[Frame]:
        return [f for f in self.frames.values() if f.joint and JOINT_DOF[f.joint] > 0]

    def q_layout(self) -> Dict[str, Tuple[int, int]]:
        """joint-frame name -> [start, end) in the flat configuration.  Robots' dofs are
        contiguous and in `self.robots` order (like C.getJointState() for scenes built
        robot after robot, rai_base_env.py:214-231)."""
        out, s = {}, 0
        for r in self.robots:
            for f in self.dof_frames():
                if f.robot == r:
                    n = JOINT_DOF[f.joint]
                    out[f.name] = (s, s + n)
                    s += n
        for f in self.dof_frames():
            assert f.name in out, f"joint {f.name} has no robot"
        return out

    @property
    def dof(self) -> int:
        return sum(JOINT_DOF[f.joint] for f in self.dof_frames())

    def robot_slices(self) -> Dict[str, Tuple[int, int]]:
        lay = self.q_layout()
        out = {}
        for r in self.robots:
            idx = [lay[f.name] for f in self.dof_frames() if f.robot == r]
            out[r] = (min(a for a, _ in idx), max(b for _, b in idx))
        return out

    def limits(self) -> np.ndarray:
        lay = self.q_layout()
        lim = np.zeros((2, self.dof))
        for f in self.dof_frames():
            s, e = lay[f.name]
            lim[0, s:e] = f.limits[:, 0]
            lim[1, s:e] = f.limits[:, 1]
        return lim

    def home(self) -> np.ndarray:
        lay = self.q_layout()
        q = np.zeros(self.dof)
        for f in self.dof_frames():
            s, e = lay[f.name]
            q[s:e] = f.q0
        return q

    # ---- forward kinematics on the host (fp64; used for relinking and tests) -----
    def fk(self, q: np.ndarray) -> Dict[str, Tf]:
        lay = self.q_layout()
        X: Dict[str, Tf] = {}
        for f in self.frames.values():
            P = X[f.parent] if f.parent is not None else Tf()
            T = P @ f.rel
            if f.name in lay:
                s, e = lay[f.name]
                T = T @ joint_tf(f.joint, np.asarray(q[s:e], np.float64))
            X[f.name] = T
        return X

    # ---- relinking (mode changes) -----------------------------------------------
    def attach(self, parent: str, child: str, q: np.ndarray) -> None:
        """rai `C.attach(parent, child)` at configuration q: child keeps its world pose,
        hangs on `parent` through a rigid joint; then `setContact(-1)` on the child
        (rai_base_env.py:801-805)."""
        X = self.fk(q)
        rel = X[parent].inv() @ X[child]
        f = self.frames[child]
        f.parent, f.rel, f.joint, f.contact = parent, rel, "rigid", -1
        # keep parents-before-children order
        fr = self.frames.pop(child)
        self.frames[child] = fr
        self._reorder()

    def _reorder(self):
        done, order = set(), []
        pending = list(self.frames.values())
        while pending:
            rest = []
            for f in pending:
                if f.parent is None or f.parent in done:
                    order.append(f)
                    done.add(f.name)
                else:
                    rest.append(f)
            assert len(rest) < len(pending), "cycle in frame tree"
            pending = rest
        self.frames = {f.name: f for f in order}

    def remove(self, name: str) -> None:
        kids = [f.name for f in self.frames.values() if f.parent == name]
        for k in kids:
            self.remove(k)
        del self.frames[name]

    # ---- link structure and the collidable-pair rule ------------------------------
    def link_of(self, name: str) -> str:
        """rai getUpwardLink(): nearest ancestor-or-self that carries a joint (any type,
        rigid included), else the root."""
        f = self.frames[name]
        while f.joint is None and f.parent is not None:
            f = self.frames[f.parent]
        return f.name

    def parent_link(self, link: str) -> Optional[str]:
        p = self.frames[link].parent
        return None if p is None else self.link_of(p)

    def can_collide(self, a: str, b: str) -> bool:
        """Restatement of rai's Shape::canCollideWith (source not in the reference repo;
        SURVEY.md 8a): contact 0 never collides; shapes on the same link never collide;
        contact = -k additionally suppresses the pair when the other shape's link is one
        of the first k ancestor links."""
        fa, fb = self.frames[a], self.frames[b]
        if not fa.shape or not fb.shape or fa.contact == 0 or fb.contact == 0:
            return False
        la, lb = self.link_of(a), self.link_of(b)
        if la == lb:
            return False
        for (l1, c1, l2) in ((la, fa.contact, lb), (lb, fb.contact, la)):
            if c1 < 0:
                p = l1
                for _ in range(-c1):
                    p = self.parent_link(p)
                    if p is None:
                        break
                    if p == l2:
                        return False
        return True

    def collision_shapes(self) -> List[str]:
        return [f.name for f in self.frames.values() if f.shape is not None and f.contact != 0]

    def collidable_pairs(self) -> List[Tuple[str, str]]:
        """all (a, b) with can_collide(a, b), a before b in frame order (same rule, link tables computed once)"""
        names = self.collision_shapes()
        link = {n: self.link_of(n) for n in names}
        up: Dict[str, Optional[str]] = {}

        def parent_link(l):
            if l not in up:
                up[l] = self.parent_link(l)
            return up[l]

        anc = {}   # shape -> the ancestor links its negative contact flag suppresses
        for n in names:
            k, p, out = -self.frames[n].contact, link[n], set()
            for _ in range(max(k, 0)):
                p = parent_link(p)
                if p is None:
                    break
                out.add(p)
            anc[n] = out
        pairs = []
        for i, a in enumerate(names):
            la, sa = link[a], anc[a]
            for b in names[i + 1:]:
                lb = link[b]
                if la != lb and lb not in sa and la not in anc[b]:
                    pairs.append((a, b))
        return pairs

    def is_moving(self, name: str) -> bool:
        f = self.frames[name]
        while True:
            if f.joint and JOINT_DOF[f.joint] > 0:
                return True
            if f.parent is None:
                return False
            f = self.frames[f.parent]

    def robot_of_shape(self, name: str) -> Optional[str]:
        f = self.frames[name]
        while True:
            if f.joint and JOINT_DOF[f.joint] > 0:
                return f.robot
            if f.parent is None:
                return None
            f = self.frames[f.parent]


# ----------------------------------------------------------------------------------
# blob layout (mirror of csrc/scene_blob.h)
# ----------------------------------------------------------------------------------
BLOB_MAGIC = 0x4D524232
BLOB_VERSION = 7
HDR_WORDS = 136
FRAME_WORDS = 16
SHAPE_WORDS = 20
LARGE_BOX_BOUND = 0.3  # boxes with a bounding radius above this use the face-normal broadphase bound
NUM_PAIR_TYPES = 8   # (point,point) (point,seg) (seg,seg) (point,box) (seg,box) (box,box) (cylz,cylz) (box,cylz)
PAIR_TYPE = {(0, 0): 0, (0, 1): 1, (1, 1): 2, (0, 2): 3, (1, 2): 4, (2, 2): 5, (3, 3): 6, (2, 3): 7}
DIRECT_TYPES = (6, 7)   # planar types the kernels evaluate straight from their pair lists (no broadphase records)
PAIR_TYPE_NAMES = ["point-point", "point-seg", "seg-seg", "point-box", "seg-box", "box-box", "cylz-cylz", "box-cylz"]
WORLD_WORDS = {CORE_POINT: 3, CORE_SEG: 6, CORE_BOX: 12, CORE_CYLZ: 3}
# header word indices
H_MAGIC, H_VERSION, H_DOF, H_NFRAMES, H_NMOV, H_NSTA, H_WORLD_WORDS, H_NCHAINS = range(8)
H_OFF_FRAMES, H_OFF_SHAPES, H_OFF_CHAINS, H_OFF_STATIC_PAIRS, H_N_STATIC_PAIRS = 8, 9, 10, 11, 12
H_TOL, H_STATIC_PEN, H_TOTAL_WORDS, H_NROBOTS = 13, 14, 15, 16
H_STAGED_WORDS = 17  # the kernels copy only this prefix into shared memory; the tail is read from global memory
H_REC_BASE = 18      # first word of the broadphase records
H_IDS_BASE = 19      # first word of the packed pair ids, one per record, in record order
H_IDS_STAGED = 38    # 1: the pair ids are part of the staged prefix
H_GPTR = 40          # [40..41] device only: the kernels park the blob's global address here after staging
H_BP_IDS = 112       # [type][sublist] -> first word of the sublist's pair ids (8 x 3 words)
H_OFF_PAIRS = 20   # [20..27]
H_N_PAIRS = 28     # [28..35]
H_OFF_SHAPE_ROBOT = 36
H_OFF_SCENTRE = 37   # static shape centres, 4 floats each
H_BP = 48            # broadphase sublists: [type][sublist] -> (record offset, count), 8 x 3 x 2 words
BP_SUBLISTS = 3      # 0: partner moving (bounding spheres), 1: partner static (bounding spheres),
                     # 2: partner is a large static box (face-normal separating-axis bound)
STAGE_IDS_MAX_BYTES = 12 * 1024   # blobs whose staged prefix stays below this keep the records' pair ids in it
CULL_SLACK = 1e-3    # broadphase keeps everything closer than 1 mm  # per-shape robot id (moving shapes: owning robot; static: -1; held: holder)


@dataclass
class CompiledScene:
    """Result of compile_blob: word arrays (u32 view) for the device (fp32) and for the
    fp64 oracle, plus bookkeeping the host/bench needs."""
    blob32: np.ndarray            # uint32
    blob64: np.ndarray            # uint64 (ints stored as uint64, floats as float64 bits)
    dof: int
    staged_words: int             # prefix of the blob the kernels keep in shared memory
    n_frames: int
    n_moving: int
    n_static: int
    world_words: int
    pair_counts: List[int]
    static_pair_count: int
    shape_names: List[str]
    pairs: List[Tuple[str, str]]
    tol: float
    unreachable_pairs: List[Tuple[str, str]] = field(default_factory=list)  # collidable, but out of each other's reach


_CHAIN_APART_CACHE: Dict[tuple, bool] = {}   # compile_blob: chain-neighbour pairs already analysed (the same for every mode)


def _seg_seg_dist(p1, q1, p2, q2):
    """distance between segments [p1, q1] and [p2, q2], row-wise ([n, 3] arrays; degenerate segments are points):
    closest points by clamping the unconstrained solution (Ericson, Real-Time Collision Detection 5.1.9)"""
    d1, d2, r = q1 - p1, q2 - p2, p1 - p2
    a, e, f = (d1 * d1).sum(1), (d2 * d2).sum(1), (d2 * r).sum(1)
    c, b = (d1 * r).sum(1), (d1 * d2).sum(1)
    eps = 1e-18
    denom = a * e - b * b
    s_ = np.where(denom > eps, np.clip((b * f - c * e) / np.where(denom > eps, denom, 1.0), 0.0, 1.0), 0.0)
    s_ = np.where(a > eps, s_, 0.0)
    t_ = np.where(e > eps, (b * s_ + f) / np.where(e > eps, e, 1.0), 0.0)
    s_lo = np.where(a > eps, np.clip(-c / np.where(a > eps, a, 1.0), 0.0, 1.0), 0.0)
    s_hi = np.where(a > eps, np.clip((b - c) / np.where(a > eps, a, 1.0), 0.0, 1.0), 0.0)
    s_ = np.where(t_ < 0.0, s_lo, np.where(t_ > 1.0, s_hi, s_))
    t_ = np.clip(t_, 0.0, 1.0)
    c1 = p1 + d1 * s_[:, None]
    c2 = p2 + d2 * t_[:, None]
    return np.linalg.norm(c1 - c2, axis=1)


def compile_blob(scene: Scene, tol: float) -> CompiledScene:
    lay = scene.q_layout()
    X0 = scene.fk(scene.home())  # static frames: any q works

    # --- kinematic chains: one per robot, strictly serial dof-joint paths ---
    dof_frames = scene.dof_frames()

    def dof_ancestor(name: str, inclusive: bool) -> Optional[str]:
        f = scene.frames[name]
        if not inclusive:
            f = scene.frames[f.parent] if f.parent else None
        while f is not None:
            if f.joint and JOINT_DOF[f.joint] > 0:
                return f.name
            f = scene.frames[f.parent] if f.parent else None
        return None

    def chain_rel(from_excl: Optional[str], to_incl: str) -> Tf:
        """product of `rel` from below `from_excl` down to and including `to_incl`
        (joint motions of intermediate frames are rigid/identity by construction)."""
        path = []
        f = scene.frames[to_incl]
        while f is not None and f.name != from_excl:
            path.append(f)
            f = scene.frames[f.parent] if f.parent else None
        T = Tf()
        for fr in reversed(path):
            T = T @ fr.rel
        return T

    chains: List[List[str]] = []
    for r in scene.robots:
        js = [f.name for f in dof_frames if f.robot == r]
        for i, j in enumerate(js):
            anc = dof_ancestor(j, inclusive=False)
            expect = js[i - 1] if i else None
            assert anc == expect, (
                f"robot {r}: joint {j} hangs on {anc}, expected {expect} -- only serial chains "
                "whose base is static are supported")
        chains.append(js)
    frame_ids: Dict[str, int] = {}
    frame_rows = []
    chain_rows = []
    for js in chains:
        start = len(frame_rows)
        for i, j in enumerate(js):
            f = scene.frames[j]
            A = chain_rel(js[i - 1] if i else None, j)
            frame_ids[j] = len(frame_rows)
            frame_rows.append((-1 if i == 0 else frame_ids[js[i - 1]], JOINT_CODE[f.joint], lay[j][0], A))
        chain_rows.append((start, len(frame_rows)))

    # --- shapes: moving ones sorted by frame id (FK emits them while walking the chain) ---
    def planar_z(name: str) -> bool:
        """True iff the frame's z axis is world z for every configuration."""
        f = scene.frames[name]
        while f is not None:
            R = f.rel.R
            if abs(R[2, 2] - 1) > 1e-9 or abs(R[0, 2]) > 1e-9 or abs(R[1, 2]) > 1e-9:
                return False
            if f.joint in ("hingeX", "hingeY"):
                return False
            f = scene.frames[f.parent] if f.parent else None
        return True

    robot_ids = {r: i for i, r in enumerate(scene.robots)}
    mov, sta = [], []
    for name in scene.collision_shapes():
        f = scene.frames[name]
        core, rad, extra = f.shape.core(planar_z(name))
        j = dof_ancestor(name, inclusive=True)
        if j is not None:
            L = chain_rel(j, name) if name != j else Tf()
            mov.append((frame_ids[j], name, core, rad, extra, L, robot_ids[scene.frames[j].robot]))
        else:
            sta.append((-1, name, core, rad, extra, X0[name], -1))
    mov.sort(key=lambda s: s[0])
    shapes = mov + sta
    shape_idx = {s[1]: i for i, s in enumerate(shapes)}
    n_mov, n_sta = len(mov), len(sta)

    def shape_data(core, extra, T: Tf) -> List[float]:
        if core == CORE_POINT:
            return list(T.t)
        if core == CORE_SEG:
            h = extra["half_len"]
            return list(T.apply([0, 0, -h])) + list(T.apply([0, 0, h]))
        if core == CORE_CYLZ:
            return list(T.t) + [extra["cyl_r"], extra["cyl_h"]]
        return list(T.t) + list(T.R.reshape(-1)) + list(extra["half"])

    def bound_radius(core, rad, extra) -> float:
        """radius of the bounding sphere around the shape's centre (segment midpoint, box centre)"""
        if core == CORE_SEG:
            return extra["half_len"] + rad
        if core == CORE_BOX:
            return float(np.linalg.norm(extra["half"])) + rad
        if core == CORE_CYLZ:
            return math.hypot(extra["cyl_r"], extra["cyl_h"])
        return rad

    woff = 0
    shape_rows = []
    bound_rs = []
    for (fid, name, core, rad, extra, T, rob) in shapes:
        data = shape_data(core, extra, T)
        data = data + [0.0] * (15 - len(data)) + [bound_radius(core, rad, extra)]   # data[15] = bounding radius
        bound_rs.append(data[15])
        shape_rows.append((core, fid, woff if fid >= 0 else -1, rad, data, rob))
        if fid >= 0:
            woff += WORLD_WORDS[core]
    world_words = woff

    # --- pairs by core-type; static-static pairs kept apart (constant per mode) ---
    typed: List[List[Tuple[int, int]]] = [[] for _ in range(NUM_PAIR_TYPES)]
    static_pairs: List[Tuple[int, int, int]] = []
    all_pairs = scene.collidable_pairs()
    for a, b in all_pairs:
        ia, ib = shape_idx[a], shape_idx[b]
        ca, cb = shapes[ia][2], shapes[ib][2]
        if ca > cb:  # order so that core(a) <= core(b)
            ia, ib, ca, cb = ib, ia, cb, ca
        if (ca, cb) not in PAIR_TYPE:
            raise NotImplementedError(f"pair {a}-{b}: an upright cylinder can only meet cylinders and boxes")
        t = PAIR_TYPE[(ca, cb)]
        if t == 7 and not (planar_z(shapes[ia][1]) and shapes[ia][3] == 0.0):
            raise NotImplementedError(f"pair {a}-{b}: upright cylinder vs a tilted or rounded box")
        # broadphase kind (bits 28..29 of the packed pair): 0 = bounding spheres; when b is a box too
        # large for its bounding sphere to cull anything, a separating-axis bound along b's face
        # normals: 1 = a's bounding sphere, 2 = a's core segment
        kind = 0
        if t == 5 and bound_rs[ia] > bound_rs[ib]:
            ia, ib = ib, ia  # box-box: the larger box is b
        if cb == CORE_BOX and bound_rs[ib] > LARGE_BOX_BOUND and t in (3, 4, 5):
            kind = 2 if ca == CORE_SEG else 1
        if ia >= n_mov and ib >= n_mov:
            static_pairs.append((t, ia, ib))
        else:
            typed[t].append((ia, ib, kind))
    for t in range(NUM_PAIR_TYPES):
        typed[t].sort()

    # --- broadphase records: per type three sublists of (X = a moving shape, Y = its partner) ---
    # record = 2 words: (byte offset of X in a W row) | (Y id << 16), float threshold
    #   sublist 0: Y moving, id = byte offset of Y;      thr = (bound_X + bound_Y + slack)^2
    #   sublist 1: Y static, id = static shape index;    thr = (bound_X + bound_Y + slack)^2
    #   sublist 2: Y large static box, id = static index; thr = r_X + r_Y + slack (X a segment: bound on
    #              its core along the box face normals) or bound_X + r_Y + slack (otherwise)
    # each record is followed (in a parallel array) by the pair's packed shape ids for the narrowphase
    def wbytes(i):
        assert 0 <= shape_rows[i][2] * 128 < 65536, "world data of a configuration exceeds 64 KB"
        return shape_rows[i][2] * 128

    # --- reach of every moving shape: a world-space ball that contains the centre of its bounding sphere for
    # EVERY joint vector (hinge angles are not assumed to respect their limits: the queries are not validated
    # against them either).  Walking up the chain, a hinge about axis a turns ball (c, r) into the ball around
    # the axial part of c with radius r + |radial part of c|; a translating joint makes the reach unbounded.
    # A pair whose reach balls, grown by the bounding radii, cannot meet never collides: it stays in the pair
    # lists (the oracle evaluates it, the algorithmic flop count includes it) but gets no broadphase record.
    HINGE_AXIS = {"hingeX": np.array([1.0, 0, 0]), "hingeY": np.array([0, 1.0, 0]), "hingeZ": np.array([0, 0, 1.0])}
    code_joint = {v: k for k, v in JOINT_CODE.items()}
    reach: List[Tuple[np.ndarray, float]] = []
    for i in range(n_mov):
        fid, L = shapes[i][0], shapes[i][5]
        c, r = np.array(L.t, dtype=np.float64), 0.0
        k = fid
        while k >= 0:
            par, jcode, _, A = frame_rows[k]
            jt = code_joint[jcode]
            if jt in HINGE_AXIS:
                a = HINGE_AXIS[jt]
                ax = a * float(a @ c)
                r += float(np.linalg.norm(c - ax))
                c = ax
            else:
                r = math.inf
            c = np.asarray(A.apply(c), dtype=np.float64)
            k = par
        reach.append((c, r + bound_rs[i]))

    any_bounded = any(math.isfinite(r) for _, r in reach)   # scenes of mobile bases: nothing to prune, skip the tests
    reach_c = [tuple(float(v) for v in c) for c, _ in reach]  # plain floats: this runs once per pair and per mode
    static_box = {}                                           # static index -> (centre, rows of R^T, half extents)
    static_ctr = {}

    # --- vertical separation: a shape whose chain only moves in the plane (translations along x / y, rotations about a z axis
    # that stays world z) keeps the world z of every one of its points.  If the z intervals of two such shapes (static shapes
    # included) are apart, a horizontal plane separates them for every joint vector: the mobile bases ride 4 cm above the
    # floor slab, and that box-box pair passed the bounding test for EVERY configuration (4 of the 5.4 box-box narrowphase
    # items per configuration of the mobile scene, the most expensive routine).
    def const_z(name: str) -> bool:
        f = scene.frames[name]
        while f is not None:
            R = f.rel.R
            if abs(R[2, 2] - 1) > 1e-9 or abs(R[0, 2]) > 1e-9 or abs(R[1, 2]) > 1e-9 or abs(R[2, 0]) > 1e-9 or abs(R[2, 1]) > 1e-9:
                return False
            if f.joint and JOINT_DOF[f.joint] > 0 and f.joint not in ("transX", "transY", "transXYPhi", "hingeZ"):
                return False
            f = scene.frames[f.parent] if f.parent else None
        return True

    def z_interval(i: int) -> Optional[Tuple[float, float]]:
        fid, name, core, rad, extra, _, _ = shapes[i]
        if fid >= 0 and not const_z(name):
            return None
        T = X0[name]
        if core == CORE_POINT:
            return float(T.t[2]) - rad, float(T.t[2]) + rad
        if core == CORE_SEG:
            za, zb = float(T.apply([0, 0, -extra["half_len"]])[2]), float(T.apply([0, 0, extra["half_len"]])[2])
            return min(za, zb) - rad, max(za, zb) + rad
        if core == CORE_CYLZ:
            return float(T.t[2]) - extra["cyl_h"], float(T.t[2]) + extra["cyl_h"]
        e = float(sum(abs(T.R[2, k]) * extra["half"][k] for k in range(3))) + rad
        return float(T.t[2]) - e, float(T.t[2]) + e

    z_iv = [z_interval(i) for i in range(len(shapes))]

    def vertically_apart(x: int, y: int) -> bool:
        a, b = z_iv[x], z_iv[y]
        return a is not None and b is not None and (a[0] - b[1] > 4 * CULL_SLACK or b[0] - a[1] > 4 * CULL_SLACK)

    # --- neighbours on a chain: a capsule / sphere pair whose relative pose depends on at most two hinge angles (a link
    # against the link two joints up, a wrist capsule against the gripper body, the first links against shapes on the robot's
    # static base) is evaluated EXACTLY on a grid over the full circle of those angles; with the Lipschitz bound of the
    # motion between grid points its distance stays positive for every joint vector, so it gets no record.  Such pairs
    # pass the bounding-sphere test for (nearly) every configuration -- 2 per arm on the UR10 scenes -- and each one is a
    # narrowphase item per configuration that can never contribute.
    chain_of = {}
    for ci, (c0, c1) in enumerate(chain_rows):
        for f in range(c0, c1):
            chain_of[f] = (ci, c0)

    _fkeys: Dict[int, tuple] = {}
    _skeys: Dict[int, tuple] = {}

    def frame_key(k: int) -> tuple:
        if k not in _fkeys:
            A = frame_rows[k][3]
            _fkeys[k] = (frame_rows[k][1],) + tuple(round(float(v), 9) for v in A.R.reshape(-1)) + tuple(round(float(v), 9) for v in A.t)
        return _fkeys[k]

    def shape_key(i: int) -> tuple:
        if i not in _skeys:
            core, _, _, rad, data, _ = shape_rows[i]
            _skeys[i] = (core, round(rad, 9)) + tuple(round(float(v), 9) for v in data[:6])
        return _skeys[i]

    def local_points(i: int):
        core, fid, _, rad, data, _ = shape_rows[i]
        if core == CORE_POINT:
            return np.array([data[:3], data[:3]], dtype=np.float64), rad
        if core == CORE_SEG:
            return np.array([data[:3], data[3:6]], dtype=np.float64), rad
        return None, rad

    def chain_apart(x: int, y: int) -> bool:
        fx = shape_rows[x][1]
        fy = shape_rows[y][1] if y < n_mov else -1
        if fy > fx:
            x, y, fx, fy = y, x, fy, fx
        if fx < 0 or shape_rows[x][0] > CORE_SEG or shape_rows[y][0] > CORE_SEG:
            return False
        ci, c0 = chain_of[fx]
        first = c0 if fy < 0 else fy + 1          # joints first .. fx move x relative to y's frame (the world for static y)
        if fx - first > 1 or fx < first or (fy >= 0 and chain_of[fy][0] != ci):
            return False
        joints = list(range(first, fx + 1))
        if any(code_joint[frame_rows[k][1]] not in HINGE_AXIS for k in joints):
            return False
        px, rx_ = local_points(x)
        py, ry_ = local_points(y)
        key = (tuple(frame_key(k) for k in joints), shape_key(x), shape_key(y), fy < 0)
        if key in _CHAIN_APART_CACHE:
            return _CHAIN_APART_CACHE[key]
        step = math.radians(0.25 if len(joints) == 1 else 1.0)
        grids = np.meshgrid(*[np.arange(0.0, 2 * math.pi, step) for _ in joints], indexing="ij")
        th = [g.reshape(-1) for g in grids]
        n = th[0].size
        # pose of frame fx relative to y's frame: product over the joints of A_k * Rot(axis_k, theta_k); the lever arm of
        # a joint = distance of x's end points from the joint's axis point (its frame origin)
        R = np.broadcast_to(np.eye(3), (n, 3, 3)).copy()
        t = np.zeros((n, 3))
        origins = []
        for k, a in zip(joints, th):
            A = frame_rows[k][3]
            t = t + R @ A.t
            R = R @ A.R
            origins.append(t.copy())
            ax = HINGE_AXIS[code_joint[frame_rows[k][1]]]
            c, s_ = np.cos(a), np.sin(a)
            K = np.array([[0, -ax[2], ax[1]], [ax[2], 0, -ax[0]], [-ax[1], ax[0], 0]])
            J = np.eye(3)[None] + s_[:, None, None] * K[None] + (1 - c)[:, None, None] * (K @ K)[None]
            R = R @ J
        a0 = np.einsum("nij,j->ni", R, px[0]) + t
        a1 = np.einsum("nij,j->ni", R, px[1]) + t
        if fy < 0:      # static partner: its data are world coordinates, and so is the chain's root transform
            b0, b1 = py[0][None], py[1][None]
        else:
            b0, b1 = py[0][None], py[1][None]
        d = _seg_seg_dist(a0, a1, np.broadcast_to(b0, a0.shape), np.broadcast_to(b1, a0.shape)) - rx_ - ry_
        lever = sum(float(np.max(np.maximum(np.linalg.norm(a0 - o, axis=1), np.linalg.norm(a1 - o, axis=1)))) + rx_ for o in origins)
        ok = bool(d.min() - 0.5 * step * lever > 4 * CULL_SLACK)
        _CHAIN_APART_CACHE[key] = ok
        return ok

    def never_meets(x: int, y: int, kind: int) -> bool:
        if vertically_apart(x, y):
            return True
        if chain_apart(x, y):
            return True
        if not any_bounded:
            return False
        rx_ = reach[x][1]
        if not math.isfinite(rx_):
            return False
        cx = reach_c[x]
        if y < n_mov:
            ry_ = reach[y][1]
            return math.isfinite(ry_) and math.dist(cx, reach_c[y]) > rx_ + ry_ + 2 * CULL_SLACK
        core, _, _, rad_y, data, _ = shape_rows[y]
        if core == CORE_BOX:  # distance from the reach centre to the static box itself (large tables)
            if y not in static_box:
                R = data[3:12]
                static_box[y] = (data[:3], [(R[0 + k], R[3 + k], R[6 + k]) for k in range(3)], data[12:15])
            ctr, cols, half = static_box[y]
            d = (cx[0] - ctr[0], cx[1] - ctr[1], cx[2] - ctr[2])
            e2 = 0.0
            for k in range(3):
                e = abs(cols[k][0] * d[0] + cols[k][1] * d[1] + cols[k][2] * d[2]) - half[k]
                if e > 0.0:
                    e2 += e * e
            return math.sqrt(e2) > rx_ + rad_y + 2 * CULL_SLACK
        if y not in static_ctr:
            static_ctr[y] = tuple(0.5 * (data[k] + data[3 + k]) for k in range(3)) if core == CORE_SEG else tuple(data[:3])
        return math.dist(cx, static_ctr[y]) > rx_ + bound_rs[y] + 2 * CULL_SLACK

    bp: List[List[List[Tuple[int, float, int]]]] = [[[] for _ in range(BP_SUBLISTS)] for _ in range(NUM_PAIR_TYPES)]
    unreachable: List[Tuple[str, str]] = []
    for t in range(NUM_PAIR_TYPES):   # (the planar direct types get records too; the kernels do not read them yet)
        for (ia, ib, kind) in typed[t]:
            x, y = (ia, ib) if ia < n_mov else (ib, ia)   # X is always a moving shape
            if t <= 5 and never_meets(x, y, kind):
                unreachable.append((shapes[ia][1], shapes[ib][1]))
                continue
            packed = ia | (ib << 16)
            rx, ry = shape_rows[x][3], shape_rows[y][3]
            if y < n_mov:
                bp[t][0].append((wbytes(x) | (wbytes(y) << 16), (bound_rs[x] + bound_rs[y] + CULL_SLACK) ** 2, packed))
            elif kind and y == ib:
                thr = (rx + ry if shape_rows[x][0] == CORE_SEG else bound_rs[x] + ry) + CULL_SLACK
                bp[t][2].append((wbytes(x) | ((y - n_mov) << 16), thr, packed))
            else:
                bp[t][1].append((wbytes(x) | ((y - n_mov) << 16), (bound_rs[x] + bound_rs[y] + CULL_SLACK) ** 2, packed))
        for sl in bp[t]:
            sl.sort(key=lambda r: (r[0] & 0xffff, r[0] >> 16))

    # --- assemble words ---
    # staged prefix (copied into shared memory by every CTA): header, frames, shapes, chains, the pair lists of the
    # directly evaluated planar types, static shape centres, broadphase records.  Tail (global memory only): the pair
    # lists of the queued types (the kernels work from the records; the oracle and the host read these), static-static
    # pairs, per-shape robot ids, and the records' packed pair ids (one read per narrowphase item).
    n_shapes = len(shapes)
    off = HDR_WORDS
    off_frames = off
    off += FRAME_WORDS * len(frame_rows)
    off_shapes = off
    off += SHAPE_WORDS * n_shapes
    off_chains = off
    off += 2 * len(chain_rows)
    off_pairs = [0] * NUM_PAIR_TYPES
    for t in DIRECT_TYPES:
        off_pairs[t] = off
        off += len(typed[t])
    off = (off + 3) // 4 * 4
    off_scentre = off
    off += 4 * n_sta
    off = (off + 1) // 2 * 2              # records are read as 8-byte words
    rec_base = off
    off_bp = [[0] * BP_SUBLISTS for _ in range(NUM_PAIR_TYPES)]
    for t in range(NUM_PAIR_TYPES):
        for k in range(BP_SUBLISTS):
            off_bp[t][k] = off
            off += 2 * len(bp[t][k])
    n_records = (off - rec_base) // 2
    # small scenes keep the pair ids in the staged prefix too (one shared-memory read per narrowphase item instead of
    # a global one: 1.3 % on the dual-arm scene); on large scenes they are what lets a fourth CTA fit on an SM
    ids_staged = (off + n_records) * 4 <= STAGE_IDS_MAX_BYTES
    ids_base = off
    if ids_staged:
        off += n_records
    staged = (off + 3) // 4 * 4           # 16-byte multiple for cp.async.bulk
    off = staged
    for t in range(NUM_PAIR_TYPES):
        if t not in DIRECT_TYPES:
            off_pairs[t] = off
            off += len(typed[t])
    off_static = off
    off += 3 * len(static_pairs)
    off_shape_robot = off
    off += n_shapes
    if not ids_staged:
        ids_base = off
        off += n_records
    total = (off + 3) // 4 * 4

    ints = np.zeros(total, np.int64)
    flts = np.zeros(total, np.float64)
    isf = np.zeros(total, bool)

    def setf(i, v):
        flts[i] = v
        isf[i] = True

    def seti(i, v):
        ints[i] = v

    seti(H_MAGIC, BLOB_MAGIC), seti(H_VERSION, BLOB_VERSION), seti(H_DOF, scene.dof)
    seti(H_NFRAMES, len(frame_rows)), seti(H_NMOV, n_mov), seti(H_NSTA, n_sta)
    seti(H_WORLD_WORDS, world_words), seti(H_NCHAINS, len(chain_rows))
    seti(H_OFF_FRAMES, off_frames), seti(H_OFF_SHAPES, off_shapes), seti(H_OFF_CHAINS, off_chains)
    seti(H_OFF_STATIC_PAIRS, off_static), seti(H_N_STATIC_PAIRS, len(static_pairs))
    setf(H_TOL, tol), setf(H_STATIC_PEN, 0.0), seti(H_TOTAL_WORDS, total)
    seti(H_NROBOTS, len(scene.robots)), seti(H_OFF_SHAPE_ROBOT, off_shape_robot)
    seti(H_STAGED_WORDS, staged), seti(H_REC_BASE, rec_base), seti(H_IDS_BASE, ids_base), seti(H_IDS_STAGED, int(ids_staged))
    for t in range(NUM_PAIR_TYPES):
        seti(H_OFF_PAIRS + t, off_pairs[t]), seti(H_N_PAIRS + t, len(typed[t]))
    for i, (par, jt, qi, A) in enumerate(frame_rows):
        b = off_frames + FRAME_WORDS * i
        sh = [k for k, row in enumerate(shape_rows) if row[1] == i]   # moving shapes are sorted by frame
        assert not sh or sh == list(range(sh[0], sh[0] + len(sh)))
        seti(b, par), seti(b + 1, jt), seti(b + 2, qi), seti(b + 3, (sh[0] if sh else 0) | (len(sh) << 16))
        for k, v in enumerate(list(A.R.reshape(-1)) + list(A.t)):
            setf(b + 4 + k, v)
    for i, (core, fid, wo, rad, data, rob) in enumerate(shape_rows):
        b = off_shapes + SHAPE_WORDS * i
        seti(b, core), seti(b + 1, fid), seti(b + 2, wo), setf(b + 3, rad)
        for k in range(SHAPE_WORDS - 4):
            setf(b + 4 + k, data[k] if k < len(data) else 0.0)
        seti(off_shape_robot + i, rob)
    for i, (s, e) in enumerate(chain_rows):
        seti(off_chains + 2 * i, s), seti(off_chains + 2 * i + 1, e)
    for t in range(NUM_PAIR_TYPES):
        for i, (a, b_, kind) in enumerate(typed[t]):
            seti(off_pairs[t] + i, a | (b_ << 16) | (kind << 28))
    seti(H_OFF_SCENTRE, off_scentre)
    for i in range(n_sta):
        core, _, _, _, data, _ = shape_rows[n_mov + i]
        ctr = [0.5 * (data[k] + data[3 + k]) for k in range(3)] if core == CORE_SEG else data[:3]
        for k in range(3):
            setf(off_scentre + 4 * i + k, ctr[k])
        setf(off_scentre + 4 * i + 3, 0.0)
    for t in range(NUM_PAIR_TYPES):
        for k in range(BP_SUBLISTS):
            n = len(bp[t][k])
            seti(H_BP + (t * BP_SUBLISTS + k) * 2, off_bp[t][k]), seti(H_BP + (t * BP_SUBLISTS + k) * 2 + 1, n)
            seti(H_BP_IDS + t * BP_SUBLISTS + k, ids_base + (off_bp[t][k] - rec_base) // 2)
            for i, (ids, thr, packed) in enumerate(bp[t][k]):
                seti(off_bp[t][k] + 2 * i, ids), setf(off_bp[t][k] + 2 * i + 1, thr)
                seti(ids_base + (off_bp[t][k] - rec_base) // 2 + i, packed)
    for i, (t, a, b_) in enumerate(static_pairs):
        seti(off_static + 3 * i, t), seti(off_static + 3 * i + 1, a), seti(off_static + 3 * i + 2, b_)

    def build(ftype, itype, utype):
        w = ints.astype(itype).view(utype).copy()
        w[isf] = flts[isf].astype(ftype).view(utype)
        return w

    return CompiledScene(
        blob32=build(np.float32, np.int32, np.uint32), blob64=build(np.float64, np.int64, np.uint64), dof=scene.dof,
        staged_words=staged,
        n_frames=len(frame_rows), n_moving=n_mov, n_static=n_sta, world_words=world_words,
        pair_counts=[len(t) for t in typed], static_pair_count=len(static_pairs),
        shape_names=[s[1] for s in shapes], pairs=all_pairs, tol=tol, unreachable_pairs=unreachable)


# Algorithmic flop convention per pair type: used for roofline.achieved only.
PAIR_FLOPS = [12, 30, 90, 30, 220, 350, 12, 30]
FK_FLOPS_PER_FRAME = 85


def algorithmic_flops_per_config(cs: CompiledScene) -> int:
    """W_cfg = W_FK + sum over the mode's collidable pairs of w(type): independent of culling
    and early exit."""
    n_transforms = cs.n_frames + cs.n_moving
    return FK_FLOPS_PER_FRAME * n_transforms + sum(n * w for n, w in zip(cs.pair_counts, PAIR_FLOPS))
