"""Reach analysis of the scene compiler (scene.py compile_blob): pairs that get no broadphase record because the
two shapes can never meet.  Checked against the oracle's FK on joint vectors far outside the limits too -- the
bound must hold for every hinge angle, queries are not validated against limits."""
import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O


def _records(cs):
    b = cs.blob64
    out = set()
    for t in range(6):
        for k in range(S.BP_SUBLISTS):
            off, n = int(b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2]), int(b[S.H_BP + (t * S.BP_SUBLISTS + k) * 2 + 1])
            for i in range(n):
                pk = int(b[int(b[S.H_IDS_BASE]) + (off - int(b[S.H_REC_BASE])) // 2 + i])
                out.add((pk & 0xffff, (pk >> 16) & 0xfff))
    return out


@pytest.mark.parametrize("name", ["box_rearrangement", "box_stacking", "mobile_wall_four"])
def test_unreachable_pairs_never_come_close(name):
    """Every pair without a broadphase record is either out of reach (the bounding spheres -- the distance to the box itself
    for large static boxes -- stay apart by more than the cull slack) or proven apart exactly: neighbours on a chain whose
    relative pose depends on at most two hinge angles, shapes separated by a horizontal plane.  For those the oracle's EXACT
    primitive distance must stay above 4 x the slack.  Joint vectors far outside the limits included."""
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    assert len(cs.unreachable_pairs) > 0.05 * sum(cs.pair_counts)
    idx = {n: i for i, n in enumerate(cs.shape_names)}
    skipped = {frozenset((idx[a], idx[c])) for a, c in cs.unreachable_pairs}
    b = cs.blob64
    ns = cs.n_moving + cs.n_static
    offS = int(b[S.H_OFF_SHAPES])
    rows = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)
    core = rows[:, 0].astype(np.int64)
    rad = rows[:, 3].view(np.float64)
    bound = rows[:, 19].view(np.float64)
    typed = []
    for t in range(6):
        n, off = int(b[S.H_N_PAIRS + t]), int(b[S.H_OFF_PAIRS + t])
        for i in range(n):
            pk = int(b[off + i])
            a, c = pk & 0xffff, (pk >> 16) & 0xfff
            if frozenset((a, c)) in skipped:
                typed.append((t, a, c))
    assert len(typed) == len(skipped)
    rng = np.random.default_rng(0)
    lim = sc.limits()
    qs = np.concatenate([rng.uniform(lim[0], lim[1], (600, sc.dof)), rng.uniform(-7.0, 7.0, (600, sc.dof))])
    sphere_gap = np.full(len(typed), np.inf)
    exact = np.full(len(typed), np.inf)
    for q in qs:
        W = O.world_shapes(b, q, ns)
        ctr = np.where((core == S.CORE_SEG)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])
        for j, (t, ia, ic) in enumerate(typed):
            x, y = (ia, ic) if ia < cs.n_moving else (ic, ia)
            if core[y] == S.CORE_BOX and y >= cs.n_moving:   # large static boxes: distance to the box itself
                R, h = W[y, 3:12].reshape(3, 3), W[y, 12:15]
                d = np.linalg.norm(np.maximum(np.abs(R.T @ (ctr[x] - W[y, :3])) - h, 0)) - bound[x] - rad[y]
            else:
                d = np.linalg.norm(ctr[ia] - ctr[ic]) - bound[ia] - bound[ic]
            sphere_gap[j] = min(sphere_gap[j], d)
            if d <= S.CULL_SLACK:      # (bounding volumes apart at this q: the exact distance is positive anyway)
                exact[j] = min(exact[j], O.pair_distance(t, W[ia], W[ic], rad[ia] + rad[ic]))
    near = sphere_gap <= S.CULL_SLACK
    assert (~near).sum() > 0 and (exact[near] > 4 * S.CULL_SLACK).all(), (
        f"a pair without a record comes within {exact[near].min() if near.any() else None} of touching")
    if name != "mobile_wall_four":
        assert near.sum() >= 4     # the chain-neighbour rule prunes something on the UR10 scenes


@pytest.mark.parametrize("name", list(SCENES))
def test_records_cover_every_reachable_pair(name):
    mk, kw = SCENES[name]
    cs = S.compile_blob(mk(), kw["tol"])
    idx = {n: i for i, n in enumerate(cs.shape_names)}
    recs = _records(cs)
    skipped = {frozenset((idx[a], idx[c])) for a, c in cs.unreachable_pairs}
    b = cs.blob64
    n_queued = 0
    for t in range(6):
        n, off = int(b[S.H_N_PAIRS + t]), int(b[S.H_OFF_PAIRS + t])
        for i in range(n):
            pk = int(b[off + i])
            a, c = pk & 0xffff, (pk >> 16) & 0xfff
            if a >= cs.n_moving and c >= cs.n_moving:
                continue
            n_queued += 1
            assert ((a, c) in recs) != (frozenset((a, c)) in skipped), "every dynamic pair is either recorded or proven unreachable"
    assert n_queued == len(recs) + len(skipped)
    if name == "2d_handover":   # translating bases reach everywhere, every shape lives at the same height
        assert not skipped


def test_vertically_separated_pairs_never_touch():
    """mobile_wall_four: the bases translate in the plane, so the reach balls prune nothing -- but shapes whose chains only
    move in the plane keep their world z, and pairs whose z intervals are apart (the bases 4 cm above the floor slab, the
    first arm links above the bricks ...) get no broadphase record.  Their EXACT distance (the oracle's primitive routines)
    stays positive for joint vectors far outside the limits too."""
    mk, kw = SCENES["mobile_wall_four"]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    idx = {n: i for i, n in enumerate(cs.shape_names)}
    skipped = {frozenset((idx[a], idx[c])) for a, c in cs.unreachable_pairs}
    assert any("base" in a + c and "table" in a + c for a, c in cs.unreachable_pairs)
    b = cs.blob64
    ns = cs.n_moving + cs.n_static
    offS = int(b[S.H_OFF_SHAPES])
    rad = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)[:, 3].view(np.float64)
    typed = []
    for t in range(6):
        n, off = int(b[S.H_N_PAIRS + t]), int(b[S.H_OFF_PAIRS + t])
        for i in range(n):
            pk = int(b[off + i])
            a, c = pk & 0xffff, (pk >> 16) & 0xfff
            if frozenset((a, c)) in skipped:
                typed.append((t, a, c))
    assert len(typed) == len(skipped) >= 8
    rng = np.random.default_rng(3)
    lim = sc.limits()
    qs = np.concatenate([rng.uniform(lim[0], lim[1], (400, sc.dof)), rng.uniform(-9.0, 9.0, (400, sc.dof))])
    closest = np.inf
    for q in qs:
        W = O.world_shapes(b, q, ns)
        for t, a, c in typed:
            closest = min(closest, O.pair_distance(t, W[a], W[c], rad[a] + rad[c]))
    assert closest > 4 * S.CULL_SLACK, f"a vertically separated pair comes within {closest} of touching"
