"""The claim behind phase A of the two-phase tiles (csrc/scene_kernels.cu table_phase): for every pair of a moving
shape X against a large static box Y, the penetration lower bound  max(0, r_X + r_Y - min_P dist(P, core Y))  over a
point's centre / five points of a segment / a box's centre never exceeds the penetration the exact routine reports.
Restated in numpy on the oracle's world data (fp64), so it runs without a GPU."""
import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O

SEG_T = (0.0, 0.25, 0.5, 0.75, 1.0)   # midpoint + u * half vector, u = -1, -0.5, 0, 0.5, 1


def _table_records(cs):
    b = cs.blob64
    out = []
    for t in (3, 4, 5):
        n = int(b[S.H_BP + (t * S.BP_SUBLISTS + 2) * 2 + 1])
        ids = int(b[S.H_BP_IDS + t * S.BP_SUBLISTS + 2])
        for i in range(n):
            pk = int(b[ids + i])
            out.append((t, pk & 0xffff, (pk >> 16) & 0xfff))
    return out


def _bound_and_exact(cs, qs):
    b = cs.blob64
    ns = cs.n_moving + cs.n_static
    offS = int(b[S.H_OFF_SHAPES])
    rad = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)[:, 3].view(np.float64)
    recs = _table_records(cs)
    worst, lb_sum, ex_sum = -np.inf, [], []
    for q in qs:
        W = O.world_shapes(b, q, ns)
        lb_q = ex_q = 0.0
        for t, a, c in recs:
            x, y = (a, c) if a < cs.n_moving else (c, a)
            rs = rad[a] + rad[c]
            ctr, R, h = W[y, :3], W[y, 3:12].reshape(3, 3), W[y, 12:15]
            pts = [W[x, :3] + u * (W[x, 3:6] - W[x, :3]) for u in SEG_T] if t == 4 else [W[x, :3]]
            dist = min(np.linalg.norm(np.maximum(np.abs(R.T @ (P - ctr)) - h, 0.0)) for P in pts)
            lb = max(0.0, rs - dist)
            ex = max(0.0, -O.pair_distance(t, W[a], W[c], rs))
            worst = max(worst, lb - ex)
            lb_q += lb
            ex_q += ex
        lb_sum.append(lb_q)
        ex_sum.append(ex_q)
    return worst, np.array(lb_sum), np.array(ex_sum), len(recs)


@pytest.mark.parametrize("name,min_retired", [("box_rearrangement", 0.35), ("box_stacking", 0.6)])
def test_bound_never_exceeds_the_exact_penetration(name, min_retired):
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    lim = sc.limits()
    qs = np.random.default_rng(7).uniform(lim[0], lim[1], (500, sc.dof))
    worst, lb, ex, n = _bound_and_exact(cs, qs)
    assert n > 0
    assert worst <= 1e-12, f"the bound exceeds the exact penetration of a pair by {worst}"
    retired = np.mean(lb > cs.tol)
    assert retired >= min_retired, retired                 # what makes the second FK pass pay
    assert retired <= np.mean(ex > cs.tol) + 1e-12


def test_bound_with_a_held_box():
    """Mode with a box in the vacuum cup: the held box is a moving box against the table (box-box records, bound from
    its centre point)."""
    from multirobot_pathplanning_benchmark_b200.scene import Scene
    mk, kw = SCENES["box_rearrangement"]
    sc = mk()
    held = sc.copy()
    held.attach("a1_ur_vacuum", "obj11", sc.home())
    cs = S.compile_blob(held, kw["tol"])
    assert any(t == 5 for t, _, _ in _table_records(cs)), "expected box-box records against the table"
    lim = held.limits()
    qs = np.random.default_rng(9).uniform(lim[0], lim[1], (300, held.dof))
    worst, lb, ex, n = _bound_and_exact(cs, qs)
    assert worst <= 1e-12
    assert np.mean(lb > cs.tol) > 0.3
