"""Pins the fp64 scene oracle (oracle/oracle_scene.c): geometric primitives against independent
numerical minimisation (scipy), forward kinematics against the host frame-tree model, the edge
loop against the reference's golden edge vectors, and the scene statistics that guard against an
always-colliding pair rule."""
import numpy as np
import pytest
from scipy.optimize import minimize, minimize_scalar

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_abstract as OA
from oracle import oracle_scene as O

PT = S.PAIR_TYPE


def rand_rot(rng):
    a = rng.normal(size=3)
    return S.axis_angle_mat(a, rng.uniform(0, np.pi))


def box_arr(c, R, h):
    return np.concatenate([c, R.reshape(-1), h, [0.0]])


def pt_box_dist(p, c, R, h):
    l = R.T @ (p - c)
    return np.linalg.norm(l - np.clip(l, -h, h))


def test_seg_seg_against_numerical_minimum():
    rng = np.random.default_rng(0)
    for _ in range(300):
        s1, s2 = rng.uniform(-1, 1, 6), rng.uniform(-1, 1, 6)
        if rng.random() < 0.15:  # parallel / degenerate cases
            s2[3:] = s2[:3] + (s1[3:] - s1[:3]) * rng.uniform(-2, 2)
        if rng.random() < 0.05:
            s2[3:] = s2[:3]
        f = lambda x: np.linalg.norm((s1[:3] + x[0] * (s1[3:] - s1[:3])) - (s2[:3] + x[1] * (s2[3:] - s2[:3])))
        best = min(minimize(f, x0, bounds=[(0, 1), (0, 1)], method="L-BFGS-B", options=dict(ftol=1e-15, gtol=1e-12)).fun
                   for x0 in ([0.5, 0.5], [0, 0], [1, 1], [0, 1], [1, 0]))
        got = O.pair_distance(PT[(1, 1)], np.r_[s1, np.zeros(10)], np.r_[s2, np.zeros(10)], 0.1)
        assert abs(got - (best - 0.1)) < 1e-6


def test_point_seg_and_point_point():
    rng = np.random.default_rng(1)
    for _ in range(200):
        p, s = rng.uniform(-1, 1, 3), rng.uniform(-1, 1, 6)
        t = minimize_scalar(lambda t: np.linalg.norm(p - (s[:3] + t * (s[3:] - s[:3]))), bounds=(0, 1), method="bounded",
                            options=dict(xatol=1e-12)).fun
        got = O.pair_distance(PT[(0, 1)], np.r_[p, np.zeros(13)], np.r_[s, np.zeros(10)], 0.05)
        assert abs(got - (t - 0.05)) < 1e-7
        p2 = rng.uniform(-1, 1, 3)
        got = O.pair_distance(PT[(0, 0)], np.r_[p, np.zeros(13)], np.r_[p2, np.zeros(13)], 0.3)
        assert abs(got - (np.linalg.norm(p - p2) - 0.3)) < 1e-14


def test_point_box_inside_and_outside():
    rng = np.random.default_rng(2)
    for _ in range(300):
        c, R, h = rng.uniform(-1, 1, 3), rand_rot(rng), rng.uniform(0.05, 0.6, 3)
        p = c + R @ (rng.uniform(-1.5, 1.5, 3) * h)
        got = O.pair_distance(PT[(0, 2)], np.r_[p, np.zeros(13)], box_arr(c, R, h), 0.03)
        l = R.T @ (p - c)
        if np.all(np.abs(l) <= h):
            assert abs(got - (-np.min(h - np.abs(l)) - 0.03)) < 1e-12
        else:
            assert abs(got - (pt_box_dist(p, c, R, h) - 0.03)) < 1e-12


def test_seg_box_against_numerical_minimum():
    rng = np.random.default_rng(3)
    n_touch = 0
    for _ in range(400):
        c, R, h = rng.uniform(-.5, .5, 3), rand_rot(rng), rng.uniform(0.03, 0.5, 3)
        a, b = rng.uniform(-1.2, 1.2, 3), rng.uniform(-1.2, 1.2, 3)
        if rng.random() < 0.2:  # axis-parallel segments hit the flat parts of f(t)
            b = a + R[:, rng.integers(3)] * rng.uniform(-1, 1)
        f = lambda t: pt_box_dist(a + t * (b - a), c, R, h)
        ts = np.linspace(0, 1, 2001)
        vals = np.array([f(t) for t in ts])
        k = int(np.argmin(vals))
        lo, hi = ts[max(k - 1, 0)], ts[min(k + 1, len(ts) - 1)]
        best = min(vals[k], minimize_scalar(f, bounds=(lo, hi), method="bounded", options=dict(xatol=1e-13)).fun)
        got = O.pair_distance(PT[(1, 2)], np.r_[a, b, np.zeros(10)], box_arr(c, R, h), 0.07)
        if best < 1e-9:
            n_touch += 1
            assert got == pytest.approx(-0.07, abs=1e-7)  # cores intersect: clamped at -(ra+rb)
        else:
            assert abs(got - (best - 0.07)) < 2e-7
    assert n_touch > 10


def test_box_box_sat_sign_and_exact_distance():
    rng = np.random.default_rng(4)
    n_sep = n_pen = 0
    for _ in range(250):
        cA, RA, hA = rng.uniform(-.4, .4, 3), rand_rot(rng), rng.uniform(0.05, 0.4, 3)
        cB, RB, hB = rng.uniform(-.4, .4, 3), rand_rot(rng), rng.uniform(0.05, 0.4, 3)
        A, B = box_arr(cA, RA, hA), box_arr(cB, RB, hB)
        # independent exact distance: minimise over points of A the distance to B (convex)
        f = lambda x: pt_box_dist(cA + RA @ x, cB, RB, hB)
        starts = [np.zeros(3)] + [np.array(s) * hA for s in np.ndindex(2, 2, 2)] + [-np.array(s) * hA for s in np.ndindex(2, 2, 2)]
        best = min(minimize(f, x0, bounds=[(-hA[i], hA[i]) for i in range(3)], method="L-BFGS-B",
                            options=dict(ftol=1e-16, gtol=1e-12)).fun for x0 in starts)
        sat = O.lib().orc_box_box_sat(O._p(A, O.C.c_double), O._p(B, O.C.c_double))
        exact = O.lib().orc_box_box_exact_dist(O._p(A, O.C.c_double), O._p(B, O.C.c_double))
        if best > 1e-6:
            n_sep += 1
            assert sat > 0 and sat <= best + 1e-9        # SAT gap is a lower bound of the distance
            assert abs(exact - best) < 5e-6               # edge-vs-box enumeration is the exact distance
            # rounded boxes: d = dist - r when the SAT gap is below r
            got = O.pair_distance(PT[(2, 2)], A, B, best + 0.01)
            assert abs(got - (-0.01)) < 5e-6
        elif best < 1e-9:
            n_pen += 1
            assert sat <= 1e-9
            # minimum-overlap depth really separates the boxes when applied (up to skipped axes)
            assert O.pair_distance(PT[(2, 2)], A, B, 0.0) == pytest.approx(sat, abs=1e-12)
    assert n_sep > 30 and n_pen > 30


def test_prism_pairs():
    rng = np.random.default_rng(5)
    for _ in range(300):
        # upright cylinder vs z-aligned box, sampled densely in 2-D
        c = np.r_[rng.uniform(-.5, .5, 2), rng.uniform(-.05, .05)]
        R = S.axis_angle_mat([0, 0, 1], rng.uniform(-np.pi, np.pi))
        h = np.r_[rng.uniform(0.05, 0.4, 2), 0.03]
        cyl = np.r_[rng.uniform(-.8, .8, 2), rng.uniform(-.05, .05)]
        r, hc = rng.uniform(0.05, 0.2), 0.02
        got = O.pair_distance(PT[(2, 3)], box_arr(c, R, h), np.r_[cyl, r, hc, np.zeros(11)], 0.0)
        l = R.T @ (cyl - c)
        ex, ey = abs(l[0]) - h[0], abs(l[1]) - h[1]
        s2 = (max(ex, ey) if (ex <= 0 and ey <= 0) else np.hypot(max(ex, 0), max(ey, 0))) - r
        sz = abs(l[2]) - h[2] - hc
        want = np.hypot(s2, sz) if (s2 > 0 and sz > 0) else (s2 if s2 > 0 else (sz if sz > 0 else max(s2, sz)))
        assert abs(got - want) < 1e-12


@pytest.mark.parametrize("name", list(SCENES))
def test_fk_matches_host_frame_tree(name):
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    rng = np.random.default_rng(6)
    lim = sc.limits()
    n_shapes = cs.n_moving + cs.n_static
    for _ in range(5):
        q = rng.uniform(lim[0], lim[1])
        W = O.world_shapes(cs.blob64, q, n_shapes)
        X = sc.fk(q)
        for i, nm in enumerate(cs.shape_names):
            f = sc.frames[nm]
            core, rad, extra = f.shape.core(planar_z=(f.shape.kind == "cylinder" and name == "2d_handover"))
            T = X[nm]
            if core == S.CORE_SEG:
                hl = extra["half_len"]
                assert np.allclose(W[i, :6], np.r_[T.apply([0, 0, -hl]), T.apply([0, 0, hl])], atol=1e-12)
            elif core == S.CORE_BOX:
                assert np.allclose(W[i, :3], T.t, atol=1e-12) and np.allclose(W[i, 3:12].reshape(3, 3), T.R, atol=1e-12)
                assert np.allclose(W[i, 12:15], extra["half"])
            else:
                assert np.allclose(W[i, :3], T.t, atol=1e-12)


def test_scene_statistics():
    """Guards the collidable-pair rule: no named scene may be always in collision, the home
    pose must be free, static geometry must not interpenetrate."""
    expect = {"2d_handover": (0.2, 0.6), "box_rearrangement": (0.2, 0.7), "box_stacking": (0.02, 0.3),
              "mobile_wall_four": (0.3, 0.8), "abstract_like": (0.7, 1.0)}
    for name, (mk, kw) in SCENES.items():
        sc = mk()
        cs = S.compile_blob(sc, kw["tol"])
        rng = np.random.default_rng(7)
        lim = sc.limits()
        q = rng.uniform(lim[0], lim[1], (4000, sc.dof))
        free, pen, mind = O.check_configs(cs.blob64, q)
        lo, hi = expect[name]
        assert lo < free.mean() < hi, (name, free.mean())
        assert O.check_configs(cs.blob64, sc.home()[None])[0][0], name
        assert O.static_penetration(cs.blob64) == 0.0


def test_binary_indices_c_matches_reference_order():
    for N in (1, 2, 3, 4, 5, 17, 64, 333, 657):
        assert tuple(O.binary_indices(N)) == OA.binary_search_indices(N)


def test_edge_loop_against_reference_golden(golden):
    """The scene oracle's edge loop on the 3-D lift of abstract.test must reproduce the
    reference's abstract.test edge flags and call counts (same discretisation, equivalent
    geometry: spheres of radius .1 in the z=0 plane, sphere obstacle, box obstacle)."""
    mk, kw = SCENES["abstract_like"]
    sc = mk()
    cs = S.compile_blob(sc, 0.0)
    q1, q2 = golden["edge_q1"], golden["edge_q2"]
    lift = lambda q: np.c_[q[:, 0:2], np.zeros(len(q)), q[:, 2:4], np.zeros(len(q))]
    free, first, checks = O.check_edges(cs.blob64, lift(q1), lift(q2), 0.01)
    # the rectangle test is `<=` and the sphere tests `<` in the reference, here both are
    # "penetration > 0": identical except on exact contact, which random edges never hit
    assert np.array_equal(free, golden["edge_free"])
    assert np.array_equal(checks, golden["edge_calls"])
