"""Second, independent oracle for the rai arithmetic: generic convex GJK + EPA on support
mappings (oracle/oracle_gjk.c -- what rai's libccd-based narrowphase does, cylinders as true convex cylinders) against
the analytic primitive routines of oracle/oracle_scene.c, which the CUDA kernels restate in fp32.

What this bounds: the conventions of the analytic oracle that nothing else pins -- a segment passing through a box is
clamped at -(r_a + r_b) instead of its true depth, cylinders in general position are capsules.  The two oracles share
forward kinematics and pair lists (pinned against the reference's own model files in tests/test_gfile.py)."""
import numpy as np
import pytest

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_gjk as G
from oracle import oracle_scene as O


def _rot(rng):
    a = rng.normal(size=(3, 3))
    qm, _ = np.linalg.qr(a)
    if np.linalg.det(qm) < 0:
        qm[:, 0] *= -1
    return qm


def _shape(core, rng, scale=0.5):
    w = np.zeros(16)
    w[:3] = rng.uniform(-1, 1, 3) * scale
    if core == 1:
        w[3:6] = w[:3] + rng.uniform(-.5, .5, 3)
    if core == 2:
        w[3:12] = _rot(rng).ravel()
        w[12:15] = rng.uniform(0.05, 0.4, 3)
    return w


@pytest.mark.parametrize("ptype,ca,cb", [(0, 0, 0), (1, 0, 1), (2, 1, 1), (3, 0, 2), (4, 1, 2), (5, 2, 2)])
def test_gjk_epa_agrees_with_the_analytic_primitives(ptype, ca, cb):
    """separated cores: distances agree to 1e-9 wherever the analytic routine returns the exact distance (box-box returns
    a SAT lower bound when clearly separated: only its sign is compared); penetrating cores: point-in-box depth and
    box-box SAT depth equal the EPA depth; a segment THROUGH a box is the documented exception (clamped at 0)."""
    rng = np.random.default_rng(ptype)
    n_pen = 0
    for _ in range(1500):
        wa, wb = _shape(ca, rng), _shape(cb, rng)
        d_an = O.pair_distance(ptype, wa, wb, 0.0)
        d_g = G.pair(ca, wa, cb, wb)
        if d_an > 1e-12:
            assert d_g > 0
            if ptype != 5:
                assert abs(d_an - d_g) < 1e-9
            else:
                assert d_g >= d_an - 1e-9          # SAT separation never exceeds the distance
        else:
            n_pen += 1
            assert d_g <= 1e-9
            if ptype in (3, 5):
                assert abs(d_an - d_g) < 1e-9      # exact interior depth / SAT minimum overlap == EPA depth
            if ptype == 4:
                assert d_an >= d_g - 1e-9          # the analytic clamp is a LOWER bound of the true penetration
    if ptype >= 3:
        assert n_pen > 20


def test_true_cylinder_support():
    """a flat disc hovering 2 cm over a large box (the 2d_handover agents over the table): distance exactly the gap"""
    box = np.zeros(16)
    box[:3] = [0, 0, 1.0]
    box[3:12] = np.eye(3).ravel()
    box[12:15] = [2, 2, 0.03]
    rng = np.random.default_rng(0)
    for _ in range(3000):
        c = np.array([rng.uniform(-1.9, 1.9), rng.uniform(-1.9, 1.9), 1.07])
        w = np.zeros(16)
        w[:3] = c
        w[3:6] = c
        w[2] -= 0.02
        w[5] += 0.02
        assert abs(G.pair(2, box, 1, w, 0.0, 0.2) - 0.02) < 1e-9


@pytest.mark.parametrize("name", list(SCENES))
def test_scene_flags_of_the_two_oracles(name):
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    lim = sc.limits()
    np.random.seed(0)
    B = 12_000 if name != "box_stacking" else 6_000
    q = np.random.uniform(lim[0], lim[1], (B, sc.dof)).astype(np.float32).astype(np.float64)
    f1, p1, m1 = O.check_configs(cs.blob64, q, nthreads=O.max_threads())
    margin = O.margin(p1, m1, cs.tol)
    # (1) same shape model as the device (cylinders in general position = capsules): flags identical on every margin-clear
    # sample although the penetration SUMS differ (segment-through-box clamp) -- the clamp never decides a flag
    f3, p3, m3 = G.check_configs(cs.blob64, q, None, nthreads=O.max_threads())
    clear = np.abs(margin) > 1e-5
    assert np.array_equal(f1[clear], f3[clear]), f"{(f1 != f3)[clear].sum()} margin-clear flags differ"
    assert (p3 >= p1 - 1e-9).all()                 # the analytic sum is a lower bound of the GJK/EPA sum
    free = p1 <= cs.tol
    assert np.max(np.abs(p1 - p3)[free]) < 1e-6    # and exact wherever it matters
    # (2) rai's cylinders as true convex cylinders: what the capsule convention costs
    cyl = G.cylinder_flags(sc, cs)
    if cyl.any():
        f2, _, _ = G.check_configs(cs.blob64, q, cyl, nthreads=O.max_threads())
        differ = f1 != f2
        print(f"{name}: {int(cyl.sum())} cylinder shapes; {differ.sum()} of {B} flags differ between capsule and true-cylinder "
              f"models, all conservative: {bool((~f1[differ]).all())}, largest |margin| {np.abs(margin[differ]).max() if differ.any() else 0:.4f}")
        assert differ.mean() < 0.005
        assert (~f1[differ]).all()                 # the capsule is a superset: it can only ADD collisions
