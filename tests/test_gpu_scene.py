"""GPU parity, primitive scenes (BASELINE configs 2-5): hand-written fp32 CUDA FK + narrowphase
vs the fp64 CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): the collision / no-collision flag is identical for every sample
whose signed clearance exceeds 1e-5; penetration sums agree to 2e-5."""
import numpy as np
import pytest
import torch

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O
from tests.parity import MARGIN, edge_disagreements_are_inside_margin

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def be(cuda_lib):
    from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
    b = SceneBackend(max_modes=16)
    b.scenes = {}
    for slot, (name, (mk, kw)) in enumerate(SCENES.items()):
        sc = mk()
        cs = S.compile_blob(sc, kw["tol"])
        b.set_mode(slot, cs)
        b.scenes[name] = (slot, sc, cs, kw)
    return b


def uniform_configs(sc, B, seed):
    lim = sc.limits()
    np.random.seed(seed)  # the reference's sampler: np.random.uniform in limits (planning_env.py:1697-1708)
    return np.random.uniform(lim[0], lim[1], (B, sc.dof)).astype(np.float32)


@pytest.mark.parametrize("name", list(SCENES))
def test_config_flags_and_penetration(be, name):
    slot, sc, cs, kw = be.scenes[name]
    B = 60_000 if name != "box_stacking" else 30_000
    q = uniform_configs(sc, B + 7, 0)  # ragged tail tile
    free, pen = be.check_configs(slot, torch.from_numpy(q).cuda(), return_penetration=True)
    free, pen = free.cpu().numpy(), pen.cpu().numpy()
    ofree, open_, omind = O.check_configs(cs.blob64, q.astype(np.float64), nthreads=O.max_threads())
    m = O.margin(open_, omind, cs.tol)
    clear = np.abs(m) > MARGIN
    assert clear.mean() > 0.99
    assert np.array_equal(free[clear], ofree[clear]), f"{(free != ofree)[clear].sum()} margin-clear flags differ"
    assert np.max(np.abs(pen - open_)) < 2e-5
    # the early-exit path must give the same flags as the full evaluation
    fast = be.check_configs(slot, torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(fast, free)
    assert 0.01 < free.mean() < 0.99


def test_home_and_tolerance_override(be):
    slot, sc, cs, kw = be.scenes["box_rearrangement"]
    q = torch.from_numpy(np.tile(sc.home().astype(np.float32), (33, 1))).cuda()
    assert bool(be.check_configs(slot, q).all())
    qs = torch.from_numpy(uniform_configs(sc, 20000, 3)).cuda()
    strict = be.check_configs(slot, qs, tol=0.0).cpu().numpy()
    loose = be.check_configs(slot, qs, tol=0.05).cpu().numpy()
    dflt = be.check_configs(slot, qs).cpu().numpy()
    assert strict.sum() < dflt.sum() < loose.sum()
    o0 = O.check_configs(cs.blob64, qs.cpu().numpy().astype(np.float64), tol=0.05, nthreads=O.max_threads())
    clear = np.abs(O.margin(o0[1], o0[2], 0.05)) > MARGIN
    assert np.array_equal(loose[clear], o0[0][clear])


def test_unaligned_and_tiny_batches(be):
    slot, sc, cs, kw = be.scenes["2d_handover"]
    q = uniform_configs(sc, 1000, 4)
    base = torch.from_numpy(np.r_[np.zeros(1, np.float32), q.reshape(-1)]).cuda()
    qt = base[1:].view(1000, sc.dof)  # 4-byte aligned only -> plain-load path
    got = be.check_configs(slot, qt).cpu().numpy()
    want = be.check_configs(slot, torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
    for B in (0, 1, 31, 32, 33):
        f = be.check_configs(slot, torch.from_numpy(q[:B]).cuda())
        assert f.shape == (B,) and np.array_equal(f.cpu().numpy(), want[:B])


@pytest.mark.parametrize("name", ["2d_handover", "box_rearrangement", "box_stacking", "mobile_wall_four", "abstract_like"])
def test_edges(be, name):
    slot, sc, cs, kw = be.scenes[name]
    E = {"box_rearrangement": 600, "box_stacking": 400}.get(name, 1500)
    q1 = uniform_configs(sc, E, 10)
    q2 = uniform_configs(sc, E, 11)
    q2[::2] = q1[::2] + np.random.default_rng(1).uniform(-0.15, 0.15, q1[::2].shape).astype(np.float32)
    if name == "box_stacking":   # uniform four-arm samples nearly always collide: start the local half from free samples
        pool = uniform_configs(sc, 60_000, 12)
        pool = pool[be.check_configs(slot, torch.from_numpy(pool).cuda()).cpu().numpy().astype(bool)]
        n = min(len(pool), len(q1[::2]))
        q1[:2 * n:2] = pool[:n]
        q2[:2 * n:2] = pool[:n] + np.random.default_rng(2).uniform(-0.1, 0.1, pool[:n].shape).astype(np.float32)
    res = kw["resolution"]
    free, first = be.check_edges(slot, torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda(), res)
    free, first = free.cpu().numpy(), first.cpu().numpy()
    ofree, ofirst, _ = O.check_edges(cs.blob64, q1.astype(np.float64), q2.astype(np.float64), res, nthreads=O.max_threads())
    # an edge may only disagree if one of its interpolated configurations is inside the margin
    n_diff = edge_disagreements_are_inside_margin(cs, q1, q2, res, free, first, ofree, ofirst)
    assert n_diff <= 0.005 * E + 2
    assert 0.02 < free.mean() < 0.98


@pytest.mark.parametrize("name,parent,child", [("box_rearrangement", "a1_ur_vacuum", "obj11"), ("2d_handover", "a1", "obj1"),
                                               ("mobile_wall_four", "a0_gripper", "obj_00"), ("box_stacking", "a2_ur_gripper_center", "obj00")])
def test_edges_in_held_object_modes(cuda_lib, name, parent, child):
    """A8 in a mode whose kinematic tree carries an object on a robot link (A7): uniform edges and planner-like local
    edges from free configurations, flags and first colliding positions against the oracle on the same compiled tree."""
    from multirobot_pathplanning_benchmark_b200.env import SceneModel
    mk, kw = SCENES[name]
    sc = mk()
    model = SceneModel(sc, kw["tol"], kw["resolution"])
    rng = np.random.default_rng(8)
    lim = sc.limits()
    base_slot = model.slot_for(())
    cand = rng.uniform(lim[0], lim[1], (8192, sc.dof)).astype(np.float32)
    ok = model.check_configs(base_slot, cand).cpu().numpy()
    slot = model.slot_for(("held",), [(parent, child, cand[np.argmax(ok)].astype(np.float64))])
    cs = model.compiled(slot)
    be = model.device.be
    pool = uniform_configs(sc, 80_000, 13)
    pool = pool[be.check_configs(slot, torch.from_numpy(pool).cuda()).cpu().numpy().astype(bool)]
    assert len(pool) >= 50
    n_loc = min(len(pool), 400)
    E_uni = 200 if name in ("box_stacking", "box_rearrangement") else 600
    q1 = np.vstack([pool[:n_loc], uniform_configs(sc, E_uni, 14)])
    q2 = np.vstack([pool[:n_loc] + rng.uniform(-0.12, 0.12, (n_loc, sc.dof)).astype(np.float32), uniform_configs(sc, E_uni, 15)])
    res = kw["resolution"]
    free, first = be.check_edges(slot, torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda(), res)
    free, first = free.cpu().numpy(), first.cpu().numpy()
    ofree, ofirst, _ = O.check_edges(cs.blob64, q1.astype(np.float64), q2.astype(np.float64), res, nthreads=O.max_threads())
    n_diff = edge_disagreements_are_inside_margin(cs, q1, q2, res, free, first, ofree, ofirst)
    assert n_diff <= 0.01 * len(q1) + 2
    assert free[:n_loc].any() and not free.all()
    # the same edges in the start mode give different answers somewhere: the held object matters
    bfree, _ = be.check_edges(base_slot, torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda(), res)
    assert (bfree.cpu().numpy() != free).any()


def test_edge_windows_and_explicit_N(be):
    slot, sc, cs, kw = be.scenes["2d_handover"]
    q1, q2 = uniform_configs(sc, 400, 20), uniform_configs(sc, 400, 21)
    t1, t2 = torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda()
    for (ns, nm, inc) in ((0, 2, False), (2, 12, False), (0, None, True), (10, 20, True)):
        f, p = be.check_edges(slot, t1, t2, 0.01, n_start=ns, n_max=nm, include_endpoints=inc)
        of, op, _ = O.check_edges(cs.blob64, q1.astype(np.float64), q2.astype(np.float64), 0.01, n_start=ns,
                                  n_max=-1 if nm is None else nm, include_endpoints=inc)
        n_diff = edge_disagreements_are_inside_margin(cs, q1, q2, 0.01, f.cpu().numpy(), p.cpu().numpy(), of, op, n_start=ns,
                                                      n_max=nm, include_endpoints=inc)
        assert n_diff <= 4
    Nh = np.full(400, 25, np.int32)
    N = torch.from_numpy(Nh).cuda()
    f, p = be.check_edges(slot, t1, t2, 0.01, N=N)
    of, op, chk = O.check_edges(cs.blob64, q1.astype(np.float64), q2.astype(np.float64), 0.01, Ns=Nh)
    assert edge_disagreements_are_inside_margin(cs, q1, q2, 0.01, f.cpu().numpy(), p.cpu().numpy(), of, op, Ns=Nh) <= 4
    assert chk.max() <= 23


@pytest.mark.parametrize("name", list(SCENES))
def test_for_robot_rule(be, name):
    """A6: is_collision_free_for_robot (rai_base_env.py:515-615) on every scene: the first robot is the robot asked about,
    every other robot counts as `other` (contacts that involve neither the robot nor only other robots are ignored)."""
    slot, sc, cs, kw = be.scenes[name]
    robots = sorted({sc.robot_of_shape(n) for n in cs.shape_names} - {None, ""})
    if len(robots) < 2:
        pytest.skip("single-robot scene")
    me = robots[0]
    rel = np.array([1 if sc.robot_of_shape(n) == me else 0 for n in cs.shape_names], np.uint8)
    oth = np.array([1 if sc.robot_of_shape(n) not in (me, None, "") else 0 for n in cs.shape_names], np.uint8)
    q = uniform_configs(sc, 40000, 30)
    got = be.check_configs_for_robot(slot, torch.from_numpy(q).cuda(), rel, oth).cpu().numpy()
    ofree, open_, omind = O.check_configs(cs.blob64, q.astype(np.float64), rel=rel, oth=oth, nthreads=O.max_threads())
    clear = (np.abs(O.margin(open_, omind, cs.tol)) > MARGIN) & (np.abs(omind) > MARGIN)
    assert np.array_equal(got[clear], ofree[clear])
    plain = be.check_configs(slot, torch.from_numpy(q).cuda()).cpu().numpy()
    assert (got >= plain).all() and got.sum() > plain.sum()
    # the single-query entry point (eight-warp kernel) gives the same flags
    few = be.query_configs_host(slot, q[:7], None, rel, oth)
    assert np.array_equal(np.asarray(few).astype(bool), got[:7].astype(bool))


@pytest.mark.parametrize("name,parent,child", [("box_rearrangement", "a1_ur_vacuum", "obj11"), ("2d_handover", "a1", "obj1"),
                                               ("mobile_wall_four", "a0_gripper", "obj_00"), ("box_stacking", "a2_ur_gripper_center", "obj00")])
def test_held_object_modes(cuda_lib, name, parent, child):
    """A7: kinematic tree of a mode in which a robot holds an object (attach + contact -1,
    rai_base_env.py:776-805): the object's shapes ride on the robot link and collide with everything
    except its parent link."""
    from multirobot_pathplanning_benchmark_b200.env import SceneModel
    mk, kw = SCENES[name]
    sc = mk()
    model = SceneModel(sc, kw["tol"], kw["resolution"])
    rng = np.random.default_rng(3)
    lim = sc.limits()
    # attach at a configuration where the scene is collision free
    base_slot = model.slot_for(())
    cand = rng.uniform(lim[0], lim[1], (4096, sc.dof)).astype(np.float32)
    ok = model.check_configs(base_slot, cand).cpu().numpy()
    q_attach = cand[np.argmax(ok)].astype(np.float64)
    slot = model.slot_for(("held",), [(parent, child, q_attach)])
    cs = model.compiled(slot)
    assert cs.n_moving == model.compiled(base_slot).n_moving + 1
    q = uniform_configs(sc, 40_000, 5)
    free, pen = model.device.be.check_configs(slot, torch.from_numpy(q).cuda(), return_penetration=True)
    ofree, open_, omind = O.check_configs(cs.blob64, q.astype(np.float64), nthreads=O.max_threads())
    clear = np.abs(O.margin(open_, omind, cs.tol)) > MARGIN
    assert np.array_equal(free.cpu().numpy()[clear], ofree[clear])
    assert np.max(np.abs(pen.cpu().numpy() - open_)) < 2e-5
    base_free = model.check_configs(base_slot, torch.from_numpy(q).cuda()).cpu().numpy()
    assert (free.cpu().numpy() != base_free).any()  # the mode really changes the answer
    # two-phase tiles in this mode: the held box is a moving box against the table (box-box records in phase A, their
    # radii read through the pair ids); same flags as the single-pass kernel and as the full evaluation
    be = model.device.be
    qd = torch.from_numpy(q).cuda()
    try:
        be.set_two_phase(slot, "never")
        single = be.check_configs(slot, qd).cpu().numpy()
        be.set_two_phase(slot, "always")
        two = be.check_configs(slot, qd).cpu().numpy()
    finally:
        be.set_two_phase(slot, "auto")
    assert np.array_equal(single, two) and np.array_equal(two, free.cpu().numpy())


def test_host_buffer_queries_equal_device_buffer_calls(be):
    """mrb200_query_configs_host / mrb200_query_edges_host (the planners' single-query seam): same answers as the
    device-buffer entry points, for one query and for small batches, plain rule and per-robot rule, edge windows"""
    slot, sc, cs, kw = be.scenes["box_rearrangement"]
    q = uniform_configs(sc, 777, 31)
    want = be.check_configs(slot, torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(be.query_configs_host(slot, q), want)
    for i in (0, 1, 500):
        assert be.query_configs_host(slot, q[i:i + 1])[0] == want[i]
    assert np.array_equal(be.query_configs_host(slot, q, tol=0.0),
                          be.check_configs(slot, torch.from_numpy(q).cuda(), tol=0.0).cpu().numpy())
    rel = np.array(["a1_" in n for n in cs.shape_names], np.uint8)
    oth = np.array(["a2_" in n for n in cs.shape_names], np.uint8)
    assert np.array_equal(be.query_configs_host(slot, q, relevant=rel, other=oth),
                          be.check_configs_for_robot(slot, torch.from_numpy(q).cuda(), rel, oth).cpu().numpy())
    q2 = q + np.random.default_rng(2).uniform(-0.3, 0.3, q.shape).astype(np.float32)
    t1, t2 = torch.from_numpy(q).cuda(), torch.from_numpy(q2).cuda()
    for kwargs in ({}, {"n_start": 2, "n_max": 9}, {"include_endpoints": True}):
        f, p = be.check_edges(slot, t1, t2, 0.01, **kwargs)
        fh, ph = be.query_edges_host(slot, q, q2, 0.01, **kwargs)
        assert np.array_equal(fh, f.cpu().numpy()) and np.array_equal(ph, p.cpu().numpy())
    N = np.full(len(q), 17, np.int32)
    f, p = be.check_edges(slot, t1, t2, 0.01, N=torch.from_numpy(N).cuda())
    fh, ph = be.query_edges_host(slot, q, q2, 0.01, N=N)
    assert np.array_equal(fh, f.cpu().numpy()) and np.array_equal(ph, p.cpu().numpy())
    assert be.query_configs_host(slot, q[:0]).shape == (0,)
    with pytest.raises(Exception):
        be.query_configs_host(slot, q[:, :5])


@pytest.mark.parametrize("pinned", [True, False])
def test_chunked_host_batches_equal_device_flags(be, pinned):
    """mrb200_check_configs_host (whole sample batches from host memory, chunks pipelined on the library's side streams):
    the same flags as the device-buffer call, for pinned and pageable buffers, ragged last chunks, and a stream of calls"""
    from multirobot_pathplanning_benchmark_b200.backend import check_configs_host
    slot, sc, cs, kw = be.scenes["box_rearrangement"]
    for B, chunk in ((100_003, 1 << 14), (40_000, 1 << 17), (33, 32), (5 * 4096, 4096)):
        q = torch.from_numpy(uniform_configs(sc, B, 77))
        out = torch.full((B,), 7, dtype=torch.uint8)
        if pinned:
            q, out = q.pin_memory(), out.pin_memory()
        want = be.check_configs(slot, q.cuda()).cpu()
        for _ in range(2):
            check_configs_host(be, slot, q, out, chunk=chunk)
            assert torch.equal(out, want)
            out.fill_(7)
    with pytest.raises(ValueError):
        check_configs_host(be, slot, torch.zeros(4, sc.dof + 1), torch.zeros(4, dtype=torch.uint8))


def test_edge_kernel_corner_cases(be):
    """multi-edge tiles: degenerate edges (N = 2: no interior point), empty windows, a single edge, mixed explicit N,
    windows beyond N, batches smaller than one tile and far larger than the grid -- all against the oracle"""
    slot, sc, cs, kw = be.scenes["2d_handover"]
    rng = np.random.default_rng(12)
    lim = sc.limits()

    def both(q1, q2, **kwargs):
        f, p = be.check_edges(slot, torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda(), 0.01,
                              **{k: (torch.from_numpy(v).cuda() if k == "N" else v) for k, v in kwargs.items()})
        okw = dict(kwargs)
        if "N" in okw:
            okw["Ns"] = okw.pop("N")
        if okw.get("n_max") is None:
            okw.pop("n_max", None)
        of, op, _ = O.check_edges(cs.blob64, q1.astype(np.float64), q2.astype(np.float64), 0.01, **okw)
        return f.cpu().numpy(), p.cpu().numpy(), of, op

    free_q = uniform_configs(sc, 20000, 40)
    free_q = free_q[be.check_configs(slot, torch.from_numpy(free_q).cuda()).cpu().numpy().astype(bool)]
    # (1) identical endpoints and sub-resolution moves: N = 2, nothing to check -> free, first = -1
    q1 = free_q[:300]
    q2 = q1 + rng.uniform(-0.004, 0.004, q1.shape).astype(np.float32)
    f, p, of, op = both(q1, q2)
    assert f.all() and (p == -1).all() and np.array_equal(f, of) and np.array_equal(p, op)
    # ... but with include_endpoints the two endpoints are checked
    f, p, of, op = both(q1, q2, include_endpoints=True)
    assert np.array_equal(f, of) and np.array_equal(p, op)
    # (2) a single edge, and a batch of 33 edges (one more than a tile)
    for n in (1, 33):
        a = free_q[:n]
        b = free_q[1000:1000 + n]
        f, p, of, op = both(a, b)
        edge_disagreements_are_inside_margin(cs, a, b, 0.01, f, p, of, op)
    # (3) explicit N of wildly different sizes in one batch, including the minimum
    n = 2000
    a, b = free_q[:n], free_q[2000:2000 + n]
    N = rng.choice([2, 3, 5, 31, 32, 33, 64, 700], n).astype(np.int32)
    f, p, of, op = both(a, b, N=N)
    assert edge_disagreements_are_inside_margin(cs, a, b, 0.01, f, p, of, op, Ns=N) <= 10
    # (4) windows: empty (n_start = n_max), beyond the end, and a late start
    for ns, nm in ((5, 5), (0, 100000), (40, None), (3, 4)):
        f, p, of, op = both(a, b, N=N, n_start=ns, n_max=nm)
        assert edge_disagreements_are_inside_margin(cs, a, b, 0.01, f, p, of, op, Ns=N, n_start=ns, n_max=nm) <= 10, (ns, nm)
        if ns == nm:
            assert f.all() and (p == -1).all()
    # (5) far more edges than resident CTAs x slots, all short: every edge gets exactly one answer
    big = 300_000
    a = free_q[rng.integers(0, len(free_q), big)]
    b = a + rng.uniform(-0.05, 0.05, a.shape).astype(np.float32)
    fl = torch.full((big,), 9, dtype=torch.uint8, device="cuda")
    f, p = be.check_edges(slot, torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), 0.01)
    assert f.shape == (big,) and set(np.unique(f.cpu().numpy())) <= {0, 1}
    sub = rng.integers(0, big, 3000)
    of, op, _ = O.check_edges(cs.blob64, a[sub].astype(np.float64), b[sub].astype(np.float64), 0.01)
    assert edge_disagreements_are_inside_margin(cs, a[sub], b[sub], 0.01, f.cpu().numpy()[sub], p.cpu().numpy()[sub], of, op) <= 15


def test_informed_batch_sampler_prunes_on_the_device(cuda_lib):
    """ellipse prune (two batch_config_cost evaluations on the device) before the collision kernel:
    every returned configuration is inside the ellipse, collision free per the oracle, and the prune really happens"""
    from multirobot_pathplanning_benchmark_b200.env import SceneModel
    from oracle import oracle_abstract as OA
    mk, kw = SCENES["box_rearrangement"]
    sc = mk()
    model = SceneModel(sc, kw["tol"], kw["resolution"])
    slot = model.slot_for(())
    home = sc.home()
    goal = home.copy()
    goal[:6] += np.array([0.8, 0.5, -0.6, 0.4, 0.3, -0.5])
    sl = np.array([list(sc.robot_slices()[r]) for r in sc.robots])
    direct = OA.batch_config_cost((goal - home)[None], sl, "euclidean", "max")[0]
    bound = 5.0 * direct
    q, drawn, pruned = model.sample_informed(slot, 500, np.stack([home, goal]), bound, "euclidean", "max", np.random.RandomState(3))
    assert len(q) == 500 and pruned > 0.5 * drawn           # most of the box lies outside the ellipse
    assert drawn < 3e8
    c = OA.batch_config_cost(q - home[None], sl, "euclidean", "max") + OA.batch_config_cost(q - goal[None], sl, "euclidean", "max")
    assert (c <= bound * (1 + 1e-12)).all()
    cs = model.compiled(slot)
    ofree, open_, omind = O.check_configs(cs.blob64, q)
    clear = np.abs(O.margin(open_, omind, cs.tol)) > MARGIN
    assert ofree[clear].all()
    # an ellipse that cannot contain anything returns nothing (and says how much it drew)
    q0, drawn0, pruned0 = model.sample_informed(slot, 10, np.stack([home, goal]), 0.5 * direct, "euclidean", "max",
                                                np.random.RandomState(4), max_rounds=2)
    assert len(q0) == 0 and pruned0 == drawn0


# ---- two-phase tiles (table / floor bound first, pooled survivors evaluated exactly) ----

@pytest.mark.parametrize("name", ["box_rearrangement", "box_stacking", "mobile_wall_four"])
@pytest.mark.parametrize("B", [4096, 4096 + 13, 50_001])
def test_two_phase_tiles_give_the_single_pass_flags(be, name, B):
    """Forced two-phase tiles, forced single pass and the full evaluation agree flag for flag (not only on the
    margin-clear samples): survivors run through the same arithmetic, retired configurations are provably colliding."""
    slot, sc, cs, kw = be.scenes[name]
    q = torch.from_numpy(uniform_configs(sc, B, 31)).cuda()
    try:
        be.set_two_phase(slot, "never")
        single = be.check_configs(slot, q).cpu().numpy()
        be.set_two_phase(slot, "always")
        two = be.check_configs(slot, q).cpu().numpy()
    finally:
        be.set_two_phase(slot, "auto")
    assert np.array_equal(single, two), f"{(single != two).sum()} flags differ between the two kernels"
    full = be.check_configs(slot, q, full_eval=True).cpu().numpy()
    assert np.array_equal(full, two)
    ofree, open_, omind = O.check_configs(cs.blob64, q.cpu().numpy().astype(np.float64), nthreads=O.max_threads())
    clear = np.abs(O.margin(open_, omind, cs.tol)) > MARGIN
    assert np.array_equal(two[clear], ofree[clear])


def test_two_phase_on_inputs_the_bound_never_decides(be):
    """All configurations at the (free) home pose: phase A retires nothing, every CTA falls back to single-pass
    tiles, the pool is flushed, every flag is written exactly once."""
    slot, sc, cs, kw = be.scenes["box_stacking"]
    q = torch.from_numpy(np.tile(sc.home().astype(np.float32), (20_000 + 5, 1))).cuda()
    try:
        be.set_two_phase(slot, "always")
        out = torch.full((q.shape[0],), 7, dtype=torch.uint8, device="cuda")
        be.check_configs(slot, q, out=out)
        assert bool((out == 1).all())
        # and on inputs it always decides: every arm folded into the table
        lim = sc.limits()
        deep = uniform_configs(sc, 10_000, 5)
        first = sc.robot_slices()[sc.robots[0]][0]
        deep[:, first + 1] = 1.2   # robot 0: shoulder lift pointing down through the table
        free = be.check_configs(slot, torch.from_numpy(deep).cuda()).cpu().numpy()
        be.set_two_phase(slot, "never")
        assert np.array_equal(free, be.check_configs(slot, torch.from_numpy(deep).cuda()).cpu().numpy())
    finally:
        be.set_two_phase(slot, "auto")


def test_two_phase_policy_settles_from_measurements(be):
    """auto: the first large batches measure how much the bound decides; the slot settles on two-phase tiles for
    the four-arm scene (most uniform samples fold an arm into the table) and on single pass for the dual-arm one."""
    for name, want in (("box_stacking", "two_phase"), ("box_rearrangement", "single_pass")):
        slot, sc, cs, kw = be.scenes[name]
        be.set_two_phase(slot, "auto")
        assert be.two_phase_info(slot)["state"] == "measuring"
        q = torch.from_numpy(uniform_configs(sc, 30_000, 8)).cuda()
        a = be.check_configs(slot, q).cpu().numpy()     # measuring launch; .cpu() synchronises
        b = be.check_configs(slot, q).cpu().numpy()     # reads the counters, settles
        info = be.two_phase_info(slot)
        assert info["state"] == want, info
        assert info["seen"] >= 30_000 and 0 <= info["decided_by_bound"] <= info["seen"]
        assert np.array_equal(a, b)
        small = be.check_configs(slot, q[:100]).cpu().numpy()   # small batches never pool
        assert np.array_equal(small, a[:100])
