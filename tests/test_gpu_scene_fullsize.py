"""GPU parity at BASELINE.json's full sizes through size-independent properties (the oracle cannot finish
millions of configurations in seconds): order independence, chunk invariance, edge results against the
configuration kernel on the reference's own interpolation points, and an oracle check of a random sample.

Semantics restated: P/problems/rai_base_env.py:442-477 (flag), :618-676 + P/problems/planning_env.py:34-51
(edge discretisation and binary order), P/ = src/multi_robot_multi_goal_planning/ in the reference."""
import numpy as np
import pytest
import torch

from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O

pytestmark = pytest.mark.gpu
FULL = {"2d_handover": 1_048_576, "box_rearrangement": 4_194_304, "box_stacking": 1_048_576, "mobile_wall_four": 4_194_304}


@pytest.fixture(scope="module")
def be(cuda_lib):
    from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
    b = SceneBackend(max_modes=8)
    b.scenes = {}
    for slot, name in enumerate(FULL):
        mk, kw = SCENES[name]
        sc = mk()
        cs = S.compile_blob(sc, kw["tol"])
        b.set_mode(slot, cs)
        b.scenes[name] = (slot, sc, cs, kw)
    return b


def uniform_device(sc, B, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    lim = torch.from_numpy(sc.limits().astype(np.float32)).cuda()
    return lim[0] + (lim[1] - lim[0]) * torch.rand((B, sc.dof), generator=g, device="cuda")


@pytest.mark.parametrize("name", list(FULL))
def test_full_batch_is_order_and_chunk_invariant_and_matches_oracle_sample(be, name):
    slot, sc, cs, kw = be.scenes[name]
    B = FULL[name]
    q = uniform_device(sc, B, 1)
    flags = be.check_configs(slot, q)
    assert flags.shape == (B,) and 0.01 < flags.float().mean().item() < 0.99
    # (1) a permuted batch gives the permuted flags (no dependence on tile / lane position)
    perm = torch.randperm(B, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    assert torch.equal(be.check_configs(slot, q[perm].contiguous()), flags[perm])
    # (2) ragged chunks (unaligned starts, tail tiles) give the same flags as one launch
    cuts = [0, 1, 33, 4097, B // 3 + 5, B // 2 - 1, B - 31, B]
    parts = [be.check_configs(slot, q[a:b].contiguous()) for a, b in zip(cuts[:-1], cuts[1:])]
    assert torch.equal(torch.cat(parts), flags)
    # (3) full evaluation (penetration sums, no early exit) agrees with the early-exit flags
    f2, pen = be.check_configs(slot, q, return_penetration=True)
    assert torch.equal(f2, flags)
    assert torch.equal(flags.bool(), ~(pen > cs.tol))
    # (4) a random sample of the full batch against the fp64 oracle
    idx = torch.randint(0, B, (20_000,), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    qs = q[idx].cpu().numpy()
    ofree, open_, omind = O.check_configs(cs.blob64, qs.astype(np.float64), nthreads=O.max_threads())
    clear = np.abs(O.margin(open_, omind, cs.tol)) > 1e-5
    assert np.array_equal(flags[idx].cpu().numpy()[clear], ofree[clear])
    assert np.max(np.abs(pen[idx].cpu().numpy() - open_)) < 2e-5


@pytest.mark.parametrize("name,E", [("2d_handover", 100_000), ("box_rearrangement", 40_000), ("box_stacking", 20_000),
                                    ("mobile_wall_four", 40_000)])
def test_edges_equal_config_kernel_on_the_reference_interpolation_points(be, name, E):
    """edge free <=> every interior interpolation point free, first colliding position = first hit in the
    reference's binary order; both sides on the device, inputs bit-identical, so the agreement is exact."""
    slot, sc, cs, kw = be.scenes[name]
    res = kw["resolution"]
    q1 = uniform_device(sc, E, 5)
    if name == "box_stacking":   # uniform four-arm samples nearly always collide: start half of the edges from free samples
        pool = uniform_device(sc, 40 * E, 4)
        pool = pool[be.check_configs(slot, pool).bool()]
        n = min(pool.shape[0], E // 2)
        q1[:n] = pool[:n]
    step = (torch.rand((E, sc.dof), device="cuda", generator=torch.Generator(device="cuda").manual_seed(6)) - 0.5)
    scale = torch.rand((E, 1), device="cuda", generator=torch.Generator(device="cuda").manual_seed(7)) * 1.2
    q2 = q1 + step * scale  # planner-like edges: 0 .. 0.6 rad per joint, N = 2 .. ~60 (120 on the 0.02 scene)
    free, first = be.check_edges(slot, q1, q2, res)
    a, b = q1.double(), q2.double()
    N = torch.clamp((torch.max(torch.abs(a - b), dim=1).values / res).to(torch.int64) + 1, min=2)
    d = (b - a) / (N - 1).double()[:, None]
    Nmax = int(N.max().item())
    # all interior points of all edges, flattened
    i = torch.arange(Nmax, device="cuda")[None, :].expand(E, Nmax)
    interior = (i >= 1) & (i < (N - 1)[:, None])
    e_id = torch.arange(E, device="cuda")[:, None].expand(E, Nmax)[interior]
    ii = i[interior]
    pts = (a[e_id] + d[e_id] * ii.double()[:, None]).float()
    pf = be.check_configs(slot, pts.contiguous()).bool()
    hit = torch.zeros((E, Nmax), dtype=torch.bool, device="cuda")
    hit[e_id, ii] = ~pf
    assert torch.equal(free.bool(), ~hit.any(dim=1))
    # first colliding position in binary order for a sample of colliding edges
    bad = torch.nonzero(~free.bool()).flatten()[:300].cpu().numpy()
    hit_h, N_h, first_h = hit.cpu().numpy(), N.cpu().numpy(), first.cpu().numpy()
    for e in bad:
        order = O.binary_indices(int(N_h[e]))
        pos = next(p for p, idx in enumerate(order) if 0 < idx < N_h[e] - 1 and hit_h[e, idx])
        assert first_h[e] == pos, (e, first_h[e], pos)
    assert (0.02 if name == "box_stacking" else 0.05) < free.float().mean().item() < 0.99
