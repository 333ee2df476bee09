"""GPU parity, abstract environment (BASELINE config 1): CUDA kernels vs golden vectors of the
unmodified reference and vs the numpy oracle.  Bit-exact (fp64 on the device)."""
import numpy as np
import pytest
import torch

from oracle import oracle_abstract as OA

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def backend(cuda_lib):
    from multirobot_pathplanning_benchmark_b200.backend import AbstractBackend
    return AbstractBackend(2, 2, [0.1, 0.1], spheres=[([0.0, 0.0], 0.2)], rects_minmax=[([-0.25, 0.15], [0.25, 0.65])])


def test_config_flags_match_reference_golden(backend, golden):
    q = torch.from_numpy(golden["cfg_q"]).cuda()
    got = backend.check_configs(q).cpu().numpy()
    assert np.array_equal(got, golden["cfg_free"])  # includes 6000 boundary-adversarial samples


def test_config_flags_match_oracle_large(backend):
    sc = OA.AbstractScene.abstract_test()
    np.random.seed(0)
    q = np.random.uniform(-2, 2, (1_000_003, 4))  # ragged size on purpose
    got = backend.check_configs(torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(got, sc.batch_flags(q))


def test_empty_batches(backend):
    assert backend.check_configs(torch.empty(0, 4, dtype=torch.float64, device="cuda")).numel() == 0
    f, p = backend.check_edges(torch.empty(0, 4, dtype=torch.float64, device="cuda"),
                               torch.empty(0, 4, dtype=torch.float64, device="cuda"), 0.01)
    assert f.numel() == 0 and p.numel() == 0


def test_edge_flags_match_reference_golden(backend, golden):
    q1, q2 = torch.from_numpy(golden["edge_q1"]).cuda(), torch.from_numpy(golden["edge_q2"]).cuda()
    free, first = backend.check_edges(q1, q2, 0.01)
    assert np.array_equal(free.cpu().numpy(), golden["edge_free"])
    # first colliding position must be what the sequential reference loop hits first
    sc = OA.AbstractScene.abstract_test()
    ofree, ofirst = sc.batch_edge_flags(golden["edge_q1"], golden["edge_q2"], 0.01)
    assert np.array_equal(first.cpu().numpy(), ofirst)


def test_edge_variants_match_reference_golden(backend, golden):
    q1, q2 = torch.from_numpy(golden["edge_q1"][:100]).cuda(), torch.from_numpy(golden["edge_q2"][:100]).cuda()
    v = golden["edge_variants"]
    assert np.array_equal(backend.check_edges(q1, q2, 0.01, include_endpoints=True)[0].cpu().numpy(), v[:, 0])
    assert np.array_equal(backend.check_edges(q1, q2, 0.1)[0].cpu().numpy(), v[:, 1])
    assert np.array_equal(backend.check_edges(q1, q2, 0.01, n_start=0, n_max=2)[0].cpu().numpy(), v[:, 2])
    assert np.array_equal(backend.check_edges(q1, q2, 0.01, n_start=2, n_max=12)[0].cpu().numpy(), v[:, 3])
    N = torch.full((100,), 40, dtype=torch.int32, device="cuda")
    assert np.array_equal(backend.check_edges(q1, q2, 0.01, N=N, n_start=1, n_max=40)[0].cpu().numpy(), v[:, 4])


def test_edges_large_against_oracle(backend):
    sc = OA.AbstractScene.abstract_test()
    rng = np.random.default_rng(5)
    q1 = rng.uniform(-2, 2, (3000, 4))
    q2 = q1 + rng.uniform(-1, 1, (3000, 4))
    free, first = backend.check_edges(torch.from_numpy(q1).cuda(), torch.from_numpy(q2).cuda(), 0.01)
    ofree, ofirst = sc.batch_edge_flags(q1, q2, 0.01)
    assert np.array_equal(free.cpu().numpy(), ofree)
    assert np.array_equal(first.cpu().numpy(), ofirst)


def test_higher_dimensional_abstract_env(cuda_lib):
    """abstract.center_rect_10d geometry (abstract_env.py:428-440): numpy's pairwise sum order."""
    from multirobot_pathplanning_benchmark_b200.backend import AbstractBackend
    n = 10
    be = AbstractBackend(2, n, [0.1, 0.1], rects_minmax=[([-0.25] * n, [0.25] * n)])
    sc = OA.AbstractScene(2, n, [0.1, 0.1], rects=[([0.0] * n, [0.5] * n)])
    rng = np.random.default_rng(6)
    q = rng.uniform(-0.36, 0.36, (200000, 2 * n))
    got = be.check_configs(torch.from_numpy(q).cuda()).cpu().numpy()
    assert np.array_equal(got, sc.batch_flags(q))
    assert 0.02 < got.mean() < 0.98


def test_host_buffer_queries_are_bit_exact_too(backend):
    """mrb200_abstract_query_*_host (single-query seam, mapped pinned staging) against the device-buffer calls and the oracle"""
    sc = OA.AbstractScene.abstract_test()
    rng = np.random.default_rng(8)
    q = rng.uniform(-2, 2, (500, 4))
    want = sc.batch_flags(q)
    assert np.array_equal(backend.query_configs_host(q), want)
    for i in range(0, 500, 37):
        assert backend.query_configs_host(q[i:i + 1])[0] == want[i]
    q2 = q + rng.uniform(-1, 1, q.shape)
    for kw in ({}, {"n_start": 3, "n_max": 11}, {"include_endpoints": True}):
        f, p = backend.query_edges_host(q, q2, 0.01, **kw)
        of, op = sc.batch_edge_flags(q, q2, 0.01, include_endpoints=kw.get("include_endpoints", False),
                                     N_start=kw.get("n_start", 0), N_max=kw.get("n_max"))
        assert np.array_equal(f, of) and np.array_equal(p, op)
    f1, p1 = backend.query_edges_host(q[:1], q2[:1], 0.01, N=np.array([40], np.int32))
    of, op = sc.batch_edge_flags(q[:1], q2[:1], 0.01, Ns=np.array([40]))
    assert f1[0] == of[0] and p1[0] == op[0]
