"""The numpy oracle (oracle/oracle_abstract.py) against golden vectors produced by the
unmodified reference (scripts/make_golden_abstract.py) and the reference's own known answers."""
import numpy as np
import pytest

from oracle import oracle_abstract as OA


def test_binary_indices_known_answers():
    # tests/test.py:36-47 of the reference
    assert OA.binary_search_indices(1) == (0,)
    assert OA.binary_search_indices(2) == (0, 1)
    assert OA.binary_search_indices(3) == (1, 0, 2)
    assert OA.binary_search_indices(4) == (1, 0, 2, 3)
    assert OA.binary_search_indices(5) == (2, 0, 3, 1, 4)


def test_binary_indices_golden(golden):
    Ns, flat = golden["bin_N"], golden["bin_idx"]
    o = 0
    for n in Ns:
        assert tuple(flat[o:o + n]) == OA.binary_search_indices(int(n))
        o += n


def test_binary_index_closed_form_matches_bfs():
    for N in list(range(1, 700)) + [1000, 1023, 1024, 1025, 4097]:
        ref = OA.binary_search_indices(N)
        got = tuple(OA.binary_index_closed_form(N, p) for p in range(N))
        assert got == ref, N


def test_config_flags_golden(golden):
    sc = OA.AbstractScene.abstract_test()
    q, free = golden["cfg_q"], golden["cfg_free"]
    assert np.array_equal(sc.batch_flags(q), free)
    for i in range(0, len(q), 97):
        assert sc.is_collision_free(q[i]) == free[i]


def test_edge_flags_and_call_counts_golden(golden):
    sc = OA.AbstractScene.abstract_test()
    q1, q2 = golden["edge_q1"], golden["edge_q2"]
    for a, b, f, calls in zip(q1, q2, golden["edge_free"], golden["edge_calls"]):
        flag, first, checks = sc.is_edge_collision_free(a, b)
        assert flag == f and checks == calls
    flags, first = sc.batch_edge_flags(q1, q2)
    assert np.array_equal(flags, golden["edge_free"])


def test_edge_variants_golden(golden):
    sc = OA.AbstractScene.abstract_test()
    for a, b, v in zip(golden["edge_q1"][:100], golden["edge_q2"][:100], golden["edge_variants"]):
        got = [sc.is_edge_collision_free(a, b, include_endpoints=True)[0],
               sc.is_edge_collision_free(a, b, resolution=0.1)[0],
               sc.is_edge_collision_free(a, b, N_start=0, N_max=2)[0],
               sc.is_edge_collision_free(a, b, N_start=2, N_max=12)[0],
               sc.is_edge_collision_free(a, b, N_start=1, N_max=40, N=40)[0]]
        assert got == list(v)


def test_edge_known_call_counts(golden):
    # tests/test.py:61-90 of the reference: 3 / 1 / 9 / 11 checks (here ordered 1,3,9,11)
    sc = OA.AbstractScene.abstract_test()
    a, b = np.array([-1, 1, 1, 1.0]), np.array([-1, 1, 1, 0.0])
    got = [sc.is_edge_collision_free(a, b, resolution=r, include_endpoints=i)[2]
           for r, i in ((0.5, False), (0.5, True), (0.1, False), (0.1, True))]
    assert got == [1, 3, 9, 11] == list(golden["edge_known_counts"])


@pytest.mark.parametrize("name", ["d22", "d77", "d333", "d25", "d14", "d6666"])
def test_metrics_golden(golden, name):
    q, pts, sl = golden[f"met_{name}_q"], golden[f"met_{name}_pts"], golden[f"met_{name}_slices"]
    for metric in ("euclidean", "sum_euclidean", "max_euclidean", "max"):
        ref = golden[f"met_{name}_dist_{metric}"]
        got = OA.batch_config_dist(q, pts, sl, metric)
        # the reference's numba kernels are fastmath: reproducible to ~1 ulp only
        assert np.allclose(got, ref, rtol=4e-16, atol=0)
    for metric in ("euclidean", "max"):
        for red in ("max", "sum"):
            ref = golden[f"met_{name}_cost_{metric}_{red}"]
            got = OA.batch_config_cost(q[None, :] - pts, sl, metric, red)
            assert np.allclose(got, ref, rtol=1e-15, atol=0)


@pytest.mark.parametrize("metric", ["max_euclidean", "euclidean", "sum_euclidean", "max"])
def test_neighbour_selection_golden(golden, metric):
    corpus, qidx, sl, k = golden["knn_corpus"], golden["knn_qidx"], golden["knn_slices"], int(golden["knn_k"])
    assert k == OA.prm_k_star(len(corpus), corpus.shape[1])
    ref_idx = golden[f"knn_idx_{metric}"]
    rr, rcnt, ridx = golden[f"knn_rad_r_{metric}"], golden[f"knn_rad_cnt_{metric}"], golden[f"knn_rad_idx_{metric}"]
    o = 0
    for j, qi in enumerate(qidx):
        d = OA.batch_config_dist(corpus[qi], corpus, sl, metric)
        assert np.array_equal(OA.knn_indices(d, k), ref_idx[j])
        # the golden radius equals one of the reference's own distances, and those are only
        # reproducible to ~1 ulp (numba fastmath): elements within 4 ulp of r are "don't care"
        tie = np.abs(d - rr[j]) <= 4 * np.finfo(np.float64).eps * rr[j]
        got = set(OA.radius_indices(d, rr[j])) - set(np.nonzero(tie)[0])
        assert got == set(ridx[o:o + rcnt[j]]) - set(np.nonzero(tie)[0])
        o += rcnt[j]
