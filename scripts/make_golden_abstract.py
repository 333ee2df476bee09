"""Generate tests/golden/abstract_golden.npz by running the UNMODIFIED reference
(/root/reference, imported with MagicMock stubs for its absent GUI / rai deps --
recipe from SURVEY.md section 8c).  Runs only in the build container; the .npz
travels, the reference does not.

    python scripts/make_golden_abstract.py
"""
import os
import sys
from unittest.mock import MagicMock, patch

for _n in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "matplotlib.patches",
           "matplotlib.collections", "mpl_toolkits", "mpl_toolkits.mplot3d", "robotic",
           "simple_parsing"):
    sys.modules.setdefault(_n, MagicMock())
sys.path.insert(0, "/root/reference/src")

import numpy as np  # noqa: E402

from multi_robot_multi_goal_planning.problems import get_env_by_name  # noqa: E402
from multi_robot_multi_goal_planning.problems.planning_env import (  # noqa: E402
    generate_binary_search_indices, State)
from multi_robot_multi_goal_planning.problems.core.configuration import (  # noqa: E402
    NpConfiguration, batch_config_dist, batch_config_cost, config_dist)

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "abstract_golden.npz")


def main():
    out = {}
    env = get_env_by_name("abstract.test")
    m = env.start_mode

    # --- binary search indices: every N in 1..64 plus a few large ones (flattened, ragged)
    Ns = list(range(1, 65)) + [100, 127, 128, 129, 255, 256, 257, 333, 500, 657, 1000, 1023, 1024, 1025]
    out["bin_N"] = np.array(Ns, np.int64)
    out["bin_idx"] = np.concatenate([np.array(generate_binary_search_indices(n), np.int64) for n in Ns])

    # --- config flags: reference's own sampler (planning_env.py:1697-1708), seed 0
    np.random.seed(0)
    B = 20000
    qs = np.random.uniform(env.limits[0], env.limits[1], (B, env.limits.shape[1]))
    # add adversarial near-boundary samples (robot-robot, sphere obstacle, rectangle)
    rng = np.random.default_rng(1)
    adv = []
    for _ in range(2000):
        a = rng.uniform(-2, 2, 2)
        th = rng.uniform(0, 2 * np.pi)
        r = 0.2 + rng.choice([0.0, 1e-12, -1e-12, 1e-9, -1e-9, 1e-16])
        b = a + r * np.array([np.cos(th), np.sin(th)])
        adv.append(np.concatenate([a, b]))
        # agent 0 near sphere obstacle boundary (R + r = 0.3)
        r2 = 0.3 + rng.choice([0.0, 1e-12, -1e-12, 1e-16])
        p = r2 * np.array([np.cos(th), np.sin(th)])
        adv.append(np.concatenate([p, rng.uniform(1, 2, 2)]))
        # agent 1 touching rectangle face: rect is [-.25,.25]x[.15,.65]; x = .25 + .1
        y = rng.uniform(0.15, 0.65)
        x = 0.25 + 0.1 + rng.choice([0.0, 1e-12, -1e-12, 1e-17])
        adv.append(np.concatenate([rng.uniform(-2, -1, 2), [x, y]]))
    qs = np.vstack([qs, np.array(adv)])
    flags = np.array([env.is_collision_free(env.start_pos.from_flat(q), m) for q in qs])
    out["cfg_q"] = qs
    out["cfg_free"] = flags

    # --- edges: uniform-uniform, default resolution, with first-collision position and the
    # number of is_collision_free calls the reference made (tests/test.py:61-90 style)
    np.random.seed(2)
    E = 400
    q1 = np.random.uniform(env.limits[0], env.limits[1], (E, 4))
    q2 = np.random.uniform(env.limits[0], env.limits[1], (E, 4))
    # make a third of them short so many are free
    q2[::3] = q1[::3] + np.random.uniform(-0.3, 0.3, (len(q1[::3]), 4))
    e_free, e_calls = [], []
    for a, b in zip(q1, q2):
        with patch.object(type(env), "is_collision_free", autospec=True,
                          side_effect=type(env).is_collision_free) as spy:
            f = env.is_edge_collision_free(env.start_pos.from_flat(a), env.start_pos.from_flat(b), m)
            e_free.append(f)
            e_calls.append(spy.call_count)
    out["edge_q1"], out["edge_q2"] = q1, q2
    out["edge_free"] = np.array(e_free)
    out["edge_calls"] = np.array(e_calls, np.int64)
    # variants: include_endpoints / N_start,N_max windows / coarse resolution
    var = []
    for a, b in zip(q1[:100], q2[:100]):
        ca, cb = env.start_pos.from_flat(a), env.start_pos.from_flat(b)
        var.append([
            env.is_edge_collision_free(ca, cb, m, include_endpoints=True),
            env.is_edge_collision_free(ca, cb, m, resolution=0.1),
            env.is_edge_collision_free(ca, cb, m, N_start=0, N_max=2),
            env.is_edge_collision_free(ca, cb, m, N_start=2, N_max=12),
            env.is_edge_collision_free(ca, cb, m, N_start=1, N_max=40, N=40),
        ])
    out["edge_variants"] = np.array(var)

    # known-answer call counts from the reference's own tests (tests/test.py:61-90)
    cnts = []
    for res, inc in ((0.5, False), (0.5, True), (0.1, False), (0.1, True)):
        with patch.object(type(env), "is_collision_free", autospec=True,
                          side_effect=type(env).is_collision_free) as spy:
            env.is_edge_collision_free(NpConfiguration.from_list([[-1, 1], [1, 1]]),
                                       NpConfiguration.from_list([[-1, 1], [1, 0]]),
                                       m, resolution=res, include_endpoints=inc)
            cnts.append(spy.call_count)
    out["edge_known_counts"] = np.array(cnts, np.int64)

    # --- metrics (configuration.py:303-329, 437-510) on the dims of tests/test_config.py
    rng = np.random.default_rng(3)
    for name, dims in (("d22", [2, 2]), ("d77", [7, 7]), ("d333", [3, 3, 3]), ("d25", [2, 5]),
                       ("d14", [14]), ("d6666", [6, 6, 6, 6])):
        D = sum(dims)
        q = NpConfiguration.from_list([rng.uniform(-3, 3, d) for d in dims])
        pts = rng.uniform(-3, 3, (257, D))
        out[f"met_{name}_q"] = q.state()
        out[f"met_{name}_pts"] = pts
        out[f"met_{name}_slices"] = np.asarray(q._array_slice, np.int64)
        for metric in ("euclidean", "sum_euclidean", "max_euclidean", "max"):
            out[f"met_{name}_dist_{metric}"] = batch_config_dist(q, pts, metric)
        for metric in ("euclidean", "max"):
            for red in ("max", "sum"):
                out[f"met_{name}_cost_{metric}_{red}"] = batch_config_cost(q, pts, metric, red)

    # --- PRM neighbour selection (prm_graph.py:440-447, 488-500) on a D=24 corpus
    rng = np.random.default_rng(4)
    dims = [6, 6, 6, 6]
    N, Q = 4096, 64
    corpus = rng.uniform(-3.2, 3.2, (N, 24))
    qidx = rng.choice(N, Q, replace=False)
    out["knn_corpus"], out["knn_qidx"] = corpus, qidx
    out["knn_slices"] = np.array([[0, 6], [6, 12], [12, 18], [18, 24]], np.int64)
    k_star = int(np.e * (1 + 1 / 24) * np.log(N)) + 1
    out["knn_k"] = np.array(k_star)
    for metric in ("max_euclidean", "euclidean", "sum_euclidean", "max"):
        idxs, rad = [], []
        for qi in qidx:
            q = NpConfiguration.from_list([corpus[qi, s:s + 6] for s in range(0, 24, 6)])
            d = batch_config_dist(q, corpus, metric)
            topk = np.argpartition(d, k_star - 1)[:k_star]
            topk = topk[np.argsort(d[topk])]
            idxs.append(topk)
            r = np.sort(d)[40]  # a radius that keeps ~40 neighbours
            rad.append((r, np.where(d < r)[0]))
        out[f"knn_idx_{metric}"] = np.array(idxs, np.int64)
        out[f"knn_rad_r_{metric}"] = np.array([r for r, _ in rad])
        out[f"knn_rad_cnt_{metric}"] = np.array([len(i) for _, i in rad], np.int64)
        out[f"knn_rad_idx_{metric}"] = np.concatenate([i for _, i in rad]).astype(np.int64)

    # --- path check counts (tests/test.py:93-122): 5/5/21/21
    path = [State(NpConfiguration.from_list([[-1, 1], [1, 1]]), m),
            State(NpConfiguration.from_list([[-1, 1], [1, 0]]), m),
            State(NpConfiguration.from_list([[-1, 1], [1, 1]]), m)]
    pc = []
    for res, order in ((0.5, False), (0.5, True), (0.1, False), (0.1, True)):
        with patch.object(type(env), "is_collision_free", autospec=True,
                          side_effect=type(env).is_collision_free) as spy:
            ok = env.is_path_collision_free(path, resolution=res, check_edges_in_order=order)
            pc.append((int(ok), spy.call_count))
    out["path_known"] = np.array(pc, np.int64)

    np.savez_compressed(OUT, **out)
    print("wrote", os.path.abspath(OUT), {k: v.shape for k, v in out.items() if k.startswith(("cfg", "edge"))})
    print("free frac", flags.mean(), "edge free frac", np.mean(e_free), "known counts", cnts, "path", pc)


if __name__ == "__main__":
    main()
