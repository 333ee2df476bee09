"""GPU box: quick correctness + timing of the tensor-core k-NN path against the exact path.
usage: python scripts/knn_check.py [N] [metric] [R]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import knn as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
metric = sys.argv[2] if len(sys.argv) > 2 else "max_euclidean"
R = int(sys.argv[3]) if len(sys.argv) > 3 else 4
D = 6 * R if metric == "max_euclidean" else 24
sl = [[6 * r, 6 * r + 6] for r in range(R)] if metric == "max_euclidean" else None
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.28, 3.28, (N, D))).cuda()
k = K.prm_k_star(N, D)
for Q in (min(N, 4096), N):
    q = c[:Q].contiguous()
    it, dt = K.batch_knn(q, c, sl, metric, k, mode="tensor", stats=True)
    st = dict(K.LAST_STATS)
    ie, de = K.batch_knn(q, c, sl, metric, k, mode="exact")
    same = torch.equal(it, ie) and torch.equal(dt, de)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3):
        K.batch_knn(q, c, sl, metric, k, mode="tensor")
    b.record(); b.synchronize()
    print(f"N={N} Q={Q} D={D} k={k} {metric}: tensor == exact: {same}; uncertified rows {st}; tensor {a.elapsed_time(b)/3:.3f} ms", flush=True)
