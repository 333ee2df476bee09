"""k-NN micro-benchmark (BASELINE config 4 shape): tensor-core path vs exact path."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import knn as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.2, 3.2, (N, 24))).cuda()
sl = [[0, 6], [6, 12], [12, 18], [18, 24]]
for metric, s in (("max_euclidean", sl), ("euclidean", None)):
    for mode in ("tensor", "exact"):
        K.batch_knn(c[:4096], c, s, metric, 33, mode=mode)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        i, d = K.batch_knn(c, c, s, metric, 33, mode=mode)
        b.record(); b.synchronize()
        print(metric, mode, "%.2f ms" % a.elapsed_time(b), "%.3g queries/s" % (N / a.elapsed_time(b) * 1e3))
        if mode == "tensor": it = i
        else: print("   identical to tensor path:", bool(torch.equal(it, i)))
