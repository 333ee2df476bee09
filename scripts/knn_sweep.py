import sys, numpy as np, torch
sys.path.insert(0, ".")
import os
from multirobot_pathplanning_benchmark_b200 import _lib
_lib.LIB_PATH = os.environ.get("VARIANT_LIB", _lib.LIB_PATH)  # experiment builds (scripts/build_variant.sh)
from multirobot_pathplanning_benchmark_b200 import knn as K
N = 100000
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.2, 3.2, (N, 24))).cuda()
for R in [int(a) for a in sys.argv[1:]] or (1, 2, 3, 4, 6, 8):
    d = 24 // R
    sl = [[r * d, (r + 1) * d] for r in range(R)]
    K.batch_knn(c[:4096], c, sl, "max_euclidean", 33, mode="tensor")
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); K.batch_knn(c, c, sl, "max_euclidean", 33, mode="tensor"); b.record(); b.synchronize()
    print("R", R, "%.2f ms" % a.elapsed_time(b))
