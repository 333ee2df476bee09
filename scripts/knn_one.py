import sys, numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import knn as K
N = 100000; R = int(sys.argv[1]) if len(sys.argv) > 1 else 4
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.2, 3.2, (N, 24))).cuda()
d = 24 // R; sl = [[r * d, (r + 1) * d] for r in range(R)]
K.batch_knn(c[:4096], c, sl, "max_euclidean", 33, mode="tensor")
torch.cuda.synchronize()
K.batch_knn(c, c, sl, "max_euclidean", 33, mode="tensor")
torch.cuda.synchronize()
