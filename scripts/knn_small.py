import sys, numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import knn as K
N = 8192
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.2, 3.2, (N, 24))).cuda()
sl = [[0, 6], [6, 12], [12, 18], [18, 24]]
a = K.batch_knn(c[:1024], c, sl, "max_euclidean", 33, mode="tensor", return_dist=False)
b = K.batch_knn(c[:1024], c, sl, "max_euclidean", 33, mode="exact", return_dist=False)
off, idx = K.batch_radius(c[:512], c, 2.0, sl, "max_euclidean")
torch.cuda.synchronize()
print("tensor == exact:", bool(torch.equal(a, b)), "radius rows", int(off[-1]))
