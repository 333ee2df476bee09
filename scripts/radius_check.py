"""GPU box: r-disc search on the tensor-core candidate generator against the fp64 kernels (identical CSR), timing.
usage: python scripts/radius_check.py [N]"""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import knn as K
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
c = torch.from_numpy(np.random.RandomState(5).uniform(-3.28, 3.28, (N, 24))).cuda()
sl = [[6 * r, 6 * r + 6] for r in range(4)]
_, d33 = K.batch_knn(c[:2048].contiguous(), c, sl, "max_euclidean", 33)
r = float(d33[:, -1].median().item())
off, idx = K.batch_radius(c, c, r, sl, "max_euclidean", mode="tensor")
st = dict(K.LAST_STATS)
offe, idxe = K.batch_radius(c, c, r, sl, "max_euclidean", mode="exact")
same = torch.equal(off, offe) and torch.equal(idx, idxe)
cnt = (off[1:] - off[:-1])
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    K.batch_radius(c, c, r, sl, "max_euclidean", mode="tensor")
b.record(); b.synchronize()
print(f"N={N} r={r:.4f} mean neighbours {cnt.float().mean().item():.1f} max {cnt.max().item()} overflow rows {st.get('radius_overflow_rows')}: "
      f"tensor == exact: {same}; tensor {a.elapsed_time(b)/3:.3f} ms", flush=True)
# single-row queries (RRT* / IT* near): the fp64 kernels
for Q1 in (1, 16, 72):
    qq = c[:Q1].contiguous()
    K.batch_radius(qq, c, r, sl, "max_euclidean", mode="exact")
    torch.cuda.synchronize()
    a.record()
    for _ in range(20):
        o1, i1 = K.batch_radius(qq, c, r, sl, "max_euclidean", mode="exact")
    b.record(); b.synchronize()
    ok = torch.equal(i1, idx[: int(off[Q1].item())])
    print(f"Q={Q1}: exact path {a.elapsed_time(b)/20*1000:.1f} us per call, equals the batch result: {ok}", flush=True)
