"""Offline analysis aid (CPU, oracle FK): broadphase records that never pass their bounding test over random joint vectors
(inside and far outside the limits).  usage: python scripts/record_census.py SCENE [B]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O
name = sys.argv[1]
mk, kw = SCENES[name]
sc = mk(); cs = S.compile_blob(sc, kw["tol"])
b = cs.blob64
I = lambda i: int(b[i])
offS = I(S.H_OFF_SHAPES); ns = cs.n_moving + cs.n_static; nm = cs.n_moving
rows = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)
core = rows[:, 0].astype(np.int64); rad = rows[:, 3].view(np.float64); bound = rows[:, 19].view(np.float64)
recs = []
for t in range(6):
    for k in range(S.BP_SUBLISTS):
        off, n = I(S.H_BP + (t * S.BP_SUBLISTS + k) * 2), I(S.H_BP + (t * S.BP_SUBLISTS + k) * 2 + 1)
        for i in range(n):
            pk = I(I(S.H_IDS_BASE) + (off - I(S.H_REC_BASE)) // 2 + i)
            recs.append((t, k, pk & 0xffff, (pk >> 16) & 0xfff))
rng = np.random.default_rng(0); lim = sc.limits()
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4000
passes = np.zeros(len(recs)); mind = np.full(len(recs), np.inf)
qs = np.concatenate([rng.uniform(lim[0], lim[1], (B, sc.dof)), rng.uniform(-7, 7, (B // 2, sc.dof))])
for q in qs:
    W = O.world_shapes(b, q, ns)
    ctr = np.where((core == 1)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])
    for j, (t, k, a, c) in enumerate(recs):
        x, y = (a, c) if a < nm else (c, a)
        if k < 2:
            gap = np.linalg.norm(ctr[x] - ctr[y]) - bound[x] - bound[y]
        else:
            R = W[y, 3:12].reshape(3, 3); h = W[y, 12:15]
            l = np.abs(R.T @ (ctr[x] - W[y, :3]))
            if core[x] == 1:
                e = np.abs(R.T @ (0.5 * (W[x, 3:6] - W[x, :3])))
                gap = np.max(l - e - h) - rad[x] - rad[y]
            else:
                gap = np.max(l - h) - bound[x] - rad[y]
        mind[j] = min(mind[j], gap)
        passes[j] += gap < 1e-3
never = passes == 0
print(name, "records", len(recs), "never passing in", len(qs), "joint vectors:", int(never.sum()), "; passing < 0.1 %:", int((passes / len(qs) < 1e-3).sum()))
order = np.argsort(-mind)
for j in order[:12]:
    t, k, a, c = recs[j]
    print(f"  closest approach {mind[j]:7.3f}  type {t} sub {k}  {cs.shape_names[a]:24s} {cs.shape_names[c]}")
