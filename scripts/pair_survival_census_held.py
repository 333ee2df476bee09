"""pair_survival_census.py for a mode in which PARENT holds CHILD.  usage: python scripts/pair_survival_census_held.py SCENE PARENT CHILD"""
import sys, collections
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O
name, parent, child = sys.argv[1:4]
mk, kw = SCENES[name]
sc = mk()
rng = np.random.default_rng(8); lim = sc.limits()
cs0 = S.compile_blob(sc, kw["tol"])
# a free configuration to grasp at
ns0 = cs0.n_moving + cs0.n_static
for _ in range(2000):
    q0 = rng.uniform(lim[0], lim[1])
    f, *_ = O.check_configs(cs0.blob64, q0[None].astype(np.float64))
    if f[0]:
        break
sc.attach(parent, child, q0)
cs = S.compile_blob(sc, kw["tol"])
skipped = {frozenset(p) for p in cs.unreachable_pairs}
b = cs.blob64
I = lambda i: int(b[i])
offS = I(S.H_OFF_SHAPES); ns = cs.n_moving + cs.n_static; nm = cs.n_moving
rows = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)
core = rows[:, 0].astype(np.int64); rad = rows[:, 3].view(np.float64); bound = rows[:, 19].view(np.float64)
B = 200
cnt = collections.Counter(); dmin = {}; dmax = {}
for q in rng.uniform(lim[0], lim[1], (B, sc.dof)):
    W = O.world_shapes(b, q, ns)
    ctr = np.where((core == 1)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])
    for t in range(6):
        n, off = I(S.H_N_PAIRS + t), I(S.H_OFF_PAIRS + t)
        for i in range(n):
            pk = I(off + i); a, c, kind = pk & 0xffff, (pk >> 16) & 0xfff, pk >> 28
            if frozenset((cs.shape_names[a], cs.shape_names[c])) in skipped:
                continue
            if kind == 0:
                ok = np.linalg.norm(ctr[a] - ctr[c]) < bound[a] + bound[c] + 1e-3
            else:
                R = W[c, 3:12].reshape(3, 3); h = W[c, 12:15]
                l = R.T @ (ctr[a] - W[c, :3])
                if kind == 2:
                    e = np.abs(R.T @ (0.5 * (W[a, 3:6] - W[a, :3])))
                    lb = np.max(np.abs(l) - e - h) - rad[a]
                else:
                    lb = np.max(np.abs(l) - h) - bound[a]
                ok = lb - rad[c] < 1e-3
            if ok:
                key = (t, cs.shape_names[a], cs.shape_names[c], kind)
                cnt[key] += 1
                d = O.pair_distance(t, W[a], W[c], rad[a] + rad[c])
                dmin[key] = min(dmin.get(key, 9), d); dmax[key] = max(dmax.get(key, -9), d)
print("survivors per configuration", sum(cnt.values()) / B)
for key, v in cnt.most_common(12):
    print(f"{v / B:5.2f}/cfg  type {key[0]} kind {key[3]}  {key[1]:24s} {key[2]:24s} exact distance in [{dmin[key]:.4f}, {dmax[key]:.4f}]")
