"""Offline analysis aid (CPU, oracle FK): which collidable pairs pass the kernels' broadphase bound most often over uniform
configurations of a named scene, and the range of their exact distance -- pairs that pass (nearly) always and never collide
are what the exact compile-time rules of scene.py compile_blob remove.  usage: python scripts/pair_survival_census.py SCENE"""
import sys, collections
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
rng = np.random.default_rng(0); lim = sc.limits()
B = 200
cnt = collections.Counter(); dmin = {}; dmax = {}
for q in rng.uniform(lim[0], lim[1], (B, sc.dof)):
    W = O.world_shapes(b, q, ns)
    ctr = np.where((core == 1)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])
    for t in range(6):
        n, off = I(S.H_N_PAIRS + t), I(S.H_OFF_PAIRS + t)
        for i in range(n):
            pk = I(off + i); a, c, kind = pk & 0xffff, (pk >> 16) & 0xfff, pk >> 28
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
for key, v in cnt.most_common(14):
    print(f"{v / B:5.2f}/cfg  type {key[0]} kind {key[3]}  {key[1]:28s} {key[2]:28s} exact distance in [{dmin[key]:.4f}, {dmax[key]:.4f}]")
