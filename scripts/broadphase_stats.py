"""Offline analysis aid: how many (configuration, pair) items survive the kernels' bounding-volume
broadphase, per pair type, for uniform random configurations of a named scene."""
import sys
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O

name = sys.argv[1] if len(sys.argv) > 1 else "box_rearrangement"
mk, kw = SCENES[name]
sc = mk(); cs = S.compile_blob(sc, kw["tol"])
b = cs.blob64
I = lambda i: int(b[i])
offS = I(S.H_OFF_SHAPES); ns = cs.n_moving + cs.n_static
rng = np.random.default_rng(0); lim = sc.limits()
B = 400
qs = rng.uniform(lim[0], lim[1], (B, sc.dof))
rows = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)
core = rows[:, 0].astype(np.int64); rad = rows[:, 3].view(np.float64); bound = rows[:, 19].view(np.float64)
tot = np.zeros(S.NUM_PAIR_TYPES); surv = np.zeros(S.NUM_PAIR_TYPES); pen = np.zeros(S.NUM_PAIR_TYPES)
for q in qs:
    W = O.world_shapes(b, q, ns)
    ctr = np.where((core == 1)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])
    for t in range(S.NUM_PAIR_TYPES):
        n, off = I(S.H_N_PAIRS + t), I(S.H_OFF_PAIRS + t)
        for i in range(n):
            pk = I(off + i); a, c, kind = pk & 0xffff, (pk >> 16) & 0xfff, pk >> 28
            tot[t] += 1
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
            surv[t] += ok
            if ok:
                pen[t] += O.pair_distance(t, W[a], W[c], rad[a] + rad[c]) < 0
for t in range(S.NUM_PAIR_TYPES):
    if tot[t]:
        print(f"{S.PAIR_TYPE_NAMES[t]:12s} pairs/cfg {tot[t]/B:7.1f}  survive {surv[t]/tot[t]*100:5.1f}%  ({surv[t]/B:6.1f}/cfg)  penetrating {pen[t]/B:5.2f}/cfg")
print("total survivors per config", surv.sum() / B, "of", tot.sum() / B)
