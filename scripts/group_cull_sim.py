"""Offline estimate for a warp-coherent group cull of the moving-moving broadphase records: records grouped by the pair of
frames their shapes ride on, one bounding test per group (anchor shape of each frame, radius over the frame's shapes), a group's
records only run if ANY of the tile's 32 samples passes.  Tiles: 32 uniform configurations, or the samples of local edges
(+-0.2 per joint, resolution as in the scene, binary order irrelevant for the estimate).
usage: python scripts/group_cull_sim.py SCENE"""
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
offS = I(S.H_OFF_SHAPES); ns = cs.n_moving + cs.n_static; nm = cs.n_moving
rows = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)
core = rows[:, 0].astype(np.int64); fid = rows[:, 1].astype(np.int64); wo = rows[:, 2].astype(np.int64)
bound = rows[:, 19].view(np.float64)
wo2shape = {int(wo[i]) * 128: i for i in range(nm)}
# records of sublist 0 (both moving), all queued types
recs = []
n_rec_total = 0
for t in range(6):
    for k in range(S.BP_SUBLISTS):
        off, n = I(S.H_BP + (t * S.BP_SUBLISTS + k) * 2), I(S.H_BP + (t * S.BP_SUBLISTS + k) * 2 + 1)
        n_rec_total += n
        if k == 0:
            for i in range(n):
                w = I(off + 2 * i)
                recs.append((t, wo2shape[w & 0xffff], wo2shape[w >> 16]))
print(name, "records", n_rec_total, "moving-moving", len(recs))
SEG = int(sys.argv[2]) if len(sys.argv) > 2 else 0    # 0: one group per frame pair; S > 0: frames of a robot merged into S segments
rob = b[I(S.H_OFF_SHAPE_ROBOT): I(S.H_OFF_SHAPE_ROBOT) + ns].astype(np.int64)
def seg_of(i):
    if SEG == 0:
        return int(fid[i])
    fr = sorted({int(fid[j]) for j in range(nm) if rob[j] == rob[i]})
    return (int(rob[i]), fr.index(int(fid[i])) * SEG // len(fr))
groups = {}
for (t, x, y) in recs:
    groups.setdefault((t, seg_of(x), seg_of(y)), []).append((x, y))
sizes = [len(v) for v in groups.values()]
print("groups", len(groups), "mean size", np.mean(sizes), "max", max(sizes))
rng = np.random.default_rng(0); lim = sc.limits()

def centres(q):
    W = O.world_shapes(b, q, ns)
    return np.where((core == 1)[:, None], 0.5 * (W[:, :3] + W[:, 3:6]), W[:, :3])

# group anchors and radii from one configuration (rigid frames: distances between shapes of a frame are constant)
c0 = centres(rng.uniform(lim[0], lim[1]))
ginfo = []
for (t, fx, fy), prs in groups.items():
    xs = sorted({x for x, _ in prs}); ys = sorted({y for _, y in prs})
    ax, ay = xs[0], ys[0]
    rx = max(np.linalg.norm(c0[x] - c0[ax]) + bound[x] for x in xs)
    ry = max(np.linalg.norm(c0[y] - c0[ay]) + bound[y] for y in ys)
    ginfo.append((ax, ay, rx + ry + 1e-3, prs))

def tile_stats(Q):
    C = np.stack([centres(q) for q in Q])        # [32, ns, 3]
    rec_tests = grp_tests = 0
    for ax, ay, thr, prs in ginfo:
        d = np.linalg.norm(C[:, ax] - C[:, ay], axis=1)
        grp_tests += 1
        if (d < thr).any():
            rec_tests += len(prs)
    return grp_tests, rec_tests

res = kw["resolution"]
for kind in ("uniform configurations", "local edges"):
    g = r = 0
    for _ in range(30):
        if kind.startswith("uniform"):
            Q = rng.uniform(lim[0], lim[1], (32, sc.dof))
        else:
            Q = []
            while len(Q) < 32:
                a = rng.uniform(lim[0], lim[1]); e = np.clip(a + rng.uniform(-0.2, 0.2, sc.dof), lim[0], lim[1])
                N = max(2, int(np.max(np.abs(e - a)) / res) + 1)
                for i in range(N):
                    Q.append(a + (e - a) * i / (N - 1))
            Q = np.array(Q[:32])
        gt, rt = tile_stats(Q)
        g += gt; r += rt
    print(f"{kind:24s}: per tile {g/30:.0f} group tests + {r/30:.0f} record tests instead of {len(recs)} record tests "
          f"(other sublists: {n_rec_total - len(recs)} records unchanged)")
