"""CPU estimate of how many interpolation points of collision-free planner-like edges a clearance-based skip rule could
retire without FK.  Rule: a sample with minimum pair distance c proves the next m samples
free while the accumulated motion bound  sum_robots sum_j rho_j |dq_j|  stays below c (rho_j: largest distance from joint
j's axis to any point of the links it carries, from the scene compiler's reach analysis).
usage: python scripts/edge_skip_sim.py [scene] [edges]"""
import sys
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from oracle import oracle_scene as O

name = sys.argv[1] if len(sys.argv) > 1 else "box_rearrangement"
E_ = int(sys.argv[2]) if len(sys.argv) > 2 else 400
mk, kw = SCENES[name]
sc = mk()
cs = S.compile_blob(sc, kw["tol"])
lim = sc.limits()
rng = np.random.default_rng(0)
# lever arms: numerically, the largest displacement of any shape centre per unit joint motion, plus nothing for the
# shape's own extent (rotation moves its far end by at most bound * |dq| more -> add the bounding radius)
ns = cs.n_moving + cs.n_static
b = cs.blob64
offS = int(b[S.H_OFF_SHAPES])
bound = b[offS: offS + ns * S.SHAPE_WORDS].reshape(ns, S.SHAPE_WORDS)[:, 19].view(np.float64)
rho = np.zeros(sc.dof)
for _ in range(60):
    q = rng.uniform(lim[0], lim[1])
    W0 = O.world_shapes(b, q, ns)
    for j in range(sc.dof):
        dq = np.zeros(sc.dof); dq[j] = 1e-4
        W1 = O.world_shapes(b, q + dq, ns)
        mv = np.linalg.norm(W1[:cs.n_moving, :3] - W0[:cs.n_moving, :3], axis=1) / 1e-4
        moved = mv > 1e-9
        if moved.any():
            rho[j] = max(rho[j], np.max(mv[moved] + bound[:cs.n_moving][moved]))
print(name, "lever arms per joint:", np.round(rho, 2))
for span in (0.05, 0.2, 0.6):
    tot = ev = nfree = 0
    tries = 0
    while nfree < E_ and tries < 20 * E_:
        tries += 1
        q1 = rng.uniform(lim[0], lim[1])
        q2 = np.clip(q1 + rng.uniform(-span, span, sc.dof), lim[0], lim[1])
        N = max(2, int(np.max(np.abs(q2 - q1)) / kw["resolution"]) + 1)
        qs = q1 + (q2 - q1) * (np.arange(N) / (N - 1))[:, None]
        free, pen, mind = O.check_configs(cs.blob64, qs)
        if not free.all():
            continue
        nfree += 1
        step = float(rho @ np.abs(q2 - q1)) / (N - 1)      # clearance lost per interpolation step, at most
        i = 0
        while i < N:
            ev += 1
            c = mind[i] if pen[i] == 0 else 0.0
            i += 1 + int(max(c, 0.0) / step) if step > 0 else N
        tot += N
    print(f"span {span:4.2f}: {nfree} free edges, {tot / max(nfree, 1):6.1f} points per edge, evaluated {ev / max(tot, 1) * 100:5.1f} % of them")
