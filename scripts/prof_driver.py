"""One workload, a few launches -- the command `ncu` wraps on the GPU box (scripts/capture_all.sh).
usage: python scripts/prof_driver.py configs SCENE B [two_phase: auto|always|never] | edges SCENE E [local|uniform] | knn N [tensor|exact] | radius N"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S, knn as K
from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
from multirobot_pathplanning_benchmark_b200.scenes import SCENES

what = sys.argv[1]
if what in ("configs", "edges"):
    name, n = sys.argv[2], int(sys.argv[3])
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    be = SceneBackend(max_modes=2)
    be.set_mode(0, cs)
    lim = sc.limits()
    rng = np.random.RandomState(1000)
    if what == "configs":
        q = torch.from_numpy(rng.uniform(lim[0], lim[1], (n, sc.dof)).astype(np.float32)).cuda()
        if len(sys.argv) > 4:
            be.set_two_phase(0, sys.argv[4])
        for _ in range(5):
            f = be.check_configs(0, q)
            torch.cuda.synchronize()   # (lets the slot's two-phase policy settle before the captured launch)
        print(name, n, "free", f.float().mean().item())
    else:
        kind = sys.argv[4] if len(sys.argv) > 4 else "local"
        a = torch.from_numpy(rng.uniform(lim[0], lim[1], (8 * n if kind == "local" else n, sc.dof)).astype(np.float32)).cuda()
        if kind == "local":
            a = a[be.check_configs(0, a).bool()][:n].contiguous()
            stp = torch.from_numpy(np.random.RandomState(11).uniform(-0.2, 0.2, tuple(a.shape)).astype(np.float32)).cuda()
            lo, hi = torch.from_numpy(lim[0].astype(np.float32)).cuda(), torch.from_numpy(lim[1].astype(np.float32)).cuda()
            b = torch.minimum(torch.maximum(a + stp, lo), hi).contiguous()
        else:
            b = torch.from_numpy(rng.uniform(lim[0], lim[1], (n, sc.dof)).astype(np.float32)).cuda()
        for _ in range(4):
            f, p = be.check_edges(0, a, b, kw["resolution"])
        torch.cuda.synchronize()
        print(name, a.shape[0], kind, "edges free", f.float().mean().item())
else:
    n = int(sys.argv[2])
    lim = SCENES["box_stacking"][0]().limits()
    c = torch.from_numpy(np.random.RandomState(5).uniform(lim[0], lim[1], (n, 24))).cuda()
    sl = [[6 * r, 6 * r + 6] for r in range(4)]
    if what == "knn":
        mode = sys.argv[3] if len(sys.argv) > 3 else "tensor"
        for _ in range(3):
            i, d = K.batch_knn(c, c, sl, "max_euclidean", 33, mode=mode)
        torch.cuda.synchronize()
        print("knn", n, mode, i[0, :5].tolist())
    else:
        _, d33 = K.batch_knn(c[:2048].contiguous(), c, sl, "max_euclidean", 33)
        r = float(d33[:, -1].median().item())
        for _ in range(3):
            off, idx = K.batch_radius(c, c, r, sl, "max_euclidean")
        torch.cuda.synchronize()
        print("radius", n, r, off[-1].item() / n)
