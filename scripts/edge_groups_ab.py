"""GPU box: the record-group cull of coherent edge tiles against the plain broadphase.  Run twice (second time with
MRB200_NO_EDGE_GROUPS=1); each run writes its flags / first positions to gpurun_out/edge_groups_<tag>.npz and prints rates;
`python scripts/edge_groups_ab.py compare` checks that the two runs agree bit for bit.
usage: python scripts/edge_groups_ab.py [run TAG | compare]"""
import os, sys
import numpy as np
sys.path.insert(0, ".")
if len(sys.argv) > 1 and sys.argv[1] == "compare":
    a, b = np.load("gpurun_out/edge_groups_on.npz"), np.load("gpurun_out/edge_groups_off.npz")
    bad = [k for k in a.files if not np.array_equal(a[k], b[k])]
    print("grouped == plain on every edge of every scene:", not bad, bad)
    sys.exit(1 if bad else 0)
import torch
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
tag = sys.argv[2] if len(sys.argv) > 2 else "on"
out = {}
for name in ("box_rearrangement", "box_stacking", "mobile_wall_four", "2d_handover"):
    mk, kw = SCENES[name]
    sc = mk(); cs = S.compile_blob(sc, kw["tol"])
    be = SceneBackend(max_modes=2); be.set_mode(0, cs)
    lim = sc.limits(); rng = np.random.RandomState(3)
    pool = torch.from_numpy(rng.uniform(lim[0], lim[1], (1 << 20, sc.dof)).astype(np.float32)).cuda()
    pool = pool[be.check_configs(0, pool).bool()][:131072].contiguous()
    lo, hi = torch.from_numpy(lim[0].astype(np.float32)).cuda(), torch.from_numpy(lim[1].astype(np.float32)).cuda()
    for span in (0.2, 0.05):
        stp = torch.from_numpy(np.random.RandomState(11).uniform(-span, span, tuple(pool.shape)).astype(np.float32)).cuda()
        q2 = torch.minimum(torch.maximum(pool + stp, lo), hi).contiguous()
        for _ in range(3):
            f, p = be.check_edges(0, pool, q2, kw["resolution"])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            f, p = be.check_edges(0, pool, q2, kw["resolution"])
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / 5
        print(f"{tag:4s} {name:18s} local +-{span}: {len(pool)} edges, {ms:.3f} ms, {len(pool) / ms * 1e3:.4g} edges/s, free {f.float().mean().item():.3f}", flush=True)
        out[f"{name}_{span}_f"] = f.cpu().numpy(); out[f"{name}_{span}_p"] = p.cpu().numpy()
    u1 = torch.from_numpy(rng.uniform(lim[0], lim[1], (16384, sc.dof)).astype(np.float32)).cuda()
    u2 = torch.from_numpy(rng.uniform(lim[0], lim[1], (16384, sc.dof)).astype(np.float32)).cuda()
    f, p = be.check_edges(0, u1, u2, kw["resolution"])
    out[f"{name}_uni_f"] = f.cpu().numpy(); out[f"{name}_uni_p"] = p.cpu().numpy()
os.makedirs("gpurun_out", exist_ok=True)
np.savez(f"gpurun_out/edge_groups_{tag}.npz", **out)
