"""Edge-kernel throughput for planner-like edges of a given length (GPU box):
python scripts/bench_edges.py [scene] -- prints edges/s and interpolation points/s per length class."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.backend import SceneBackend
from multirobot_pathplanning_benchmark_b200.scenes import SCENES

names = sys.argv[1:] or ["box_rearrangement", "box_stacking", "mobile_wall_four", "2d_handover"]
be = SceneBackend(max_modes=4)
for name in names:
    mk, kw = SCENES[name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    be.set_mode(0, cs)
    lim = torch.from_numpy(sc.limits().astype(np.float32)).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    E = 1 << 18
    # free start points so that edges are not decided by their first tile
    q = lim[0] + (lim[1] - lim[0]) * torch.rand((8 * E, sc.dof), generator=g, device="cuda")
    q = q[be.check_configs(0, q).bool()][:E].contiguous()
    E = q.shape[0]
    for span in (0.05, 0.2, 0.6, 2.0):
        d = (torch.rand((E, sc.dof), generator=g, device="cuda") - 0.5) * 2 * span
        q2 = torch.minimum(torch.maximum(q + d, lim[0]), lim[1]).contiguous()
        N = torch.clamp((torch.max(torch.abs(q.double() - q2.double()), dim=1).values / kw["resolution"]).long() + 1, min=2)
        free, first = be.check_edges(0, q, q2, kw["resolution"])
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            be.check_edges(0, q, q2, kw["resolution"])
        b.record()
        b.synchronize()
        dt = a.elapsed_time(b) / 5 * 1e-3
        # points actually needed: all interior points of free edges, up to the first hit otherwise
        need = torch.where(free.bool(), N - 2, first.long() + 1).clamp(min=0).sum().item()
        print(f"{name:18s} span {span:4.2f}: mean N {N.float().mean().item():6.1f}  free {free.float().mean().item():.3f}  "
              f"{E / dt:.3e} edges/s  {need / dt:.3e} required points/s  ({(N - 2).clamp(min=0).sum().item() / dt:.3e} if all points counted)")
