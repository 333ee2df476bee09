"""e2e (pinned host -> H2D -> kernel -> D2H) rate of check_configs_host by chunk size (GPU box)."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.backend import SceneBackend, check_configs_host
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
mk, kw = SCENES["box_rearrangement"]
sc = mk(); cs = S.compile_blob(sc, kw["tol"])
be = SceneBackend(max_modes=2); be.set_mode(0, cs)
B = 4_194_304
lim = sc.limits()
q = torch.from_numpy(np.random.RandomState(0).uniform(lim[0], lim[1], (B, sc.dof)).astype(np.float32)).pin_memory()
out = torch.empty(B, dtype=torch.uint8).pin_memory()
for chunk in (1 << 19, 1 << 18, 1 << 17, 1 << 16, 1 << 15, 1 << 14):
    st = {}
    for _ in range(2): check_configs_host(be, 0, q, out, chunk=chunk, state=st)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): check_configs_host(be, 0, q, out, chunk=chunk, state=st)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"chunk {chunk:7d}: {ms:.3f} ms/step  {B / ms * 1e-6:.4f}e9 configs/s  H2D {B * sc.dof * 4 / ms * 1e-6:.1f} GB/s")
