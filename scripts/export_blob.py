"""Write a compiled scene blob (+ joint limits) for native/bench_main: python scripts/export_blob.py box_rearrangement out.blob"""
import sys
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200 import scene as S
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
name, out = sys.argv[1], sys.argv[2]
mk, kw = SCENES[name]
sc = mk()
cs = S.compile_blob(sc, kw["tol"])
with open(out, "wb") as f:
    f.write(cs.blob32.tobytes())
    f.write(sc.limits().astype(np.float32).tobytes())
print(out, len(cs.blob32) * 4, "bytes, dof", sc.dof)
