"""Golden geometry derived from the reference's own robot model files (build container only).

Reads the `.g` files under /root/reference/src/multi_robot_multi_goal_planning/assets/models/rai with
multirobot_pathplanning_benchmark_b200.gfile, assembles the three arm / mobile-base scenes of BASELINE.json
from them, and stores what the collision path depends on -- joint layout, limits, home configuration, every
collision shape (kind, size, contact flag, link), the collidable pair list and the world pose of every shape
at fixed configurations -- in tests/golden/g_models.json.  tests/test_gfile.py checks scenes.py (the
transcription that ships, and that runs where the reference is absent) against it.

    python scripts/make_golden_gmodels.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multirobot_pathplanning_benchmark_b200 import scenes  # noqa: E402

MODELS = "/root/reference/src/multi_robot_multi_goal_planning/assets/models/rai"
MAKERS = {"box_rearrangement": scenes.make_box_rearrangement, "box_stacking": scenes.make_box_stacking,
          "mobile_wall_four": scenes.make_mobile_wall}


def summary(sc, n_cfg=4, seed=0):
    lim = sc.limits()
    rng = np.random.RandomState(seed)
    qs = np.vstack([sc.home()[None], rng.uniform(lim[0], lim[1], (n_cfg, sc.dof))])
    shapes = sc.collision_shapes()
    out = {"dof": sc.dof, "robots": sc.robots, "limits": lim.tolist(), "home": sc.home().tolist(),
           "shapes": {n: {"kind": sc.frames[n].shape.kind, "size": list(map(float, sc.frames[n].shape.size)),
                          "contact": sc.frames[n].contact, "link": sc.link_of(n)} for n in shapes},
           "pairs": sorted(sorted(p) for p in sc.collidable_pairs()), "configs": qs.tolist(), "poses": []}
    for q in qs:
        X = sc.fk(q)
        out["poses"].append({n: np.concatenate([X[n].t, X[n].R.ravel()]).round(12).tolist() for n in shapes})
    return out


if __name__ == "__main__":
    res = {"source": "reference .g files parsed by gfile.py: ur10/ur10_vacuum.g, ur10/ur10_two_finger.g, "
                     "mobile-manipulator-restricted.g (+ includes)",
           "scenes": {k: summary(mk(models_dir=MODELS)) for k, mk in MAKERS.items()}}
    path = os.path.join(ROOT, "tests", "golden", "g_models.json")
    with open(path, "w") as f:
        json.dump(res, f)
    print(path, os.path.getsize(path), "bytes")
