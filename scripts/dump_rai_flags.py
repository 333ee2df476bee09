#!/usr/bin/env python
"""Dumps rai's own collision answers on the seeded batches the GPU parity tests use -- the missing pin of the scene
oracle ("parity UNPINNED against rai").  Run this on ANY machine that has the reference installed together
with its rai backend (`pip install robotic>=0.2.2,<0.3.0`, reference pyproject.toml:27); it needs no GPU and nothing of
this repository except this file:

    python scripts/dump_rai_flags.py [--out tests/golden] [--batch 60000] [--edges 1500]

For each of the four BASELINE scenes (rai.2d_handover, rai.box_rearrangement, rai.box_stacking,
rai.dep_mobile_wall_four) it draws exactly the inputs of tests/test_gpu_scene.py (`np.random.seed(s);
np.random.uniform(limits[0], limits[1], (B, D))`, seeds 0 / 10 / 11, rounded to fp32 like the device inputs) and stores, in
`rai_flags_<scene>.npz`:

    q            [B, D] float32   the configurations (start mode)
    free         [B] bool         env.is_collision_free_np(q, start mode)        (P/problems/rai_base_env.py:480-513)
    pen          [B] float64      C.getCollisionsTotalPenetration() after setJointState(q)   (the margin of the decision)
    e_q1, e_q2   [E, D] float32   edge endpoints (seeds 10 / 11, every other edge local like the tests)
    e_free       [E] bool         env.is_edge_collision_free(q1, q2, start mode)  (:618-676)
    limits, tol, resolution, joint_names, robotic_version

tests/test_rai_golden.py consumes these files when they exist: flags of the fp64 oracle AND of the CUDA kernels must agree
with rai on every sample whose rai penetration is more than 1e-5 away from the tolerance -- cylinders excepted where the
oracle models them as capsules (reported separately).  The sampling mirrors the reference's own benchmark loop
(P/scripts/show_problems.py:167-185, P/scripts/compare_rai_vamp_coll_checking.py:163-168)."""
import argparse
import os
import sys

import numpy as np

SCENES = {"2d_handover": "rai.2d_handover", "box_rearrangement": "rai.box_rearrangement", "box_stacking": "rai.box_stacking",
          "mobile_wall_four": "rai.dep_mobile_wall_four"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
    ap.add_argument("--batch", type=int, default=60_000)
    ap.add_argument("--edges", type=int, default=1500)
    ap.add_argument("--reference-src", default=None, help="path of the reference's src/ directory if it is not installed")
    args = ap.parse_args()
    if args.reference_src:
        sys.path.insert(0, args.reference_src)
    try:
        import robotic as ry
    except ImportError:
        raise SystemExit("this script needs the rai backend of the reference: pip install 'robotic>=0.2.2,<0.3.0'")
    from multi_robot_multi_goal_planning.problems import get_env_by_name
    os.makedirs(args.out, exist_ok=True)
    for name, env_name in SCENES.items():
        env = get_env_by_name(env_name)
        lim = np.asarray(env.limits, np.float64)
        D = lim.shape[1]
        m0 = env.get_start_mode()
        np.random.seed(0)
        q = np.random.uniform(lim[0], lim[1], (args.batch, D)).astype(np.float32)
        free = np.zeros(len(q), bool)
        pen = np.zeros(len(q))
        env.set_to_mode(m0)
        for i, row in enumerate(q):
            row64 = row.astype(np.float64)
            free[i] = env.is_collision_free_np(row64, m0)
            env.C.setJointState(row64)
            env.C.computeCollisions()
            pen[i] = env.C.getCollisionsTotalPenetration()
        np.random.seed(10)
        q1 = np.random.uniform(lim[0], lim[1], (args.edges, D)).astype(np.float32)
        np.random.seed(11)
        q2 = np.random.uniform(lim[0], lim[1], (args.edges, D)).astype(np.float32)
        q2[::2] = q1[::2] + np.random.default_rng(1).uniform(-0.15, 0.15, q1[::2].shape).astype(np.float32)
        e_free = np.zeros(len(q1), bool)
        for i in range(len(q1)):
            a = env.start_pos.from_flat(q1[i].astype(np.float64))
            b = env.start_pos.from_flat(q2[i].astype(np.float64))
            e_free[i] = env.is_edge_collision_free(a, b, m0)
        joint_names = [str(n) for n in env.C.getJointNames()] if hasattr(env.C, "getJointNames") else []
        np.savez_compressed(os.path.join(args.out, f"rai_flags_{name}.npz"), q=q, free=free, pen=pen, e_q1=q1, e_q2=q2, e_free=e_free,
                            limits=lim, tol=float(env.collision_tolerance), resolution=float(env.collision_resolution),
                            joint_names=np.array(joint_names), robotic_version=str(getattr(ry, "__version__", "unknown")))
        print(f"{name}: {free.mean():.3f} of {len(q)} configs free, {e_free.mean():.3f} of {len(q1)} edges free")


if __name__ == "__main__":
    main()
