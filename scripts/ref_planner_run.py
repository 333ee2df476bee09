#!/usr/bin/env python
"""Time-to-first-solution of the reference's own planners on a b200 environment (BASELINE metric 3).

    python scripts/ref_planner_run.py ENV PLANNER --device cuda|cpu --seeds 1,2,3 [--max-time S] [--optimize]
                                      [--no-speculation] [--warmup]

ENV      b200 environment class: box_stacking | box_rearrangement | 2d_handover | dep_mobile_wall_four | abstract_test,
         or `ref:abstract.test` for the reference's own numpy environment (the true reference CPU path of config 1)
PLANNER  composite_prm | rrt_star | birrt_star | aitstar | eitstar      (P/scripts/run_planner.py:72-85)
--device cuda: the product path (libmrb200.so);  cpu: the fp64 oracle device, one query per call, single threaded --
         the stand-in for "the reference CPU backend" where rai cannot run (bench.py's CPU arm; test infrastructure)

Prints one JSON line per seed: info["times"][0] (ttfs_s), costs, wall time of plan(), time spent inside the device and the
number of device round trips against the number of planner queries."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ENVS = {"box_stacking": "b200_box_stacking", "box_rearrangement": "b200_box_rearrangement", "2d_handover": "b200_two_dim_handover",
        "dep_mobile_wall_four": "b200_dep_mobile_wall_four", "abstract_test": "b200_abstract_test"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("env")
    ap.add_argument("planner")
    ap.add_argument("--device", default="cuda", choices=["cuda", "cpu"])
    ap.add_argument("--seeds", default="1")
    ap.add_argument("--max-time", type=float, default=300.0)
    ap.add_argument("--optimize", action="store_true")
    ap.add_argument("--no-speculation", action="store_true")
    ap.add_argument("--no-mode-validation", action="store_true")
    ap.add_argument("--warmup", action="store_true", help="one short untimed run first (CUDA context, numba compilation)")
    args = ap.parse_args()

    from multirobot_pathplanning_benchmark_b200 import env as E
    from multirobot_pathplanning_benchmark_b200 import refplanners as RP
    if not E.HAVE_REFERENCE:
        print(json.dumps({"error": "reference package not importable (baseline/_ref missing?)"}))
        return 1

    kw = {}
    if args.no_mode_validation or (args.env.endswith("abstract_test") or args.env.startswith("ref:abstract")) and args.planner != "composite_prm":
        kw["with_mode_validation"] = False   # abstract envs have no per-robot rule

    def build():
        if args.env.startswith("ref:"):
            from multi_robot_multi_goal_planning.problems import get_env_by_name
            return get_env_by_name(args.env[4:]), None
        cls = getattr(E, ENVS[args.env])
        if args.env == "abstract_test":
            if args.device == "cpu":
                from oracle.oracle_device import OracleAbstractDevice
                dev = OracleAbstractDevice()
            else:
                tmp = cls(speculate=not args.no_speculation)
                dev = tmp.device
            meter = RP.DeviceMeter(dev)
            return cls(device=meter, speculate=not args.no_speculation), meter
        if args.device == "cpu":
            from oracle.oracle_device import OracleSceneDevice
            dev = OracleSceneDevice(nthreads=1)
        else:
            dev = E.CudaDevice()
        meter = RP.DeviceMeter(dev)
        return cls(device=meter, speculate=not args.no_speculation), meter

    if args.warmup:
        env, _ = build()
        RP.run_planner(env, args.planner, 12345, min(args.max_time, 20.0), optimize=False, **kw)

    for seed in [int(s) for s in args.seeds.split(",") if s]:
        t = time.perf_counter()
        env, meter = build()
        t_build = time.perf_counter() - t
        base = meter.snapshot() if meter else None
        res = RP.run_planner(env, args.planner, seed, args.max_time, optimize=args.optimize, **kw)
        path = res.pop("_path")
        res.update({"env": args.env, "device": args.device if not args.env.startswith("ref:") else "reference-numpy",
                    "env_build_s": t_build, "valid_plan": bool(path is not None and env.is_valid_plan(path))})
        if meter:
            snap = meter.snapshot()
            res["backend_s"] = snap["seconds"] - base["seconds"]
            res["device_calls"] = {k: v - base["calls"].get(k, 0) for k, v in snap["calls"].items()}
            res["device_items"] = {k: v - base["items"].get(k, 0) for k, v in snap["items"].items()}
            res["device_round_trips"] = int(sum(res["device_calls"].values()))
        sc = getattr(env, "spec_cache", None)
        if sc is not None:
            res["speculation"] = dict(sc.stats)
            res["planner_queries"] = int(sum(sc.stats.get(k, 0) for k in ("config_queries", "edge_queries", "robot_queries")))
        print(json.dumps(res), flush=True)
    return 0


if __name__ == "__main__":
    sys.exit(main())
