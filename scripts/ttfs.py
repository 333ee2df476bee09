"""Time-to-first-solution of the batch-native PRM (planner.py): B200 backend vs the CPU oracle backend
answering the very same batch calls.  usage: python scripts/ttfs.py [scene ...] [--cpu] [--seeds N]"""
import json, sys, time
import numpy as np
sys.path.insert(0, ".")
from multirobot_pathplanning_benchmark_b200.env import SceneModel, CudaDevice
from multirobot_pathplanning_benchmark_b200.planner import BatchedPRM, SeqTask
from multirobot_pathplanning_benchmark_b200.scenes import SCENES
from multirobot_pathplanning_benchmark_b200 import problems as P


def handover_tasks(start):
    return [SeqTask(["a1"], np.array([0.0, 0.77, 0.0]), ("a1", "obj1")),
            SeqTask(["a1", "a2"], np.array([-1.2, 1.37, 0.0, -1.2, 0.58, 0.0]), ("a2", "obj1")),
            SeqTask(["a1"], np.array([0.5, -1.13, 0.0]), ("a1", "obj2")),
            SeqTask(["a1"], np.array([1.3, 1.57, 0.0]), ("table", "obj2")),
            SeqTask(["a2"], np.array([1.22, 0.4, np.pi / 2]), ("table", "obj1")),
            SeqTask(["a1", "a2"], start.copy(), None)]


def goto_tasks(model, sc, seed):
    """every robot moves to a sampled collision-free goal (the others are free), then all return home"""
    rng = np.random.RandomState(seed)
    lim = sc.limits()
    slot = model.slot_for(())
    while True:
        cand = rng.uniform(lim[0], lim[1], (4096, sc.dof)).astype(np.float32)
        ok = model.check_configs(slot, cand)
        ok = ok.cpu().numpy() if hasattr(ok, "cpu") else np.asarray(ok)
        if ok.any():
            goal = cand[int(np.argmax(ok))].astype(np.float64)
            break
    sl = sc.robot_slices()
    tasks = [SeqTask([r], goal[sl[r][0]:sl[r][1]].copy(), None) for r in sc.robots]
    tasks.append(SeqTask(list(sc.robots), sc.home(), None))
    return tasks


def run(scene_name, backend, seed, n0, t0, max_time, goto=False, n_moves=4):
    mk, kw = SCENES[scene_name]
    sc = mk()
    if backend == "b200":
        import torch
        from multirobot_pathplanning_benchmark_b200 import knn as K
        model = SceneModel(sc, kw["tol"], kw["resolution"])
        knn = lambda q, c, sl, metric, k: K.batch_knn(torch.from_numpy(q).cuda(), torch.from_numpy(c).cuda(), sl, metric, k, return_dist=False)
    else:
        from oracle import oracle_abstract as OA, oracle_scene as O
        class Dev:  # the CPU backend: fp64 oracle on all host threads, answering the same batch calls
            def __init__(self): self.cs = {}
            def set_mode(self, slot, cs): self.cs[slot] = cs
            def check_configs(self, slot, q, tol=None):
                return O.check_configs(self.cs[slot].blob64, np.asarray(q, np.float64), nthreads=O.max_threads())[0]
            def check_edges(self, slot, q1, q2, resolution, **kw):
                f, p, _ = O.check_edges(self.cs[slot].blob64, np.asarray(q1, np.float64), np.asarray(q2, np.float64), resolution, nthreads=O.max_threads())
                return f, p
        model = SceneModel(sc, kw["tol"], kw["resolution"], device=Dev())
        def knn(q, c, sl, metric, k):
            out = np.full((len(q), k), -1, np.int64)
            sl = np.asarray(sl)
            for i, row in enumerate(q):
                idx = OA.knn_indices(OA.batch_config_dist(row, c, sl, metric), k)
                out[i, :len(idx)] = idx
            return out
    if scene_name == "2d_handover":
        tasks = handover_tasks(sc.home())
    elif scene_name in P.PROBLEMS and not goto:
        # pick / place sequence with held objects (keyframes by numerical IK, checked by this backend)
        tasks = [SeqTask(list(t.robots), t.goal, t.frames) for t in P.manipulation_tasks(scene_name, model, n_moves=n_moves, seed=0)]
    else:
        tasks = goto_tasks(model, sc, 100 + seed)
    prm = BatchedPRM(model, tasks, sc.home(), knn, seed=seed, samples_per_mode=n0, transitions_per_mode=t0)
    res = prm.plan(max_time=max_time)
    return {"solved": res.path is not None, "time_s": res.time_s, "cost": res.cost, **res.stats}


if __name__ == "__main__":
    scenes = [a for a in sys.argv[1:] if a in SCENES] or ["2d_handover"]
    cpu = "--cpu" in sys.argv
    seeds = int(sys.argv[sys.argv.index("--seeds") + 1]) if "--seeds" in sys.argv else 3
    for s in scenes:
        n0, t0 = (500, 60) if s == "2d_handover" else (4000, 400)
        run(s, "b200", 99, 200, 30, 30)  # warm-up: CUDA context, kernel load
        for backend in ["b200"] + (["cpu"] if cpu else []):
            rs = [run(s, backend, seed, n0, t0, 120) for seed in range(seeds)]
            print(json.dumps({"scene": s, "backend": backend, "median_time_s": float(np.median([r["time_s"] for r in rs])), "runs": rs}))
