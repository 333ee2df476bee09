#!/usr/bin/env python
"""Benchmark of the collision / proximity hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input: `check_configs` over the workload's
configuration batch (uniform in the joint limits, exactly like the reference's sampler,
P/problems/planning_env.py:1697-1708).  Default workload: the dual-arm scene of BASELINE.json (box_rearrangement,
4 194 304 configurations), the configuration north_star's >= 1e9 checks/s target is quoted on.  ONE JSON line on rank 0.

  value         whole-job configuration checks/s with the batch resident in HBM
  e2e           the same through the host-buffer API: pinned host configs -> H2D -> kernel -> D2H flags
  roofline      FP32-SIMT roofline of the FK + narrowphase kernel on ALGORITHMIC flop + what the pipes
                really execute (`executed`, `issue_active`: ncu capture of the timed build, profiles/traffic.json)
  edges         BASELINE metric 2: edge checks/s (uniform and planner-like local edges), W_edge roofline, CPU edges/s
  knn           BASELINE config 4: 100k x 100k k-NN (tcgen05 path == exact path asserted) and r-disc
  planners      BASELINE metric 3: the reference's own PRM / EIT* (baseline/_ref) on b200.box_stacking, CUDA device vs the
                CPU oracle device, same seeds, median time to first solution; abstract.test with the five planners
  modes         held-object modes, all-free batches next to the headline
  scenes        the other named scenes, same protocol, fewer steps
  strong_sweep  (N > 1) BASELINE config 5: the 64M mobile_wall_four sweep sharded over the ranks
  cpu_baseline  the fp64 CPU oracle (a port: rai itself cannot be installed) on all host cores

--impl reference times the reference's CPU path for the same metric / config: rai (`robotic`) is an un-vendored wheel
that cannot run here, so this arm runs the fp64 oracle port with every host core (thread count passed explicitly:
torchrun's OMP_NUM_THREADS=1 does not apply).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (scene, configs per GPU, edges per GPU)
    "box_rearrangement_4M": ("box_rearrangement", 4_194_304, 16_384),
    "2d_handover_1M": ("2d_handover", 1_048_576, 100_000),
    "box_stacking_1M": ("box_stacking", 1_048_576, 8_192),
    "mobile_wall_four_8M": ("mobile_wall_four", 8_388_608, 16_384),
    # BASELINE config 5: one 64M-configuration sweep, SHARDED over the ranks (strong scaling)
    "mobile_wall_four_64M": ("mobile_wall_four", 67_108_864, 16_384),
}
STRONG = {"mobile_wall_four_64M"}  # total work fixed, split across ranks; every other workload is per GPU (weak)
DEFAULT = "box_rearrangement_4M"
METRIC = "config collision checks/sec"
SM_COUNT, FP32_LANES = 148, 128


def uniform_configs(lim, B, seed, chunk=4_194_304):
    """uniform in the joint limits like the reference's sampler (np.random.uniform, fp64) rounded to fp32;
    large batches are drawn chunk by chunk from the same stream"""
    rng = np.random.RandomState(seed)
    if B <= chunk:
        return rng.uniform(lim[0], lim[1], (B, lim.shape[1])).astype(np.float32)
    out = np.empty((B, lim.shape[1]), np.float32)
    for i in range(0, B, chunk):
        n = min(chunk, B - i)
        out[i:i + n] = rng.uniform(lim[0], lim[1], (n, lim.shape[1]))
    return out


def config_block(args, scene_name, B, world, D, cs, free_frac=None, strong=False):
    """the `config` object: identical keys for both arms (--impl b200 / reference)"""
    cfg = {"workload": args.workload, "scene": scene_name, "configs_per_gpu": B, "configs_total": world * B, "dof": D,
           "collidable_pairs": int(sum(cs.pair_counts)), "tolerance": cs.tol,
           "inputs": "uniform in joint limits (np.random.uniform), fp32",
           "l2": f"input batch {B * D * 4 / 1e6:.0f} MB > 126 MB L2, streamed once per step",
           "exchange": "all_gather of flag bytes (NCCL)" if world > 1 else "none"}
    return cfg


def capture(section, key, expect=None):
    """a per-launch figure from the committed `ncu --set full` capture of the timed build (profiles/traffic.json, written
    by scripts/ncu_summary.py): None if absent or taken at another problem size"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(section)
        if not t:
            return None
        if expect is not None and any(int(t.get(k, -1)) != int(v) for k, v in expect.items()):
            return None
        return t.get(key)
    except Exception:
        return None


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        sm.sort()
        busy = [s for s in sm if s > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": busy[len(busy) // 2] if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
def host_threads():
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_configs_rate(cs, lim, budget_s, B_full):
    """fp64 oracle port on all host cores: the full batch if it fits the time budget, else a bounded sample"""
    from oracle import oracle_scene as O
    nthreads = host_threads()
    probe = uniform_configs(lim, 20_000, 123).astype(np.float64)
    O.check_configs(cs.blob64, probe[:2000], nthreads=nthreads)   # thread pool start-up
    t = time.perf_counter()
    O.check_configs(cs.blob64, probe, nthreads=nthreads)
    rate = len(probe) / (time.perf_counter() - t)
    n = int(min(max(rate * budget_s, 50_000), B_full))
    return rate, n, nthreads


def cpu_baseline(cs, lim, B_full, kind_note, budget_s=12.0):
    from oracle import oracle_scene as O
    rate, n, nthreads = cpu_configs_rate(cs, lim, budget_s, B_full)
    q = uniform_configs(lim, n, 0).astype(np.float64)
    t = time.perf_counter()
    O.check_configs(cs.blob64, q, nthreads=nthreads)
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": "configs/s", "cores": nthreads, "kind": "port",
            "sample": f"one pass over {n} of {B_full} uniform configs of the same workload, fp64 C oracle ({kind_note}), "
                      f"{nthreads} OpenMP threads (set explicitly), {dt:.1f} s"}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of this path.  rai (`robotic`) is an un-vendored third-party
    wheel that cannot be installed offline, so this is the fp64 oracle port with every host core; one step = one pass
    over the workload's batch (the full batch whenever the whole run fits --ref-seconds, else a bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from multirobot_pathplanning_benchmark_b200 import scene as S
    from multirobot_pathplanning_benchmark_b200.scenes import SCENES
    from oracle import oracle_scene as O
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    scene_name, B, _ = WORKLOADS[args.workload]
    if args.workload in STRONG:
        B = B // max(world, 1)
    mk, kw = SCENES[scene_name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    lim = sc.limits()
    total_steps = args.steps + args.warmup
    rate, n, nthreads = cpu_configs_rate(cs, lim, args.ref_seconds / max(total_steps, 1), B)
    q = uniform_configs(lim, n, 1000).astype(np.float64)
    for _ in range(args.warmup):
        O.check_configs(cs.blob64, q, nthreads=nthreads)
    t = time.perf_counter()
    for _ in range(args.steps):
        O.check_configs(cs.blob64, q, nthreads=nthreads)
    dt = time.perf_counter() - t
    value = n * args.steps / dt
    sample = (f"{n} of {B} uniform configs per step ({'the full batch' if n == B else 'bounded sample'}), fp64 C oracle port of "
              f"the rai query, {nthreads} OpenMP threads (set explicitly; OMP_NUM_THREADS={os.environ.get('OMP_NUM_THREADS', 'unset')} ignored)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "configs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args, scene_name, B, world, sc.dof, cs),
        "cpu_baseline": {"value": value, "unit": "configs/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
def planners_block(args, log):
    """BASELINE metric 3 with the reference's OWN planners (baseline/_ref): same environment classes, same seeds, the
    CUDA device against the CPU oracle device (one query per call, single threaded, no speculation -- the stand-in for
    rai).  Each arm runs in its own processes (scripts/ref_planner_run.py); CPU seeds run concurrently, one core each."""
    script = os.path.join(ROOT, "scripts", "ref_planner_run.py")
    gseeds = ",".join(str(s) for s in range(1, 1 + args.planner_seeds))
    cseeds = [str(s) for s in range(1, 1 + args.planner_cpu_seeds)]

    def parse(out):
        rows = []
        for line in out.splitlines():
            line = line.strip()
            if line.startswith("{"):
                try:
                    rows.append(json.loads(line))
                except ValueError:
                    pass
        return rows

    def summarise(rows):
        ok = [r for r in rows if r.get("ttfs_s") is not None]
        if not ok:
            return {"runs": rows, "solved": 0}
        med = lambda k: float(np.median([r[k] for r in ok if r.get(k) is not None])) if any(r.get(k) is not None for r in ok) else None
        return {"solved": len(ok), "of": len(rows), "median_ttfs_s": med("ttfs_s"), "median_wall_s": med("wall_s"),
                "median_backend_s": med("backend_s"), "median_device_round_trips": med("device_round_trips"),
                "median_first_cost": med("first_cost"), "valid_plans": all(r.get("valid_plan") for r in ok),
                "runs": [{k: r.get(k) for k in ("seed", "ttfs_s", "wall_s", "backend_s", "first_cost", "device_round_trips",
                                               "device_calls", "speculation")} for r in rows]}

    res = {"note": "reference planners unmodified (baseline/_ref), optimize=False, info['times'][0]; cpu arm = fp64 oracle "
                   "device, single thread, one query per call, no speculation (kind: port, rai cannot run)"}
    jobs = [("box_stacking", "composite_prm", args.planner_max_time), ("box_stacking", "eitstar", args.planner_max_time)]
    for env_name, planner, max_time in jobs:
        key = f"{env_name}/{planner}"
        entry = {}
        try:
            cpu_procs = []
            if not args.no_cpu:
                for s in cseeds:
                    cpu_procs.append(subprocess.Popen(
                        [sys.executable, script, env_name, planner, "--device", "cpu", "--seeds", s, "--max-time", str(max_time),
                         "--no-speculation"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
                        env={**os.environ, "OMP_NUM_THREADS": "1", "NUMBA_NUM_THREADS": "1"}))
            g = subprocess.run([sys.executable, script, env_name, planner, "--device", "cuda", "--seeds", gseeds, "--max-time",
                                str(max_time), "--warmup"], capture_output=True, text=True, timeout=max_time * (args.planner_seeds + 2))
            entry["cuda"] = summarise(parse(g.stdout))
            if g.returncode:
                entry["cuda"]["stderr"] = g.stderr[-400:]
            if cpu_procs:
                rows = []
                for p in cpu_procs:
                    out, _ = p.communicate(timeout=max_time + 120)
                    rows += parse(out)
                entry["cpu"] = summarise(rows)
                a, b = entry["cuda"].get("median_ttfs_s"), entry["cpu"].get("median_ttfs_s")
                if a and b:
                    entry["ttfs_ratio_cpu_over_cuda"] = b / a
                a, b = entry["cuda"].get("median_backend_s"), entry["cpu"].get("median_backend_s")
                if a and b:
                    entry["backend_time_ratio_cpu_over_cuda"] = b / a
                # same seeds -> the same first solution unless a query fell inside the 1e-5 margin
                both = {r["seed"]: r for r in entry["cpu"]["runs"]}
                same = [abs(r["first_cost"] - both[r["seed"]]["first_cost"]) < 1e-9 for r in entry["cuda"]["runs"]
                        if r["seed"] in both and r.get("first_cost") is not None and both[r["seed"]].get("first_cost") is not None]
                entry["same_first_solution"] = f"{sum(same)} of {len(same)} common seeds"
        except Exception as e:  # the headline line must survive a planner hiccup
            entry["error"] = repr(e)
        res[key] = entry
        log(f"planners {key}: {json.dumps({k: v for k, v in entry.items() if k not in ('cuda', 'cpu')})}")
    # BASELINE config 1: abstract.test with the five planners (run_planner.py semantics, seed 1), the reference's own numpy
    # environment against the CUDA-backed one (bit-identical flags -> identical plans)
    ab = {}
    for planner in ("composite_prm", "rrt_star", "birrt_star", "aitstar", "eitstar"):
        try:
            procs = {}
            if not args.no_cpu:
                procs["reference_numpy"] = subprocess.Popen([sys.executable, script, "ref:abstract.test", planner, "--seeds", "1", "--max-time", "10"],
                                                            stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            g = subprocess.run([sys.executable, script, "abstract_test", planner, "--device", "cuda", "--seeds", "1", "--max-time", "10",
                                "--warmup"], capture_output=True, text=True, timeout=120)
            rows = parse(g.stdout)
            e = {"cuda": {k: rows[0].get(k) for k in ("ttfs_s", "first_cost", "wall_s", "backend_s", "device_round_trips")} if rows else None}
            for name, p in procs.items():
                out, _ = p.communicate(timeout=120)
                r = parse(out)
                e[name] = {k: r[0].get(k) for k in ("ttfs_s", "first_cost", "wall_s")} if r else None
            if e.get("cuda") and e.get("reference_numpy"):
                e["identical_first_cost"] = e["cuda"]["first_cost"] == e["reference_numpy"]["first_cost"]
            ab[planner] = e
        except Exception as ex:
            ab[planner] = {"error": repr(ex)}
    res["abstract.test (seed 1, max_time 10, first solution)"] = ab
    return res


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=list(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary blocks (edges, knn, modes, scenes, planners)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU legs")
    ap.add_argument("--no-planners", action="store_true", help="skip the reference-planner block")
    ap.add_argument("--ref-seconds", type=float, default=90.0, help="--impl reference: CPU time budget of the whole run")
    ap.add_argument("--planner-seeds", type=int, default=5)
    ap.add_argument("--planner-cpu-seeds", type=int, default=5)
    ap.add_argument("--planner-max-time", type=float, default=180.0)
    ap.add_argument("--verbose", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    if world_env > 1:  # secondary blocks and the CPU leg are N = 1 only
        args.no_extra = args.no_cpu = True

    if args.impl == "reference":
        return run_reference(args)

    def log(msg):
        if args.verbose:
            print(f"[bench {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)

    import torch
    import torch.distributed as dist

    from multirobot_pathplanning_benchmark_b200 import scene as S
    from multirobot_pathplanning_benchmark_b200.backend import (SceneBackend, check_configs_host, fp32_fma_peak_tflops,
                                                                launch_count)
    from multirobot_pathplanning_benchmark_b200.scenes import SCENES

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = world_env
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # one rank per GPU: run on the CPUs (and, by first touch, allocate the pinned staging memory) of the GPU's NUMA node
    from multirobot_pathplanning_benchmark_b200.dist import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local_rank) if world > 1 else {"bound": False, "note": "single rank: not bound"}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / reps * 1e-3

    scene_name, B, E = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    if strong:  # this rank's shard of the fixed-size sweep (contiguous row block, dist.shard_range)
        from multirobot_pathplanning_benchmark_b200.dist import shard_range
        lo, hi = shard_range(B, rank, world)
        B = hi - lo
    mk, kw = SCENES[scene_name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    lim = sc.limits()
    be = SceneBackend(max_modes=8, device=dev)
    be.set_mode(0, cs)
    D = sc.dof

    # every rank owns its own shard of the batch (weak scaling: B configurations per GPU)
    q_host = torch.from_numpy(uniform_configs(lim, B, 1000 + rank)).pin_memory()
    q_dev = q_host.to(dev)
    flags = torch.empty(B, dtype=torch.uint8, device=dev)
    gathered = torch.empty(world * B, dtype=torch.uint8, device=dev) if world > 1 else None

    def step():
        be.check_configs(0, q_dev, out=flags)
        if world > 1:  # the path's only exchange: the final flag gather (NCCL over NVLink)
            dist.all_gather_into_tensor(gathered, flags)

    for _ in range(args.warmup):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    l0 = launch_count()
    kernel_events = []
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t0.record()
    for _ in range(args.steps):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        be.check_configs(0, q_dev, out=flags)
        k1.record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, flags)
        kernel_events.append((k0, k1))
    t1.record()
    barrier()
    launches = launch_count() - l0
    total_ms = max_over_ranks(t0.elapsed_time(t1))
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kernel_events]))
    value = world * B * args.steps / (total_ms * 1e-3)
    free_frac = float(flags.float().mean().item())
    log(f"resident: {value:.4g} configs/s, {total_ms / args.steps:.3f} ms/step")

    # ---- end to end through the host-buffer API (pinned host in, pinned host out) ----
    out_host = torch.empty(B, dtype=torch.uint8).pin_memory()
    state = {}
    for _ in range(2):
        check_configs_host(be, 0, q_host, out_host, state=state)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(e2e_steps):
        check_configs_host(be, 0, q_host, out_host, state=state)   # returns after the flags have landed in out_host
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    # the host link alone: the same chunks copied H2D without any kernel (is the end-to-end rate link bound?)
    h0 = time.perf_counter()
    for _ in range(3):
        q_dev.copy_(q_host, non_blocking=True)
    torch.cuda.synchronize()
    h2d_gbs = 3 * B * D * 4 / (time.perf_counter() - h0) / 1e9
    h2d_gbs_min = -max_over_ranks(-h2d_gbs)
    clk = clocks.stop() if rank == 0 else None   # sampled over both timed regions (resident and end-to-end)
    assert torch.equal(out_host.to(dev), flags), "host-API flags differ from the device-resident run"
    log(f"e2e: {e2e_value:.4g} configs/s; H2D alone {h2d_gbs:.1f} GB/s")

    # ---- N > 1: BASELINE config 5, the 64M mobile_wall_four sweep sharded over the ranks (strong scaling) ----
    strong_sweep = None
    if world > 1 and not strong:
        from multirobot_pathplanning_benchmark_b200.dist import shard_range
        del q_dev, flags, gathered
        torch.cuda.empty_cache()
        mk5, kw5 = SCENES["mobile_wall_four"]
        sc5 = mk5()
        cs5 = S.compile_blob(sc5, kw5["tol"])
        be.set_mode(2, cs5)
        B5 = 67_108_864
        lo, hi = shard_range(B5, rank, world)
        n5 = hi - lo
        lim5 = torch.from_numpy(sc5.limits().astype(np.float32)).to(dev)
        gen = torch.Generator(device=dev).manual_seed(77 + rank)
        q5 = torch.empty((n5, sc5.dof), dtype=torch.float32, device=dev)
        for i in range(0, n5, 4_194_304):   # uniform in the limits, drawn on the device chunk by chunk
            j = min(n5, i + 4_194_304)
            q5[i:j] = lim5[0] + (lim5[1] - lim5[0]) * torch.rand((j - i, sc5.dof), generator=gen, device=dev)
        f5 = torch.empty(n5, dtype=torch.uint8, device=dev)
        sizes = [shard_range(B5, r, world) for r in range(world)]
        g5 = torch.empty(max(b - a for a, b in sizes) * world, dtype=torch.uint8, device=dev) if all(b - a == n5 for a, b in sizes) else None

        def sweep():
            be.check_configs(2, q5, out=f5)
            if g5 is not None:
                dist.all_gather_into_tensor(g5, f5)
        for _ in range(2):
            sweep()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(5):
            sweep()
        s1.record()
        barrier()
        sweep_ms = max_over_ranks(s0.elapsed_time(s1)) / 5
        strong_sweep = {"workload": "mobile_wall_four_64M", "configs_total": B5, "configs_per_gpu": n5, "n_gpus": world,
                        "ms_per_sweep": sweep_ms, "configs_per_s": B5 / (sweep_ms * 1e-3), "scaling": "strong",
                        "inputs": "uniform in joint limits, drawn on the device (torch.rand), fp32, resident",
                        "free_fraction": float(f5.float().mean().item()),
                        "exchange": "all_gather of flag bytes (NCCL)" if g5 is not None else "none (ragged shards)"}
        del q5, f5, g5

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    flop_cfg = S.algorithmic_flops_per_config(cs)
    sm_max_mhz = float((clk or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
    fp32_peak = SM_COUNT * FP32_LANES * 2 * sm_max_mhz * 1e6 / 1e12       # 74.5 TFLOP/s at 1965 MHz
    fp32_probe = fp32_fma_peak_tflops(dev)
    achieved = flop_cfg * B / (kernel_ms * 1e-3) / 1e12
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_cfg = 4 * D + 1
    cap_expect = {"configs": B}
    executed_flop = capture(args.workload, "executed_fp32_flop", cap_expect)
    roofline = {
        "bound": "fp32_simt", "kernel": "check_configs_kernel", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak,
        "traffic": capture(args.workload, "dram_bytes", cap_expect),
        "algorithmic_bytes": bytes_cfg * B,
        "bound_note": "FK + narrowphase is FP32-SIMT work (SURVEY.md 8d); `achieved` counts ALGORITHMIC flop (every collidable pair of "
                      "the mode, no credit taken for culling), so frac > 1 means the broadphase culls more than the pipe could "
                      "compute; `executed` / `issue_active` say how busy the SM really is",
        "peak_source": f"theoretical: {SM_COUNT} SMs x {FP32_LANES} lanes x 2 flop x {sm_max_mhz:.0f} MHz (MEASURED_PEAKS.json has no FP32-SIMT "
                       f"figure); live FMA probe (mrb200_fp32_probe) reached {fp32_probe:.1f} TFLOP/s",
        "peak_probe": fp32_probe,
        "algorithmic_flop_per_config": flop_cfg, "kernel_ms": kernel_ms,
        "executed": None if executed_flop is None else {
            "fp32_flop_per_launch": executed_flop, "tflops": executed_flop / (kernel_ms * 1e-3) / 1e12,
            "frac_of_peak": executed_flop / (kernel_ms * 1e-3) / 1e12 / fp32_peak,
            "source": "profiles/traffic.json (ncu --set full of this build: 2*FFMA + FADD + FMUL thread instructions)"},
        "issue_active": capture(args.workload, "issue_active_pct", cap_expect),
        "warps_active": capture(args.workload, "warps_active_pct", cap_expect),
        "capture": capture(args.workload, "source", cap_expect),
        "hbm": {"achieved": bytes_cfg * B / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_cfg * B / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }

    out = {
        "metric": METRIC, "value": value, "unit": "configs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_block(args, scene_name, B, world, D, cs),   # the same keys and values as the reference arm's line
        "free_fraction": free_frac,
        "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": world * B * D * 4,
                "d2h_bytes_per_step": world * B, "steps": e2e_steps, "timing": "CUDA events on the calling stream (which waits for the copy streams) around "
                "calls that return after the flags have landed in host memory, max over ranks",
                "h2d_only_gbs_per_rank_min": h2d_gbs_min, "numa": numa, "h2d_needed_gbs_per_rank_at_value": value / world * D * 4 / 1e9,
                "path": "mrb200_check_configs_host: pinned host -> H2D -> check_configs -> D2H, 256k-config chunks on 3 side streams of the library"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline,
    }
    if strong_sweep is not None:
        out["strong_sweep"] = strong_sweep

    # ---- secondary blocks (N = 1) ----
    if not args.no_extra:
        def local_edges(be_, slot, lim_, n, seed):
            """planner-like local edges: free start, every joint moves by at most +-0.2 (mean N ~ 19 at resolution 0.01)"""
            l1 = torch.from_numpy(uniform_configs(lim_, n, seed)).to(dev)
            l1 = l1[be_.check_configs(slot, l1).bool()].contiguous()
            stp = torch.from_numpy(np.random.RandomState(seed + 1).uniform(-0.2, 0.2, tuple(l1.shape)).astype(np.float32)).to(dev)
            lo_t, hi_t = torch.from_numpy(lim_[0].astype(np.float32)).to(dev), torch.from_numpy(lim_[1].astype(np.float32)).to(dev)
            return l1, torch.minimum(torch.maximum(l1 + stp, lo_t), hi_t).contiguous()

        def edge_stats(be_, slot, a, b, res, W_cfg, reps=3):
            fr, first = be_.check_edges(slot, a, b, res)
            t = timed(lambda: be_.check_edges(slot, a, b, res), reps)
            N = torch.clamp((torch.max(torch.abs(a.double() - b.double()), dim=1).values / res).long() + 1, min=2)
            interior = (N - 2).clamp(min=0)
            required = torch.where(fr.bool(), interior, first.long() + 1).sum().item()   # up to and including the first hit, binary order
            n = a.shape[0]
            return {"edges": int(n), "edges_per_s": n / t, "ms": t * 1e3, "free_frac": float(fr.float().mean().item()),
                    "mean_N": float(N.float().mean().item()), "required_samples": int(required), "interior_samples": int(interior.sum().item()),
                    "required_samples_per_s": required / t,
                    "roofline": {"bound": "fp32_simt", "kernel": "check_edges_kernel", "unit": "TFLOP/s", "peak": fp32_peak,
                                 "achieved": required * W_cfg / t / 1e12, "frac": required * W_cfg / t / 1e12 / fp32_peak,
                                 "upper_figure_all_interior": float(interior.sum().item()) * W_cfg / t / 1e12,
                                 "definition": "W_edge = (#samples the reference visits up to and including the first collision) x W_cfg (SURVEY.md 8d)"}}

        # --- edges of the headline scene (BASELINE metric 2)
        q1 = torch.from_numpy(uniform_configs(lim, E, 8)).to(dev)
        q2 = torch.from_numpy(uniform_configs(lim, E, 9)).to(dev)
        l1, l2 = local_edges(be, 0, lim, 131_072, 10)
        edges = {"scene": scene_name, "resolution": kw["resolution"],
                 "uniform": edge_stats(be, 0, q1, q2, kw["resolution"], flop_cfg),
                 "local": edge_stats(be, 0, l1, l2, kw["resolution"], flop_cfg),
                 "capture": {k: capture("edges_" + scene_name, k) for k in ("issue_active_pct", "no_instruction_stall", "warps_active_pct", "source")}}
        if not args.no_cpu:
            from oracle import oracle_scene as O
            nth = host_threads()
            n_cpu = min(20_000, l1.shape[0])
            a64, b64 = l1[:n_cpu].cpu().numpy().astype(np.float64), l2[:n_cpu].cpu().numpy().astype(np.float64)
            t = time.perf_counter()
            O.check_edges(cs.blob64, a64, b64, kw["resolution"], nthreads=nth)
            edges["local"]["cpu_port_edges_per_s"] = n_cpu / (time.perf_counter() - t)
            u1, u2 = q1.cpu().numpy().astype(np.float64), q2.cpu().numpy().astype(np.float64)
            t = time.perf_counter()
            O.check_edges(cs.blob64, u1, u2, kw["resolution"], nthreads=nth)
            edges["uniform"]["cpu_port_edges_per_s"] = len(u1) / (time.perf_counter() - t)
            edges["cpu_cores"] = nth
            edges["cpu_sample"] = f"{n_cpu} local + {len(u1)} uniform edges, fp64 oracle port, {nth} threads"
        out["edges"] = edges
        log(f"edges: uniform {edges['uniform']['edges_per_s']:.3g}/s local {edges['local']['edges_per_s']:.3g}/s")
        del q1, q2, l1, l2

        # --- held-object modes and all-free batches of the headline scene
        if scene_name == "box_rearrangement":
            from multirobot_pathplanning_benchmark_b200.env import SceneModel
            model = SceneModel(sc, kw["tol"], kw["resolution"], device=None)
            base_slot = model.slot_for(())
            cand = uniform_configs(lim, 8192, 21)
            okc = model.check_configs(base_slot, torch.from_numpy(cand).to(dev)).cpu().numpy().astype(bool)
            q_att = cand[int(np.argmax(okc))].astype(np.float64)
            mslots = {"start": base_slot,
                      "a1_holds_obj11": model.slot_for(("a1",), [("a1_ur_vacuum", "obj11", q_att)]),
                      "a2_holds_obj00": model.slot_for(("a2",), [("a2_ur_vacuum", "obj00", q_att)])}
            qm = torch.from_numpy(uniform_configs(lim, 2_097_152, 22)).to(dev)
            modes = {}
            for mname, mslot in mslots.items():
                tm = timed(lambda: model.check_configs(mslot, qm), 5)
                fm = model.check_configs(mslot, qm)
                # all-free batch (what dense validation near a solution sees): the free samples of 4 x 2M draws, tiled to 1M
                pool = torch.cat([qm[fm.bool()]] + [(lambda t_: t_[model.check_configs(mslot, t_).bool()])(
                    torch.from_numpy(uniform_configs(lim, 2_097_152, 30 + i)).to(dev)) for i in range(2)])
                allfree = pool[:1_048_576].contiguous() if pool.shape[0] >= 1_048_576 else pool.repeat((1_048_576 // max(pool.shape[0], 1)) + 1, 1)[:1_048_576].contiguous()
                ta = timed(lambda: model.check_configs(mslot, allfree), 5)
                modes[mname] = {"configs_per_s": qm.shape[0] / tm, "free_frac": float(fm.float().mean().item()),
                                "all_free_batch_configs_per_s": allfree.shape[0] / ta,
                                "collidable_pairs": int(sum(model.compiled(mslot).pair_counts))}
                del pool, allfree
            third = qm.shape[0] // 3
            parts = [qm[i * third:(i + 1) * third].contiguous() for i in range(3)]
            tmix = timed(lambda: [model.check_configs(sl_, p_) for sl_, p_ in zip(mslots.values(), parts)], 5)
            modes["mixed_thirds"] = {"configs_per_s": 3 * third / tmix}
            out["modes"] = modes
            del qm, parts
            log(f"modes: {json.dumps({k: round(v['configs_per_s'] / 1e9, 3) for k, v in modes.items()})}")

        # --- BASELINE config 4: batched k-NN / r-disc for PRM / EIT graph building, 100k samples of one mode, D = 24
        from multirobot_pathplanning_benchmark_b200 import knn as K
        Nk, Dk, kk = 100_000, 24, K.prm_k_star(100_000, 24)
        slk = [[6 * r, 6 * r + 6] for r in range(4)]
        limk = SCENES["box_stacking"][0]().limits()
        corpus = torch.from_numpy(np.random.RandomState(5).uniform(limk[0], limk[1], (Nk, Dk))).to(dev)
        knn_res = {"N": Nk, "Q": Nk, "D": Dk, "k": kk, "metric": "max_euclidean"}
        idx_by_mode = {}
        for mode in ("tensor", "exact"):
            K.batch_knn(corpus[:4096], corpus, slk, "max_euclidean", kk, mode=mode)
            tk = timed(lambda: K.batch_knn(corpus, corpus, slk, "max_euclidean", kk, mode=mode), 2)
            idx_by_mode[mode] = K.batch_knn(corpus, corpus, slk, "max_euclidean", kk, mode=mode)
            knn_res[f"{mode}_queries_per_s"] = Nk / tk
            knn_res[f"{mode}_ms"] = tk * 1e3
        same = torch.equal(idx_by_mode["tensor"][0], idx_by_mode["exact"][0]) and torch.equal(idx_by_mode["tensor"][1], idx_by_mode["exact"][1])
        assert same, "k-NN: the tcgen05 path and the exact fp64 path returned different neighbours"
        knn_res["tensor_equals_exact"] = bool(same)
        flop = 2.0 * Nk * Nk * Dk
        tf32_peak = 0.5 * peaks.get("bf16_tflops", 1643.0)
        knn_res["roofline"] = {"bound": "tensor", "kernel": "knn_tc_kernel", "unit": "TFLOP/s", "achieved": flop / (knn_res["tensor_ms"] * 1e-3) / 1e12,
                               "peak": tf32_peak, "frac": flop / (knn_res["tensor_ms"] * 1e-3) / 1e12 / tf32_peak,
                               "peak_source": "TF32 ceiling = half of MEASURED_PEAKS.json bf16_tflops", "algorithmic_flop": flop,
                               "tensor_pipe_active_pct": capture("knn_tc", "tensor_pipe_active_pct"),
                               "kernel_ms_in_capture": capture("knn_tc", "kernel_ms"), "capture": capture("knn_tc", "source")}
        _, d33 = K.batch_knn(corpus[:2048].contiguous(), corpus, slk, "max_euclidean", kk)
        r_sel = float(d33[:, -1].median().item())
        tr = timed(lambda: K.batch_radius(corpus, corpus, r_sel, slk, "max_euclidean"), 2)
        off, _ = K.batch_radius(corpus, corpus, r_sel, slk, "max_euclidean")
        knn_res["radius"] = {"r": r_sel, "ms": tr * 1e3, "mean_neighbours": float(off[-1].item()) / Nk,
                             "note": "selective radius (median distance of the 33rd neighbour); the PRM* r* of prm_graph.py:479-500 "
                                     "spans nearly the whole corpus at D = 24"}
        if not args.no_cpu:
            from oracle import oracle_abstract as OA
            cn = corpus.cpu().numpy()
            t = time.perf_counter()
            nq = 40
            for j in range(nq):
                OA.knn_indices(OA.batch_config_dist(cn[j], cn, np.array(slk), "max_euclidean"), kk)
            knn_res["cpu_port_queries_per_s_1core"] = nq / (time.perf_counter() - t)
        out["knn"] = knn_res
        del corpus, idx_by_mode
        log(f"knn: tensor {knn_res['tensor_ms']:.2f} ms exact {knn_res['exact_ms']:.1f} ms")

        # --- the other named scenes, same protocol, fewer steps
        scenes = {}
        for wname, (sname, Bw, Ew) in WORKLOADS.items():
            if wname in STRONG or wname == args.workload:
                continue
            mk2, kw2 = SCENES[sname]
            sc2 = mk2()
            cs2 = S.compile_blob(sc2, kw2["tol"])
            be.set_mode(1, cs2)
            lim2 = sc2.limits()
            W2 = S.algorithmic_flops_per_config(cs2)
            Bc = min(Bw, 2_097_152)
            qd = torch.from_numpy(uniform_configs(lim2, Bc, 7)).to(dev)
            tc = timed(lambda: be.check_configs(1, qd), 5)
            u1 = torch.from_numpy(uniform_configs(lim2, Ew, 8)).to(dev)
            u2 = torch.from_numpy(uniform_configs(lim2, Ew, 9)).to(dev)
            a1, a2 = local_edges(be, 1, lim2, 131_072, 10)
            scenes[wname] = {"configs": Bc, "configs_per_s": Bc / tc, "config_free_frac": float(be.check_configs(1, qd).float().mean().item()),
                             "algorithmic_flop_per_config": W2, "roofline_frac": W2 * Bc / tc / 1e12 / fp32_peak,
                             "edge_resolution": kw2["resolution"],
                             "edges_uniform": edge_stats(be, 1, u1, u2, kw2["resolution"], W2),
                             "edges_local": edge_stats(be, 1, a1, a2, kw2["resolution"], W2)}
            del qd, u1, u2, a1, a2
        out["scenes"] = scenes
        # BASELINE config 1 (abstract.test, D = 4): micro-batches on the fp64 abstract kernels, bit-exact with the reference
        from multirobot_pathplanning_benchmark_b200.backend import AbstractBackend
        ab = AbstractBackend(2, 2, [0.1, 0.1], spheres=[([0.0, 0.0], 0.2)], rects_minmax=[([-0.25, 0.15], [0.25, 0.65])], device=dev)
        rs = np.random.RandomState(3)
        qa = torch.from_numpy(rs.uniform(-2, 2, (1_000_000, 4))).to(dev)
        ea, eb = torch.from_numpy(rs.uniform(-2, 2, (100_000, 4))).to(dev), torch.from_numpy(rs.uniform(-2, 2, (100_000, 4))).to(dev)
        ta = timed(lambda: ab.check_configs(qa), 5)
        tea = timed(lambda: ab.check_edges(ea, eb, 0.01), 3)
        abstract = {"configs": 1_000_000, "configs_per_s": 1_000_000 / ta, "edges": 100_000, "edges_per_s": 100_000 / tea,
                    "resolution": 0.01, "free_frac": float(ab.check_configs(qa).float().mean().item())}
        if not args.no_cpu:
            from oracle import oracle_abstract as OA
            osc = OA.AbstractScene.abstract_test()
            qh = qa[:200_000].cpu().numpy()
            t = time.perf_counter()
            of = osc.batch_flags(qh)
            abstract["cpu_port_vectorised_configs_per_s_1core"] = len(qh) / (time.perf_counter() - t)
            t = time.perf_counter()
            for row in qh[:2000]:
                osc.is_collision_free(row)
            abstract["cpu_port_per_call_configs_per_s_1core"] = 2000 / (time.perf_counter() - t)
            abstract["flags_identical"] = bool(np.array_equal(of, ab.check_configs(qa[:200_000]).cpu().numpy()))
        out["abstract_test"] = abstract
        del qa, ea, eb
        log("scenes + abstract done")

        # --- BASELINE metric 3: the reference's own planners on the device
        if not args.no_planners:
            torch.cuda.synchronize()
            out["planners"] = planners_block(args, log)
            # batch-native PRM (planner.py): whole-batch calls instead of the reference's lazy per-query loop
            try:
                sys.path.insert(0, os.path.join(ROOT, "scripts"))
                import ttfs
                ttfs.run("2d_handover", "b200", 99, 200, 30, 30)   # warm-up
                gpu_runs = [ttfs.run("box_stacking", "b200", seed, 3000, 300, 120, n_moves=4) for seed in range(3)]
                entry = {"scene": "box_stacking", "samples_per_mode": 3000, "pick_place_moves": 4, "modes": 9,
                         "b200_median_s": float(np.median([r["time_s"] for r in gpu_runs]))}
                if not args.no_cpu:
                    cpu_run = ttfs.run("box_stacking", "cpu", 0, 3000, 300, 600, n_moves=4)
                    entry["cpu_port_s"] = cpu_run["time_s"]
                    entry["cpu_cores"] = host_threads()
                    entry["same_plan"] = abs(cpu_run["cost"] - gpu_runs[0]["cost"]) < 1e-9
                out["batched_prm"] = entry
            except Exception as e:
                out["batched_prm"] = {"error": repr(e)}

    if not args.no_cpu:
        out["cpu_baseline"] = cpu_baseline(cs, lim, B, "restatement of the rai query, not rai")
    else:
        out["cpu_baseline"] = None

    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
