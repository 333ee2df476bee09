#!/usr/bin/env python
"""Benchmark of the collision hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic input: `check_configs` over
the workload's configuration batch (uniform in the joint limits, exactly like the reference's
sampler, P/problems/planning_env.py:1697-1708).  Default workload: the dual-arm scene of
BASELINE.json (box_rearrangement, 4 194 304 configurations), the configuration north_star's
>= 1e9 checks/s target is quoted on.  Prints ONE JSON line on rank 0.

  value      whole-job configuration checks/s with the batch resident in HBM
  e2e        the same through the host-buffer API: pinned host configs -> H2D -> kernel -> D2H flags
  roofline   FP32-SIMT roofline of the FK+narrowphase kernel (algorithmic flop)
  cpu_baseline  the fp64 CPU oracle (a port: rai itself cannot be installed) on the host cores
  extra      edge checks/s and the other named scenes, same protocol, fewer steps

--impl reference times the reference's CPU path for the same metric/config: the reference's own
backend (rai) is an un-vendored wheel and cannot run here, so this arm runs the fp64 oracle port
with all host threads on a bounded sample.
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


def ncu_capture(workload, B, key):
    """A per-launch figure of this workload's kernel from the committed `ncu --set full` capture
    (profiles/traffic.json, filled from scripts/ncu_summary.py output): `dram_bytes` = dram__bytes_read.sum +
    dram__bytes_write.sum, `executed_fp32_flop` = 2*FFMA + FADD + FMUL thread instructions.  None if the capture
    was taken at another batch size."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(workload)
        return float(t[key]) if t and int(t["configs"]) == int(B) else None
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
def cpu_baseline(cs, lim, kind_note, budget_s=12.0, edges=None, resolution=None):
    """fp64 oracle port on all host threads, bounded sample of the same workload."""
    from oracle import oracle_scene as O
    nthreads = O.max_threads()
    probe = uniform_configs(lim, 20_000, 123).astype(np.float64)
    t = time.perf_counter()
    O.check_configs(cs.blob64, probe, nthreads=nthreads)
    rate = len(probe) / (time.perf_counter() - t)
    n = int(min(max(rate * budget_s, 50_000), 4_000_000))
    q = uniform_configs(lim, n, 0).astype(np.float64)
    reps = max(1, int(round(budget_s / max(n / rate, 1e-3))))
    t = time.perf_counter()
    for _ in range(reps):
        O.check_configs(cs.blob64, q, nthreads=nthreads)
    dt = time.perf_counter() - t
    out = {"value": n * reps / dt, "unit": "configs/s", "cores": nthreads, "kind": "port",
           "sample": f"{reps} passes over {n} uniform configs of the same workload, fp64 C oracle ({kind_note}), "
                     f"{nthreads} OpenMP threads, {dt:.1f} s"}
    if edges is not None:
        q1, q2 = edges
        m = min(len(q1), 2000)
        t = time.perf_counter()
        O.check_edges(cs.blob64, q1[:m].astype(np.float64), q2[:m].astype(np.float64), resolution, nthreads=nthreads)
        out["edges_per_s"] = m / (time.perf_counter() - t)
    return out


def run_reference(args):
    """Reference arm: the reference's CPU implementation of this path.  rai (`robotic`) is an
    un-vendored third-party wheel that cannot be installed offline, so this is the fp64 oracle port
    with every host thread; one step = one bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from multirobot_pathplanning_benchmark_b200 import scene as S
    from multirobot_pathplanning_benchmark_b200.scenes import SCENES
    from oracle import oracle_scene as O
    scene_name, B, _ = WORKLOADS[args.workload]
    mk, kw = SCENES[scene_name]
    sc = mk()
    cs = S.compile_blob(sc, kw["tol"])
    lim = sc.limits()
    nthreads = O.max_threads()
    probe = uniform_configs(lim, 20_000, 123).astype(np.float64)
    t = time.perf_counter()
    O.check_configs(cs.blob64, probe, nthreads=nthreads)
    rate = len(probe) / (time.perf_counter() - t)
    total_steps = args.steps + args.warmup
    n = int(min(max(rate * args.ref_seconds / max(total_steps, 1), 2_000), B))  # whole run ~ref_seconds
    q = uniform_configs(lim, n, 0).astype(np.float64)
    for _ in range(args.warmup):
        O.check_configs(cs.blob64, q, nthreads=nthreads)
    t = time.perf_counter()
    for _ in range(args.steps):
        O.check_configs(cs.blob64, q, nthreads=nthreads)
    dt = time.perf_counter() - t
    value = n * args.steps / dt
    sample = f"{n} of {B} uniform configs per step, fp64 C oracle port of the rai query, {nthreads} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "configs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "scene": scene_name, "configs_per_step": n},
        "cpu_baseline": {"value": value, "unit": "configs/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ----------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT, choices=list(WORKLOADS))
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary scenes / edge rates")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    ap.add_argument("--ref-seconds", type=float, default=60.0, help="--impl reference: CPU time budget of the whole run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:  # secondary scenes and the CPU leg are N = 1 only
        args.no_extra = args.no_cpu = True

    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    from multirobot_pathplanning_benchmark_b200 import scene as S
    from multirobot_pathplanning_benchmark_b200.backend import (SceneBackend, check_configs_host, fp32_fma_peak_tflops,
                                                                launch_count)
    from multirobot_pathplanning_benchmark_b200.scenes import SCENES

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
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

    # ---- end to end through the host-buffer API (pinned host in, pinned host out) ----
    out_host = torch.empty(B, dtype=torch.uint8).pin_memory()
    state = {}
    for _ in range(2):
        check_configs_host(be, 0, q_host, out_host, state=state)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        check_configs_host(be, 0, q_host, out_host, state=state)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = world * B * e2e_steps / (e2e_ms * 1e-3)
    clk = clocks.stop() if rank == 0 else None   # sampled over both timed regions (resident and end-to-end)
    assert torch.equal(out_host.to(dev), flags), "host-API flags differ from the device-resident run"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----
    flop_cfg = S.algorithmic_flops_per_config(cs)
    fp32_peak = fp32_fma_peak_tflops(dev)
    achieved = flop_cfg * B / (kernel_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    bytes_cfg = 4 * D + 1
    roofline = {
        "bound": "fp32_simt", "kernel": "check_configs_kernel", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": achieved / fp32_peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this workload, from the committed
        # `ncu --set full` capture (profiles/r1_check_configs_v2_broadphase.txt): 203.5 MB + 17.7 MB
        "traffic": ncu_capture(args.workload, B, "dram_bytes"),
        "algorithmic_bytes": (4 * D + 1) * B,
        "bound_note": "FK + narrowphase is FP32-FMA bound (SURVEY.md 8d); the HBM view is reported under 'hbm'",
        "peak_source": "measured live by mrb200_fp32_probe (MEASURED_PEAKS.json has no FP32-SIMT figure)",
        "algorithmic_flop_per_config": flop_cfg, "kernel_ms": kernel_ms,
        # what the SIMT pipes really execute after culling / early exit (ncu capture): the algorithmic figure above
        # counts every collidable pair of the mode and can therefore exceed the pipe's peak
        "executed": (lambda f: None if f is None else {
            "fp32_flop_per_launch": f, "tflops": f / (kernel_ms * 1e-3) / 1e12, "frac_of_peak": f / (kernel_ms * 1e-3) / 1e12 / fp32_peak,
            "source": "profiles/traffic.json (ncu --set full, smsp__sass_thread_inst_executed_op_{ffma,fadd,fmul}_pred_on)"})(
                ncu_capture(args.workload, B, "executed_fp32_flop")),
        "hbm": {"achieved": bytes_cfg * B / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": bytes_cfg * B / (kernel_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"},
    }

    # ---- secondary numbers: edges on this scene, other scenes ----
    extra = {}
    if not args.no_extra:
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

        edge_inputs = None
        for wname, (sname, Bw, Ew) in WORKLOADS.items():
            if wname in STRONG and wname != args.workload:
                continue  # same scene as its per-GPU sibling
            mk2, kw2 = SCENES[sname]
            sc2 = mk2()
            cs2 = S.compile_blob(sc2, kw2["tol"])
            be.set_mode(1, cs2)
            lim2 = sc2.limits()
            Bc = min(Bw, 2_097_152)
            qd = torch.from_numpy(uniform_configs(lim2, Bc, 7)).to(dev)
            q1 = torch.from_numpy(uniform_configs(lim2, Ew, 8)).to(dev)
            q2 = torch.from_numpy(uniform_configs(lim2, Ew, 9)).to(dev)
            tc = timed(lambda: be.check_configs(1, qd), 5)
            fr, first = be.check_edges(1, q1, q2, kw2["resolution"])
            te = timed(lambda: be.check_edges(1, q1, q2, kw2["resolution"]), 3)
            # planner-like local edges: every joint moves by at most +-0.2 (mean N ~ 19 at resolution 0.01)
            El = 131_072
            l1 = torch.from_numpy(uniform_configs(lim2, El, 10)).to(dev)
            l1 = l1[be.check_configs(1, l1).bool()].contiguous()
            step = torch.from_numpy(np.random.RandomState(11).uniform(-0.2, 0.2, tuple(l1.shape)).astype(np.float32)).to(dev)
            lo_t, hi_t = torch.from_numpy(lim2[0].astype(np.float32)).to(dev), torch.from_numpy(lim2[1].astype(np.float32)).to(dev)
            l2 = torch.minimum(torch.maximum(l1 + step, lo_t), hi_t).contiguous()
            lf, _ = be.check_edges(1, l1, l2, kw2["resolution"])
            tl = timed(lambda: be.check_edges(1, l1, l2, kw2["resolution"]), 3)
            extra[wname] = {"configs_per_s": Bc / tc, "local_edges_per_s": l1.shape[0] / tl, "local_edge_free_frac": float(lf.float().mean().item()),
                            "local_edges": int(l1.shape[0]), "config_free_frac": float(be.check_configs(1, qd).float().mean().item()),
                            "edges_per_s": Ew / te, "edge_free_frac": float(fr.float().mean().item()),
                            "edge_resolution": kw2["resolution"], "edges": Ew, "configs": Bc,
                            "algorithmic_flop_per_config": S.algorithmic_flops_per_config(cs2)}
            if wname == args.workload:
                edge_inputs = (q1.cpu().numpy(), q2.cpu().numpy())
            del qd, q1, q2
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
        extra["abstract_test"] = abstract
        del qa, ea, eb
        # modes of the default scene: start mode, a box held by a1, a box held by a2, and a mode-mixed batch
        # (the reference's benchmark draws a random reachable mode per sample, scripts/show_problems.py:177-181)
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
                modes[mname] = {"configs_per_s": qm.shape[0] / tm, "free_frac": float(model.check_configs(mslot, qm).float().mean().item()),
                                "collidable_pairs": int(sum(model.compiled(mslot).pair_counts))}
            third = qm.shape[0] // 3
            parts = [qm[i * third:(i + 1) * third].contiguous() for i in range(3)]
            tmix = timed(lambda: [model.check_configs(sl_, p_) for sl_, p_ in zip(mslots.values(), parts)], 5)
            modes["mixed_thirds"] = {"configs_per_s": 3 * third / tmix}
            extra["modes_box_rearrangement"] = modes
            del qm, parts
        # BASELINE config 4: batched k-NN for PRM/EIT graph building, 100k samples of one mode, D = 24
        from multirobot_pathplanning_benchmark_b200 import knn as K
        Nk, Dk, kk = 100_000, 24, K.prm_k_star(100_000, 24)
        slk = [[6 * r, 6 * r + 6] for r in range(4)]
        limk = SCENES["box_stacking"][0]().limits()
        corpus = torch.from_numpy(np.random.RandomState(5).uniform(limk[0], limk[1], (Nk, Dk))).to(dev)
        knn_res = {"N": Nk, "Q": Nk, "D": Dk, "k": kk, "metric": "max_euclidean"}
        for mode in ("tensor", "exact"):
            K.batch_knn(corpus[:4096], corpus, slk, "max_euclidean", kk, mode=mode)
            tk = timed(lambda: K.batch_knn(corpus, corpus, slk, "max_euclidean", kk, mode=mode), 2)
            knn_res[f"{mode}_queries_per_s"] = Nk / tk
            knn_res[f"{mode}_ms"] = tk * 1e3
        knn_res["tensor_algorithmic_tflops"] = 2.0 * Nk * Nk * Dk / (knn_res["tensor_ms"] * 1e-3) / 1e12
        if not args.no_cpu:
            from oracle import oracle_abstract as OA
            cn = corpus.cpu().numpy()
            t = time.perf_counter()
            nq = 40
            for j in range(nq):
                OA.knn_indices(OA.batch_config_dist(cn[j], cn, np.array(slk), "max_euclidean"), kk)
            knn_res["cpu_port_queries_per_s_1core"] = nq / (time.perf_counter() - t)
        extra["knn_box_stacking_100k"] = knn_res
        del corpus
        # time-to-first-solution of the batch-native PRM (planner.py): same planner code, same seeds, B200 backend
        # vs the CPU oracle backend answering the same batch calls (BASELINE.md plan item 4)
        sys.path.insert(0, os.path.join(ROOT, "scripts"))
        import ttfs
        ttfs.run("2d_handover", "b200", 99, 200, 30, 30)   # warm-up
        tt = {}
        # 2d_handover: pick / handover / place sequence (6 modes); box_rearrangement: 2 pick-and-place moves with the
        # vacuum tools (5 modes); box_stacking: the four arms stack 4 boxes (9 modes); mobile_wall_four: two robots
        # move their wall columns (4 moves, 9 modes); keyframes from problems.py
        for sname, n0, t0, cpu_seeds, n_moves in (("2d_handover", 500, 60, 3, 0), ("box_rearrangement", 2000, 200, 1, 2),
                                                  ("box_stacking", 3000, 300, 1, 4), ("mobile_wall_four", 1500, 150, 1, 4)):
            gpu_runs = [ttfs.run(sname, "b200", seed, n0, t0, 120, n_moves=n_moves) for seed in range(3)]
            entry = {"samples_per_mode": n0, "pick_place_moves": n_moves, "modes": 6 if sname == "2d_handover" else 2 * n_moves + 1,
                     "b200_median_s": float(np.median([r["time_s"] for r in gpu_runs])), "b200_runs": gpu_runs}
            if not args.no_cpu:
                cpu_runs = [ttfs.run(sname, "cpu", seed, n0, t0, 600, n_moves=n_moves) for seed in range(cpu_seeds)]
                entry["cpu_port_median_s"] = float(np.median([r["time_s"] for r in cpu_runs]))
                entry["cpu_runs"] = cpu_runs
                entry["same_plans"] = all(abs(a["cost"] - b["cost"]) < 1e-9 for a, b in zip(gpu_runs, cpu_runs))
            tt[sname] = entry
        extra["prm_time_to_first_solution"] = tt
    else:
        edge_inputs = None

    cpu = None
    if not args.no_cpu:
        cpu = cpu_baseline(cs, lim, "restatement of the rai query, not rai", edges=edge_inputs, resolution=kw["resolution"])

    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "configs/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": scene_name, "configs_per_gpu": B, "configs_total": world * B, "dof": D,
                   "collidable_pairs": int(sum(cs.pair_counts)), "tolerance": cs.tol,
                   "inputs": "uniform in joint limits (np.random.uniform), fp32, resident in HBM",
                   "l2": f"input batch {B * D * 4 / 1e6:.0f} MB > 126 MB L2, streamed once per step",
                   "free_fraction": free_frac,
                   "exchange": "all_gather of flag bytes (NCCL)" if world > 1 else "none"},
        "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": world * B * D * 4,
                "d2h_bytes_per_step": world * B, "steps": e2e_steps,
                "path": "pinned host -> H2D -> check_configs -> D2H, 512k-config chunks on 2 streams"},
        "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu, "extra": extra,
    }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
