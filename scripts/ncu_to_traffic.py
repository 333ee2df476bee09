"""Reads the `ncu --set full` captures brought back from the GPU box (gpurun_out/cap_TAG_*.ncu-rep, scripts/capture_all.sh)
and writes (a) profiles/TAG_<name>.txt, the key metrics of each capture, and (b) profiles/traffic.json, the per-launch
figures bench.py quotes (`roofline.traffic`, `roofline.executed`, `issue_active`, tensor-pipe activity).
usage: python scripts/ncu_to_traffic.py TAG"""
import csv, glob, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAG = sys.argv[1] if len(sys.argv) > 1 else "r2"
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__", "sm__inst_executed_pipe_", "sm__pipe_tensor",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warp", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct", "smsp__cycles_elapsed.avg", "sm__issue_active.avg", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__sass_thread_inst_executed_op_f", "dram__throughput.avg.pct",
        "lts__t_sector_hit_rate.pct", "sm__cycles_active.avg", "smsp__cycles_active.avg", "smsp__warp_issue_stalled", "sm__cycles_elapsed.max")
WORKLOAD_OF = {"configs_box_rearrangement_4M": ("box_rearrangement_4M", 4194304), "configs_box_stacking_1M": ("box_stacking_1M", 1048576),
               "configs_mobile_wall_four_2M": ("mobile_wall_four_2M", 2097152)}


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


traffic = {}
for rep in sorted(glob.glob(os.path.join(ROOT, "gpurun_out", f"cap_{TAG}_*.ncu-rep"))):
    name = os.path.basename(rep)[len(f"cap_{TAG}_"):-len(".ncu-rep")]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("empty report", rep)
        continue
    hdr, units, v = rows[0], rows[1], rows[2]
    m = {h: (num(x), u) for h, u, x in zip(hdr, units, v)}
    txt = os.path.join("profiles", f"{TAG}_{name}.txt")
    with open(os.path.join(ROOT, txt), "w") as f:
        f.write(f"# ncu --set full --clock-control none, one launch after warm-up, build of round {TAG}: {name}\n")
        f.write(f"## {v[hdr.index('Kernel Name')]}  grid {v[hdr.index('Grid Size')]} block {v[hdr.index('Block Size')]}\n")
        for h, u, x in sorted(zip(hdr, units, v)):
            if h.startswith(KEEP):
                f.write(f"{h} [{u}] = {x}\n")

    def g(key, scale_unit=True):
        val, unit = m.get(key, (None, ""))
        if val is None:
            return None
        if scale_unit:
            val *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-3, "ms": 1.0, "ns": 1e-6, "s": 1e3}.get(unit, 1)
        return val
    # thread instructions per elapsed cycle (summed over the SM sub-partitions) x elapsed cycles
    cyc = g("smsp__cycles_elapsed.avg", False) or 0
    ffma, fadd, fmul = ((g(f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed", False) or 0) * cyc
                        for o in ("ffma", "fadd", "fmul"))
    stall_noinst = g("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", False)
    if stall_noinst is None:
        stall_noinst = g("smsp__average_warp_latency_issue_stalled_no_instruction.ratio", False)
    entry = {"source": f"{txt} (ncu --set full of the round-{TAG} build)", "kernel": v[hdr.index("Kernel Name")][:60],
             "kernel_ms": g("gpu__time_duration.sum"), "dram_bytes": (g("dram__bytes_read.sum") or 0) + (g("dram__bytes_write.sum") or 0),
             "executed_fp32_flop": 2 * ffma + fadd + fmul, "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
             "sm_throughput_pct": g("sm__throughput.avg.pct_of_peak_sustained_elapsed", False),
             "warp_instructions": g("smsp__inst_executed.sum", False),
             "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active", False),
             "registers": g("launch__registers_per_thread", False), "no_instruction_stall": stall_noinst,
             "tensor_pipe_active_pct": g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", False)}
    if entry["tensor_pipe_active_pct"] is None:
        for h in hdr:
            if h.startswith("sm__pipe_tensor") and "pct_of_peak_sustained_active" in h:
                entry["tensor_pipe_active_pct"] = m[h][0]
                break
    if name in WORKLOAD_OF:
        w, B = WORKLOAD_OF[name]
        entry["configs"] = B
        traffic[w] = entry
    else:
        traffic[name] = entry
    print(name, json.dumps(entry))
p = os.path.join(ROOT, "profiles", "traffic.json")
old = {}
try:
    old = json.load(open(p))
except Exception:
    pass
old.update(traffic)
json.dump(old, open(p, "w"), indent=1)
