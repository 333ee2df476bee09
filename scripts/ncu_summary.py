"""Key metrics of an ncu report as `name = value` lines: python scripts/ncu_summary.py report.ncu-rep [header text]"""
import csv, io, subprocess, sys
KEEP = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "launch__block_size", "launch__grid_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit",
        "sm__inst_executed_pipe_alu.avg.pct", "sm__inst_executed_pipe_fma.avg.pct", "sm__inst_executed_pipe_fmaheavy.avg.pct",
        "sm__inst_executed_pipe_lsu.avg.pct", "sm__inst_executed_pipe_xu.avg.pct", "sm__inst_executed_pipe_fp64.avg.pct",
        "sm__inst_executed_pipe_tensor", "sm__pipe_tensor", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warps_issue_stalled", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct", "smsp__thread_inst_executed_per_inst_executed.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "dram__throughput.avg.pct", "lts__t_sector_hit_rate.pct",
        "sm__cycles_active.avg", "smsp__cycles_active.avg")
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units, vals = rows[0], rows[1], rows[2:]
if len(sys.argv) > 2:
    print("# " + " ".join(sys.argv[2:]))
for v in vals:
    print(f"## {v[hdr.index('Kernel Name')]}  grid {v[hdr.index('Grid Size')]} block {v[hdr.index('Block Size')]}")
    for h, u, x in sorted(zip(hdr, units, v)):
        if h.startswith(KEEP):
            print(f"{h} [{u}] = {x}")
