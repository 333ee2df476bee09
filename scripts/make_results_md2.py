"""profiles/r2_results.md from a round-2 bench.py JSON line (top-level blocks):
python scripts/make_results_md2.py profiles/r2_bench_line.json [profiles/r2_ref_line.json] > profiles/r2_results.md"""
import json, sys
d = json.load(open(sys.argv[1]))
ref = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
r, e, c = d["roofline"], d["e2e"], d.get("cpu_baseline")
print("# Round 2 measured results (B200, one GPU; `python bench.py --verbose`, default arguments, final build)\n")
print(f"* headline: **{d['value']:.4g} configs/s** resident ({d['config']['workload']}, {d['ms_per_step']:.3f} ms/step), end to end through host "
      f"buffers **{e['value']:.4g} configs/s** ({e['h2d_bytes_per_step'] / 1e6:.0f} MB H2D + {e['d2h_bytes_per_step'] / 1e6:.1f} MB D2H per step; "
      f"H2D alone {e['h2d_only_gbs_per_rank_min']:.1f} GB/s, the path needs {e['value'] * 48 / 1e9:.1f}); {d['gpu_launches']} kernel launches in the timed region")
x = r["executed"]
print(f"* roofline (FP32 SIMT): algorithmic {r['achieved']:.1f} of {r['peak']:.1f} TFLOP/s = {r['frac']:.3f} (probe {r['peak_probe']:.1f}); executed after culling "
      f"{x['tflops']:.2f} TFLOP/s = {x['frac_of_peak']:.3f}; issue slots {r['issue_active']:.1f} % busy, warp slots {r['warps_active']:.1f} % occupied; "
      f"DRAM {r['traffic'] / 1e6:.1f} MB per launch vs {r['algorithmic_bytes'] / 1e6:.1f} MB algorithmic  [{r['capture']}]")
print(f"* CPU port ({c['cores']} threads): {c['value']:.4g} configs/s -> {d['value'] / c['value']:.0f}x resident, {e['value'] / c['value']:.0f}x end to end")
if ref:
    print(f"* reference arm (`--impl reference`, {ref['cpu_baseline']['cores']} threads, {ref['cpu_baseline']['kind']}): {ref['value']:.4g} configs/s")
print(f"* clocks during the timed region: {d['clocks']}\n")
print("| mode (dual-arm scene) | configs/s, uniform samples | free | configs/s, all-free batch |\n|---|---|---|---|")
for k, v in d["modes"].items():
    print(f"| {k} | {v['configs_per_s']:.4g} | {v.get('free_frac', float('nan')):.3f} | {v.get('all_free_batch_configs_per_s', float('nan')):.4g} |")
print("\n| scene | configs/s | free | edges/s uniform (W_edge frac) | edges/s local (W_edge frac) |\n|---|---|---|---|---|")
for k, v in d["scenes"].items():
    u, l = v.get("edges_uniform", {}), v.get("edges_local", {})
    print(f"| {k} | {v['configs_per_s']:.4g} | {v['config_free_frac']:.3f} | {u.get('edges_per_s', float('nan')):.4g} ({u.get('roofline', {}).get('frac', float('nan')):.2f}) | "
          f"{l.get('edges_per_s', float('nan')):.4g} ({l.get('roofline', {}).get('frac', float('nan')):.2f}) |")
ed = d["edges"]
print(f"\nHeadline scene edges: uniform {ed['uniform']['edges_per_s']:.4g}/s (W_edge frac {ed['uniform']['roofline']['frac']:.2f}), local "
      f"{ed['local']['edges_per_s']:.4g}/s (frac {ed['local']['roofline']['frac']:.2f}); CPU port {ed['uniform'].get('cpu_port_edges_per_s', float('nan')):.3g} / "
      f"{ed['local'].get('cpu_port_edges_per_s', float('nan')):.3g} edges/s.")
k = d["knn"]
print(f"\nk-NN {k['N']} x {k['Q']}, D = {k['D']}, k = {k['k']}, {k['metric']}: tensor path {k['tensor_ms']:.2f} ms, exact {k['exact_ms']:.1f} ms, identical: "
      f"{k['tensor_equals_exact']}; {k['roofline']['achieved']:.1f} TFLOP/s algorithmic = {k['roofline']['frac']:.3f} of the TF32 ceiling, tensor pipe "
      f"{k['roofline']['tensor_pipe_active_pct']:.1f} % busy; r-disc (selective radius, {k['radius']['mean_neighbours']:.0f} neighbours per row): {k['radius']['ms']:.2f} ms.")
print("\n| reference planner (four-arm stacking problem) | first solution CUDA | CPU port | ratio | backend time CUDA / CPU | device round trips |\n|---|---|---|---|---|---|")
for name, v in d["planners"].items():
    if not isinstance(v, dict) or "cuda" not in v:
        continue
    cu, cp = v["cuda"], v["cpu"]
    print(f"| {name} | {cu['median_ttfs_s']:.3f} s | {cp['median_ttfs_s']:.2f} s | {v['ttfs_ratio_cpu_over_cuda']:.1f}x | {cu['median_backend_s']:.2f} / {cp['median_backend_s']:.2f} s = "
          f"{v['backend_time_ratio_cpu_over_cuda']:.0f}x | {cu['median_device_round_trips']:.0f} |")
b = d["batched_prm"]
print(f"\nBatch-native PRM ({b['scene']}, {b['samples_per_mode']} samples per mode, {b['modes']} modes): {b['b200_median_s']:.3f} s against {b['cpu_port_s']:.2f} s on "
      f"{b['cpu_cores']} CPU threads, same plan: {b['same_plan']}.")
a = d.get("abstract_test")
if a:
    print(f"\nabstract.test (BASELINE config 1): {json.dumps(a)[:900]}")
