"""profiles/r1_results.md from a bench.py JSON line: python scripts/make_results_md.py gpurun_out/bench.json [ref.json] > profiles/r1_results.md"""
import json, sys
d = json.load(open(sys.argv[1]))
ref = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else None
r, e, c = d["roofline"], d["e2e"], d.get("cpu_baseline")
print("# Round 1 measured results (B200, one GPU unless stated; `python bench.py`, default arguments)\n")
print(f"* headline: **{d['value']:.4g} configs/s** resident ({d['config']['workload']}, {d['config']['configs_per_gpu']} configs, "
      f"{d['ms_per_step']:.3f} ms/step), e2e through host buffers **{e['value']:.4g} configs/s** "
      f"({e['h2d_bytes_per_step'] / 1e6:.0f} MB H2D + {e['d2h_bytes_per_step'] / 1e6:.1f} MB D2H per step)")
print(f"* roofline (FP32 SIMT, algorithmic flop): achieved {r['achieved']:.1f} TFLOP/s of {r['peak']:.1f} measured = {r['frac']:.3f}; "
      f"HBM {r['hbm']['achieved']:.1f} GB/s of {r['hbm']['peak']:.0f} = {r['hbm']['frac']:.4f}; DRAM traffic per launch (ncu) "
      f"{(r['traffic'] or 0) / 1e6:.1f} MB vs {r['algorithmic_bytes'] / 1e6:.1f} MB algorithmic")
if r.get("executed"):
    x = r["executed"]
    print(f"* executed FP32 work after culling (ncu): {x['fp32_flop_per_launch'] / d['config']['configs_per_gpu']:.0f} flop/config, "
          f"{x['tflops']:.2f} TFLOP/s = {x['frac_of_peak']:.3f} of the FP32 peak")
if c:
    print(f"* CPU baseline ({c['kind']}, {c['cores']} threads): {c['value']:.4g} configs/s, {c.get('edges_per_s', float('nan')):.4g} edges/s "
          f"-> GPU/CPU = {d['value'] / c['value']:.0f}x resident, {e['value'] / c['value']:.0f}x end to end  [{c['sample']}]")
if ref:
    print(f"* reference arm (`--impl reference`, {ref['cpu_baseline']['cores']} threads, {ref['cpu_baseline']['kind']}): {ref['value']:.4g} configs/s")
print(f"* clocks during the timed region: {d['clocks']}\n")
x = d.get("extra", {})
print("| scene | configs/s | edges/s, uniform endpoints | edges/s, local (±0.2 per joint) | free configs | free uniform edges | free local edges |")
print("|---|---|---|---|---|---|---|")
for k, v in x.items():
    if "config_free_frac" in v:
        print(f"| {k} | {v['configs_per_s']:.4g} | {v['edges_per_s']:.4g} | {v.get('local_edges_per_s', float('nan')):.4g} | {v['config_free_frac']:.3f} | "
              f"{v['edge_free_frac']:.4f} | {v.get('local_edge_free_frac', float('nan')):.3f} |")
a = x.get("abstract_test")
if a:
    print(f"\n* abstract.test (D=4, fp64, bit-exact): {a['configs_per_s']:.4g} configs/s, {a['edges_per_s']:.4g} edges/s at resolution {a['resolution']}"
          + (f"; CPU port on one core: {a['cpu_port_vectorised_configs_per_s_1core']:.4g} configs/s vectorised, "
             f"{a['cpu_port_per_call_configs_per_s_1core']:.4g}/s one call per configuration like the reference's planners; flags identical: {a['flags_identical']}"
             if "flags_identical" in a else ""))
m = x.get("modes_box_rearrangement")
if m:
    print("\n* modes of the dual-arm scene (2 097 152 configs): " + "; ".join(
        f"{k} {v['configs_per_s']:.4g}/s" + (f" ({v['collidable_pairs']} pairs, {v['free_frac']:.3f} free)" if "free_frac" in v else "")
        for k, v in m.items()))
k = x.get("knn_box_stacking_100k")
if k:
    print(f"\n* k-NN {k['N']} x {k['Q']}, D={k['D']}, 4 robots, k={k['k']}, {k['metric']}: tcgen05 path {k['tensor_ms']:.1f} ms "
          f"({k['tensor_queries_per_s']:.4g} queries/s), exact fp64 path {k['exact_ms']:.1f} ms, reference-style numpy loop "
          f"{k.get('cpu_port_queries_per_s_1core', float('nan')):.3g} queries/s on one core")
t = x.get("prm_time_to_first_solution")
if t:
    print("\n| PRM time-to-first-solution (planner.py, same seeds, identical plans) | B200 median s | CPU oracle backend s | ratio |")
    print("|---|---|---|---|")
    for s, v in t.items():
        cpu = v.get("cpu_port_median_s")
        print(f"| {s} ({v.get('modes', '?')} modes, {v['samples_per_mode']} samples/mode) | {v['b200_median_s']:.3f} | {cpu if cpu is None else round(cpu, 2)} | "
              f"{'' if cpu is None else str(round(cpu / v['b200_median_s'])) + 'x'} |")
print("\nMulti-GPU: profiles/r1_scaling.md.  Kernel profiles: profiles/r1_check_configs_v3_*.txt, r1_check_edges_v4.txt.")
