"""Markdown rows of DESIGN.md section 7 from a bench line and the ncu table: python tools/design_table.py profiles/bench_r2_1gpu.json"""
import json, sys
d = json.load(open(sys.argv[1])); t = json.load(open("profiles/ncu_traffic.json"))["kernels"]
rows = [(d["config"]["workload"], d)] + list(d.get("configs", {}).items())
for k, v in rows:
    r, n = v["roofline"], t.get(k, {})
    e2e = v["e2e"]["value"] if "e2e" in v else v.get("e2e_value", 0.0)
    res = v["resident_mode"]["value_per_gpu"] if "resident_mode" in v else v.get("resident_mode_value_per_gpu", 0.0)
    print(f"| {k} | {v['value']:.3g} (e2e {e2e:.3g}) | {res:.3g} | {v['ms_per_step']:.1f} | {r['frac']:.3f} ({r['phonon_steps_per_s_in_kernel']:.3g} in the kernel, "
          f"{1e3 * r['kernel_ms_per_launch']:.0f} us / launch) | {r['kernel_share_of_step']:.2f} | ncu: {n.get('dram_bytes_per_slot', 0):.0f} B per slot, "
          f"issue {n.get('issue_active_pct', 0):.0f} %, {n.get('warp_instructions_per_warp_step', 0):.0f} warp-instr per warp-step, {n.get('registers', 0):.0f} regs x {n.get('block', 0):.0f} |")
