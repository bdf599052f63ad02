"""profiles/ncu_traffic.json from the ncu --set full captures of the steady k_step launch of each workload
(gpurun_out/prof_r2_final_<wl>.ncu-rep, tools/gpu_r2_final.sh): DRAM bytes per launch, duration, issue-slot utilisation,
warp instructions per warp-step.  bench.py reads the file for roofline.traffic (ncu cannot run inside the bench).
usage: python tools/ncu_traffic.py [tag]      (tag: the capture name infix, default r2_final)"""
import csv, io, json, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r2_final"
KEY = {"slab": "C2-slab100nm-si", "film": "C1-film100nm-si", "wire": "C3-wire32x32-si", "tube": "C4-tube-si", "bulk": "C5-bulk128-si"}
out = {"source": f"ncu --set full --clock-control none, one steady (S=1, full population) k_step launch per workload: gpurun_out/prof_{tag}_<wl>.ncu-rep "
                 f"(summaries: profiles/r2_k_step_<wl>_S1.md); python tools/ab_run.py <wl> -s 20 -c 1", "kernels": {}}
for wl, key in KEY.items():
    raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{tag}_{wl}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    if len(rows) < 3:
        continue
    hdr, units, r = rows[0], rows[1], rows[2]
    def val(name):
        i = hdr.index(name); v = float(r[i].replace(",", "")); u = units[i]
        return v * {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}.get(u, 1)
    b = val("dram__bytes_read.sum") + val("dram__bytes_write.sum"); t = val("gpu__time_duration.sum")
    out["kernels"][key] = {"kernel": r[hdr.index("Kernel Name")], "dram_bytes_per_launch": b, "us": t * 1e6, "gbs": b / t / 1e9,
                           "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                           "warp_instructions": val("smsp__inst_executed.sum"), "registers": val("launch__registers_per_thread"),
                           "block": val("launch__block_size"), "smem_dynamic_bytes": val("launch__shared_mem_per_block_dynamic"),
                           "lanes_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio")}
    k = out["kernels"][key]
    k["slots"] = 148 * int(k["block"]) * 32                     # tools/ab_run.py: 32 tiles per warp in the captured launch
    k["dram_bytes_per_slot"] = k["dram_bytes_per_launch"] / k["slots"]
    k["warp_instructions_per_warp_step"] = k["warp_instructions"] / (k["slots"] / 32)
json.dump(out, open("profiles/ncu_traffic.json", "w"), indent=1)
print(json.dumps(out, indent=1))
