"""Turn ncu outputs into the committed summaries under profiles/.
  python tools/profile_summary.py launches <launches.csv> <out.md> "<command>"
  python tools/profile_summary.py kernel <report.ncu-rep> <out.md> "<command>"
"""
import collections, csv, io, subprocess, sys


def launches(path, out, cmd):
    rows = list(csv.reader(open(path)))
    h = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[h]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    agg = collections.OrderedDict(); seq = []
    for r in rows[h + 1:]:
        if len(r) <= vi: continue
        name = r[ki].split("(")[0].replace("void ", ""); v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v * 1e3 if r[ui] in ("ms", "msecond") else v)
        a = agg.setdefault(name, [0, 0.0, r[gi], r[bi]]); a[0] += 1; a[1] += v; seq.append((name, v))
    tot = sum(a[1] for a in agg.values())
    with open(out, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none)\n\n`{cmd}`\n\n")
        f.write("Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.\n\n")
        f.write("| kernel | launches | total us | share | grid | block |\n|---|---:|---:|---:|---|---|\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {a[0]} | {a[1]:.1f} | {100 * a[1] / tot:.2f} % | {a[2]} | {a[3]} |\n")
        ks = [v for n, v in seq if "k_step" in n]
        if ks:
            f.write(f"\nk_step launch durations (us), in launch order, every 8th: " + " ".join(f"{v:.0f}" for v in ks[::8]) + "\n")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max"]


def kernel(rep, out, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw))); hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        f.write(f"# ncu --set full summary\n\n`{cmd}`\n\nreport: `{rep}` (scratch, not committed)\n")
        for r in rows[2:]:
            f.write(f"\n## {r[hdr.index('Kernel Name')]}  (launch id {r[0]})\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr: f.write(f"| {k} | {r[hdr.index(k)]} | {units[hdr.index(k)]} |\n")
            try:
                rd = float(r[hdr.index("dram__bytes_read.sum")].replace(",", "")); wr = float(r[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                u = units[hdr.index("dram__bytes_read.sum")]; mul = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}[u]
                t = float(r[hdr.index("gpu__time_duration.sum")].replace(",", "")); tu = units[hdr.index("gpu__time_duration.sum")]
                t *= {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9}[tu]
                f.write(f"\nDRAM traffic {(rd + wr) * mul / 1e6:.1f} MB per launch -> {(rd + wr) * mul / t / 1e9:.0f} GB/s over {t * 1e6:.0f} us\n")
            except Exception as e:
                f.write(f"\n(traffic summary unavailable: {e})\n")
            names = [h for h in hdr if "average_warps_issue_stalled" in h and h.endswith("_per_issue_active.ratio")]
            vals = sorted([(float(r[hdr.index(n)].replace(",", "") or 0), n) for n in names], reverse=True)[:8]
            f.write("\nwarp stall reasons (warps per issue-active cycle): " + ", ".join(
                f"{n.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')} {v:.2f}" for v, n in vals) + "\n")
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name", "regex:k_step"],
                             capture_output=True, text=True).stdout
        open("/tmp/_src.csv", "w").write(src)
        lines = subprocess.run([sys.executable, "tools/ncu_lines.py", "/tmp/_src.csv", "25"], capture_output=True, text=True).stdout
        f.write("\n## hottest CUDA source lines (warp-state samples)\n\n```\n" + lines + "```\n")
        sass = subprocess.run(["cuobjdump", "-sass", "montecarlocpp_b200/libmcb.so"], capture_output=True, text=True).stdout
        f.write(f"\nSASS evidence in libmcb.so: UBLKCP (TMA bulk copy) x{sass.count('UBLKCP')}, SYNCS (mbarrier) x{sass.count('SYNCS.')}, "
                f"MATCH.ANY x{sass.count('MATCH.ANY')}, REDG.E.ADD.F64 x{sass.count('REDG.E.ADD.F64')}, tensor-core ops (UTC*MMA/HMMA) x{sass.count('MMA')}\n")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4] if len(sys.argv) > 4 else "")
