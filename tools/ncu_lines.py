"""Aggregate an ncu source page (cuda,sass view) by CUDA source line: samples, instructions, thread efficiency.
usage: ncu -i X.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_step > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, func, hdr, agg, seen_func = None, None, None, {}, set()
first_func = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]
        if first_func is None: first_func = func
        continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or func != first_func: continue
    if r[0].isdigit():
        j, ie, te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        k = (fname, int(r[0]))
        a = agg.setdefault(k, [0, 0, 0, r[1]])
        for idx, col in enumerate((j, ie, te)):
            try: a[idx] += int(r[col])
            except ValueError: pass
tot = sum(a[0] for a in agg.values()); ti = sum(a[1] for a in agg.values())
print(f"kernel {first_func}: samples {tot}, warp instructions {ti}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    eff = a[2] / a[1] / 32 if a[1] else 0
    print(f"{a[0]:6d} {100*a[0]/tot:5.1f}% inst {100*a[1]/ti:5.1f}% lanes {eff:4.2f}  {f}:{ln:<4d} {a[3].strip()[:105]}")
