"""Group an ncu source page (cuda,sass csv) of k_step by the function a source line belongs to.
usage: python tools/ncu_groups.py src.csv"""
import csv, sys, collections, re, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def starts(path):
    out = []
    for n, ln in enumerate(open(path), 1):
        m = re.match(r"\s*(?:template.*)?(?:__device__|__global__)[^(]*?(\w+)\(", ln)
        if m: out.append((n, m.group(1)))
        m = re.match(r"struct (\w+)", ln)
        if m: out.append((n, "struct " + m.group(1)))
    return out
S = {f: starts(os.path.join(ROOT, "montecarlocpp_b200/csrc", f)) for f in ("mcb_device.cuh", "mcb_kernels.cuh")}
def group(f, l):
    if f not in S: return f
    name = "?"
    for n, nm in S[f]:
        if n <= l: name = nm
        else: break
    return f.split("_")[1][0] + ":" + name
rows = list(csv.reader(open(sys.argv[1])))
fname = hdr = func = first = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        func = r[1]; first = first or func; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or func != first: continue
    if r[0].isdigit():
        j, ie, te = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        a = agg[group(fname, int(r[0]))]
        for idx, col in enumerate((j, ie, te)):
            try: a[idx] += int(r[col])
            except ValueError: pass
ti = sum(a[1] for a in agg.values()); ts = sum(a[0] for a in agg.values())
nws = float(sys.argv[2]) if len(sys.argv) > 2 else 113664.0
print(f"{first}: samples {ts}, warp instructions {ti} ({ti/nws:.0f} per warp-step)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:34s} inst {100*v[1]/ti:5.1f}% ({v[1]/nws:7.1f}/warp-step) lanes {v[2]/max(v[1],1)/32:4.2f} samples {100*v[0]/ts:5.1f}%")
