"""Instruction budget of k_step by function group and the per-line difference between two captures, from ncu source pages.
  ncu -i A.ncu-rep --page source --print-source cuda,sass --csv --kernel-name regex:k_step > a.csv   (same for B)
  python tools/ncu_budget.py a.csv <warp-steps in A> b.csv <warp-steps in B>
(warp-steps = 32-slot tiles x loop trips of the captured launch)"""
import csv, sys


def load(path):
    rows = list(csv.reader(open(path)))
    fname = func = hdr = first = None; agg = {}
    for r in rows:
        if not r: continue
        if r[0] == "File Path": fname = r[1].split("/")[-1]; continue
        if r[0] == "Function Name":
            func = r[1]; first = first or func; continue
        if r[0] == "Line No": hdr = r; continue
        if hdr is None or func != first or not r[0].isdigit(): continue
        ie, te, j = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
        a = agg.setdefault((fname, int(r[0])), [0, 0, 0, r[1]])
        for idx, col in enumerate((ie, te, j)):
            try: a[idx] += int(r[col])
            except ValueError: pass
    return agg


def group(src):
    """function group of a source line, from its text (line numbers move; the statements do not)"""
    t = src.strip()
    keys = [("philox", ("__umulhi(0x", "rk[2 * r]", "^ c1 ^")), ("reciprocal / division (Newton)", ("fma(-b, y, 1.0)", "rcp_fast(b)", "fma(-b, q, a)", "rcp.approx")),
            ("log (free path)", ("c_k[19]", "c_k[17]", "c_k[18]", "hfsq", "c_k[23] - (double)x", "div_fast(f, 2.0 + f)", "__hiloint2double(hi")),
            ("sincospi", ("c_k[0], z", "c_k[6], z", "c_k[24]", "swap ?", "<< 30")),
            ("intrinsic event draw", ("inv_bucket", "T.wprob[r]", "T.pprob[k]", "T.lambda[wp]", "sth", "o.wp = wp", "draw_event(")),
            ("advect + move", ("sd.offl[b]", "t < d ||", "key < 6", "sn < c_k[27]", "sg.ex = sg.bx", "ph.ps++", "div_fast(num", "advect_move<", "double d = ph.sn")),
            ("isInside", ("s0 = x + sd.offl", "setp.lt", "is_inside<")),
            ("1-D tally set-up", ("t1_", "__double2int_rd(bcd)", "t.w0 =", "t.w1 =", "t.clo", "t.np =", "flag < -1", "fabs(ecd - bcd)", "t.c1 =")),
            ("1-D tally deposits", ("red.shared.add.u32", "MCB_MAGIC", "__double2loint(s)", "deposit_fx<", "base[r] = amt[r]", "t.np ==", "t.np >=")),
            ("payload (accumAmt)", ("fx_scale", "T.inv_vel[ph.wp()]", "sg.ex - sg.bx", "rbase", "slow")),
            ("wall events / stop test", ("MCB_BDRY", "cb.", "renorm", "ph.kill", "maxscat32", "collide<", "n2 - 1.0", "0.5 * n2", "h.nx")),
            ("tile: state load / store / prefetch", ("lds2(", "st_stream", "ld_stream", "tma_bulk", "mbar_", "store_group", "ph.store(", "load_shared", "a_buf", "a_bar", "MCB_GROUP_BYTES")),
            ("tile: free list + counters + ballots", ("my_free", "s_wcnt", "wc.", "__nvvm_vote", "ballot", "fm", "nlisted", "__popc", "__nvvm_bar_warp", "__syncwarp", "__shfl", "shfl")),
            ("tile: loop control / flush / invariants", ("for (int g =", "for (int s = 0", "since_flush", "threadIdx.x & 31u", "gstride", "ngroups", "T.", "P.", "const int i = g"))]
    for name, pats in keys:
        if any(p in t for p in pats): return name
    return "other"


a, wa, b, wb = load(sys.argv[1]), float(sys.argv[2]), load(sys.argv[3]), float(sys.argv[4])
ta, tb = sum(v[0] for v in a.values()), sum(v[0] for v in b.values())
print(f"A: {ta / wa:.0f} warp-instructions per warp-step, B: {tb / wb:.0f}\n")
ga, gb = {}, {}
for agg, g, w in ((a, ga, wa), (b, gb, wb)):
    for k, v in agg.items():
        e = g.setdefault(group(v[3]), [0.0, 0.0]); e[0] += v[0] / w; e[1] += v[1] / w
print("| group | A: instructions per warp-step | lanes | B: instructions per warp-step | lanes |\n|---|---:|---:|---:|---:|")
for name in sorted(set(ga) | set(gb), key=lambda n: -(gb.get(n, [0])[0])):
    x, y = ga.get(name, [0, 0]), gb.get(name, [0, 0])
    print(f"| {name} | {x[0]:.0f} | {x[1] / max(x[0], 1e-9) / 32:.2f} | {y[0]:.0f} | {y[1] / max(y[0], 1e-9) / 32:.2f} |")
