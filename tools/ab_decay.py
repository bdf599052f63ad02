"""Decay-phase schedule A/B in ONE process (GPU box): whole solves at the bench sizes for a list of (decay_mode, decay_pct)
variants, interleaved.  usage: python tools/ab_decay.py "2:0 0:8 0:12 0:20" slab film wire   (AB_S=1 streaming, AB_S=0 library default)"""
import os, sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
W = {"slab": ("slab", [100e-9] * 3, [100, 0, 0], 10_000_000, 1000), "film": ("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 10_000_000, 100),
     "wire": ("wire", [1e-6, 1e-7, 1e-7], [0, 32, 32], 20_000_000, 100), "tube": ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 20_000_000, 100),
     "bulk": ("bulk", [1e-6] * 3, [128, 128, 128], 20_000_000, 100)}
variants = [tuple(int(x) for x in v.split(":")) for v in sys.argv[1].split()]
S = int(os.environ.get("AB_S") or "1")
for wl in sys.argv[2:]:
    kind, dim, div, n, ms = W[wl]
    dom = hostapi.Domain(kind, dim, div, 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", n, ms)
    ctx.upload_domain(dom.desc)
    raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
    best = {v: 1e9 for v in variants}; stv = {}; ref = None
    for rep in range(4):
        for v in variants:
            ctx.set_options(steps_per_launch=S, slots=0, decay_mode=v[0], decay_pct=v[1])
            raw.zero_(); torch.cuda.synchronize()
            t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=7); dt = time.perf_counter() - t
            if rep: best[v] = min(best[v], dt)
            stv[v] = st
            if rep == 1:
                f = raw.clone()
                if ref is None: ref = (f, st)
                else:
                    sc = ref[0].abs().max().item()
                    assert (st["steps"], st["esc"], st["emitted"]) == (ref[1]["steps"], ref[1]["esc"], ref[1]["emitted"]), (v, st, ref[1])
                    assert (f - ref[0]).abs().max().item() <= 1e-9 * sc, (v, (f - ref[0]).abs().max().item() / sc)
    for v in variants:
        st = stv[v]
        print(f"  S={S} {wl:5s} decay_mode {v[0]} pct {v[1]:2d}: {st['steps']/best[v]:.3e} steps/s  {best[v]*1e3:7.1f} ms  launches {st['launches']:4d}"
              f"  k_step {st['step_ms']:.1f} ms  lane occupancy {st['steps']/max(st['slot_steps'],1):.3f}  compactions {st['compactions']}", flush=True)
