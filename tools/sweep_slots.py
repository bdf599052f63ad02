"""Sweep resident-slot count / S for the C2 bench workload (silicon slab, 1e7 phonons)."""
import sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
wl = sys.argv[1] if len(sys.argv) > 1 else "slab"
if wl == "slab":
    dom = hostapi.Domain("slab", [100e-9] * 3, [100, 0, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 10_000_000, 1000)
else:
    dom = hostapi.Domain("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 4_000_000, 100)
ctx = capi.Context(0); ctx.upload_material(mat.desc); ctx.upload_domain(dom.desc)
import torch
raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
for S in (1, 16):
    for k in (4, 8, 16, 32):
        ctx.set_options(steps_per_launch=S, slots=148 * 768 * k)
        for rep in range(2):
            raw.zero_(); torch.cuda.synchronize()
            t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=rep); dt = time.perf_counter() - t
        print(f"{wl} S={S:2d} slots=148*768*{k:2d} wall={dt*1e3:7.1f}ms rate={st['steps']/dt:.3e}/s launches={st['launches']} step_ms={st['step_ms']:.1f} "
              f"stores={st['state_stores']:.3e} GB/s(128B)={st['state_stores']*128/st['step_ms']/1e6:.0f}", flush=True)
