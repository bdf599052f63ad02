"""Whole-solve A/B of library variants on the bench workloads (full BASELINE sizes): MCB_LIBMCB=lib python tools/ab_whole.py slab film ..."""
import os, sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
tag = os.environ.get("AB_TAG", "")
W = {"slab": ("slab", [100e-9] * 3, [100, 0, 0], 10_000_000, 1000), "film": ("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 10_000_000, 100),
     "wire": ("wire", [1e-6, 1e-7, 1e-7], [0, 32, 32], 20_000_000, 100), "tube": ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 20_000_000, 100),
     "bulk": ("bulk", [1e-6] * 3, [128, 128, 128], 20_000_000, 100)}
for wl in (sys.argv[1:] or ["slab", "film"]):
    kind, dim, div, n, ms = W[wl]
    dom = hostapi.Domain(kind, dim, div, 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", n, ms)
    ctx.upload_domain(dom.desc)
    raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
    tiles = int(os.environ.get("AB_TILES") or "0")            # tiles per warp (0: library default, 32)
    blk = 768 if wl in ("slab", "film") else 640
    ctx.set_options(steps_per_launch=int(os.environ.get("AB_S") or "1"), slots=148 * blk * tiles, compact_pct=int(os.environ.get("AB_COMPACT") or "0"),
                    sort_mode=int(os.environ.get("AB_SORT") or "0"))
    best = 1e9
    for rep in range(4):
        raw.zero_(); torch.cuda.synchronize()
        t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=rep); dt = time.perf_counter() - t
        if rep: best = min(best, dt)
    print(f"  {tag:10s} {wl}: whole solve {st['steps']/best:.3e} steps/s  {best*1e3:7.1f} ms  launches {st['launches']}  k_step {st['step_ms']:.1f} ms (steady {st['steady_ms']:.1f})"
          f"  compactions {st['compactions']} sorts {st['sorts']}", flush=True)
