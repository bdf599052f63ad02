"""One A/B leg (GPU box): solve the bench workloads with the library named by MCB_LIBMCB and print steady-phase rates.
usage: MCB_LIBMCB=path python tools/ab_run.py [workloads...]   (tools/ab_libs.sh drives it)"""
import os, sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
tag = os.environ.get("AB_TAG", "")
NEMIT = int(os.environ.get("AB_NEMIT") or "4000000")       # enough phonons that launch 21 (ncu -s 20) is a steady one
for wl in (sys.argv[1:] or ["slab", "film", "wire"]):
    if wl == "slab":
        dom = hostapi.Domain("slab", [100e-9] * 3, [100, 0, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 10_000_000, 1000)
    elif wl == "wire":
        dom = hostapi.Domain("wire", [1e-6, 1e-7, 1e-7], [0, 32, 32], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", NEMIT, 100)
    elif wl == "tube":
        dom = hostapi.Domain("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", NEMIT, 100)
    elif wl == "bulk":
        dom = hostapi.Domain("bulk", [1e-6] * 3, [128, 128, 128], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", NEMIT, 100)
    else:
        dom = hostapi.Domain("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", NEMIT, 100)
    ctx.upload_domain(dom.desc)
    raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
    for S, k in ((1, 32), (16, 8)):
        extra = {k_: int(v_) for k_, v_ in (kv.split("=") for kv in os.environ.get("AB_OPTS", "").split())}
        blk = int(os.environ.get("AB_BLOCK") or "768")
        ctx.set_options(steps_per_launch=S, slots=148 * blk * k, **extra)
        best, bs = 1e9, 0.0
        for rep in range(3):
            raw.zero_(); torch.cuda.synchronize()
            t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=rep); best = min(best, time.perf_counter() - t)
            bs = max(bs, st["steady_steps"] / max(st["steady_ms"], 1e-9) * 1e3)
        print(f"  {tag:10s} {wl} S={S:2d}: {st['steps']/best:.3e} steps/s  wall {best*1e3:6.1f} ms  steady {bs:.3e}/s", flush=True)
