"""A small solve of every kernel family (1-D warp histograms, CTA histogram, global tally, N-D walk, non-box cells are covered by
the tests; this is the quick one) for compute-sanitizer:  compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from montecarlocpp_b200 import capi, hostapi, materials
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=64))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
for kind, dim, div, pk, n in (("slab", [1e-7] * 3, [20, 0, 0], "multi", 6000), ("film", [1e-6, 1e-7, 1e-6], [0, 10, 0], "temp", 6000),
                              ("wire", [1e-6, 1e-7, 1e-7], [0, 8, 8], "multi", 4000), ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 4, 4, 2], "cumflux", 4000),
                              ("bulk", [1e-6] * 3, [64, 4, 4], "flux", 3000)):
    dom = hostapi.Domain(kind, dim, div, 1.0)
    prob = hostapi.FieldProblem(mat, dom, pk, n, 30, size=3 if pk.startswith("cum") else 0)
    ctx.upload_domain(dom.desc)
    for opts in (dict(steps_per_launch=1, slots=2048), dict(steps_per_launch=8, slots=1024)):
        ctx.set_options(**opts)
        sol, st = ctx.solve(prob.desc, seed=11)
        print(kind, pk, opts, st["steps"], st["esc"], flush=True)
ctx.close()
