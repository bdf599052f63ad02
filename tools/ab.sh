#!/bin/bash
# A/B a compile-time variant on the GPU box: tools/ab.sh "<EXTRA flags>" -> rebuild libmcb.so there and run the sweep
set -e
cd montecarlocpp_b200/csrc && make clean >/dev/null && make EXTRA="$1" 2>&1 | grep -E "rror" || true
grep -A2 "k_stepILi4ELi0ELi0ELb0" ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
cd ../..
python - <<'PY'
import sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
import os
for wl in os.environ.get("AB_WL", "slab film wire").split():
    if wl == "slab":
        dom = hostapi.Domain("slab", [100e-9] * 3, [100, 0, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 10_000_000, 1000)
    elif wl == "wire":
        dom = hostapi.Domain("wire", [1e-6, 1e-7, 1e-7], [0, 32, 32], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 4_000_000, 100)
    elif wl == "bulk":
        dom = hostapi.Domain("bulk", [1e-6] * 3, [128, 128, 128], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 4_000_000, 100)
    else:
        dom = hostapi.Domain("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", 4_000_000, 100)
    ctx.upload_domain(dom.desc)
    raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
    for S, k in ((1, 32), (16, 8)):
        ctx.set_options(steps_per_launch=S, slots=148 * 768 * k)
        best = 1e9
        for rep in range(3):
            raw.zero_(); torch.cuda.synchronize()
            t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=rep); best = min(best, time.perf_counter() - t)
        sm = st["steady_ms"]
        print(f"  {wl} S={S:2d} k={k:2d}: {st['steps']/best:.3e} steps/s  wall {best*1e3:6.1f} ms  steady {st['steady_steps']/max(sm,1e-9)*1e3:.3e}/s ({st['steady_launches']} launches, {sm/max(1,st['steady_launches'])*1e3:.0f} us each)", flush=True)
PY
