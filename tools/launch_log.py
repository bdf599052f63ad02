"""Per-launch trace of one solve (MCB_LAUNCH_LOG=1): python tools/launch_log.py slab 0|1   (second argument: steps_per_launch, 0 = library default)"""
import os, sys, tempfile
os.environ["MCB_LAUNCH_LOG"] = "1"
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
W = {"slab": ("slab", [100e-9] * 3, [100, 0, 0], 10_000_000, 1000), "film": ("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 10_000_000, 100),
     "wire": ("wire", [1e-6, 1e-7, 1e-7], [0, 32, 32], 20_000_000, 100)}
kind, dim, div, n, ms = W[sys.argv[1]]
dom = hostapi.Domain(kind, dim, div, 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", n, ms)
ctx.upload_domain(dom.desc)
raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
ctx.set_options(steps_per_launch=int(sys.argv[2]), slots=0)
ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=0)   # warm-up (also traced)
st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=1)
print({k: st[k] for k in ("steps", "launches", "device_ms", "step_ms", "steady_ms", "tail_ms", "tail_steps")})
