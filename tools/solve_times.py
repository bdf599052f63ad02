"""Wall and device time of N consecutive solves (variance probe): python tools/solve_times.py slab 0 8"""
import sys, time, tempfile
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, hostapi, materials
import torch
d = tempfile.mkdtemp()
mat = hostapi.Material(*materials.write_silicon(d, nw=1000))
ctx = capi.Context(0); ctx.upload_material(mat.desc)
W = {"slab": ("slab", [100e-9] * 3, [100, 0, 0], 10_000_000, 1000), "film": ("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 10_000_000, 100)}
kind, dim, div, n, ms = W[sys.argv[1]]
dom = hostapi.Domain(kind, dim, div, 1.0); prob = hostapi.FieldProblem(mat, dom, "multi", n, ms)
ctx.upload_domain(dom.desc)
raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
for S in (1, int(sys.argv[2]), 1, int(sys.argv[2])):
    ctx.set_options(steps_per_launch=S, slots=0)
    out = []
    for rep in range(int(sys.argv[3])):
        raw.zero_(); torch.cuda.synchronize()
        t = time.perf_counter(); st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=100 + rep); dt = time.perf_counter() - t
        out.append(f"{dt*1e3:.1f}/{st['device_ms']:.1f}/{st['step_ms']:.1f}/{st['launches']}")
    print(f"S={S}: wall/device/k_step ms/launches:", " ".join(out), flush=True)
