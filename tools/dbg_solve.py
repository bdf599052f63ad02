"""Debug helper (GPU box): run the solve parity cases and print per-row errors against the oracle."""
import sys
import numpy as np
sys.path.insert(0, ".")
from tests import cases
from tests.test_gpu_parity import SOLVE_CASES, SEED
import tempfile
from montecarlocpp_b200 import materials
from montecarlocpp_b200 import capi
from oracle import pyoracle as orc

d = tempfile.mkdtemp()
files = materials.write_all(d, nw=1000)
files["silicon_small"] = materials.write_silicon(tempfile.mkdtemp(), nw=64)
omats = {k: orc.Material(*v) for k, v in files.items()}
print("silicon min vel", min(omats["silicon"].desc.vel[i] for i in range(3000)))
ctx = capi.Context(0)
for mname, dname, pkind, size in SOLVE_CASES:
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(ctx, mat, dom)
    ctx.set_options(slots=0, steps_per_launch=0, tally_mode=0)
    prob = orc.Problem(mat, dom, pkind, 20000, 30, size=size)
    ref, rst = prob.solve(rng=orc.RNG_PHILOX, seed=SEED)
    try:
        got, gst = ctx.solve(prob.desc, seed=SEED)
    except Exception as e:
        print(mname, dname, pkind, "ERROR", e); continue
    scale = np.abs(ref).max(axis=1, keepdims=True)
    err = np.abs(got - ref) / np.where(scale > 0, scale, 1)
    print(mname, dname, pkind, "max rel err per row", err.max(axis=1), "scale", scale.ravel(), flush=True)
