"""Quick GPU probe: time the C2-like slab / film workloads at a few schedule settings."""
import sys, time, tempfile
import numpy as np
sys.path.insert(0, ".")
from montecarlocpp_b200 import capi, materials
from oracle import pyoracle as orc
from tests import cases

d = tempfile.mkdtemp()
files = materials.write_all(d)
ctx = capi.Context(0)
for mname in ("grey", "silicon"):
    mat = orc.Material(*files[mname])
    for dname, dom in (("slab", cases.slab(ncell=100)), ("film", cases.film())):
        cases.upload(ctx, mat, dom)
        prob = orc.Problem(mat, dom, "multi", int(sys.argv[1]) if len(sys.argv) > 1 else 2000000, 100)
        for S in (1, 16):
            for tm in (1, 3):
                ctx.set_options(steps_per_launch=S, tally_mode=tm)
                t = time.time(); sol, st = ctx.solve(prob.desc, seed=1); dt = time.time() - t
                print(f"{mname:8s} {dname:5s} S={S:3d} tally={tm} steps={st['steps']:.3e} wall={dt*1e3:8.1f}ms dev={st['device_ms']:8.1f}ms "
                      f"step_ms={st['step_ms']:8.1f} launches={st['launches']} rate={st['steps']/st['device_ms']*1e3:.3e}/s "
                      f"slot_eff={st['steps']/max(st['slot_steps'],1):.2f} esc={st['esc']}", flush=True)
