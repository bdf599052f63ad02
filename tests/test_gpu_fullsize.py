"""GPU tests at BASELINE.json's full problem sizes, through size-independent properties of the domain
(the oracle cannot run these sizes in seconds): particle/step bookkeeping, zero escapes, flux continuity,
mirror symmetries, bulk conductivity.  Configs C2-C5 as pinned in SURVEY.md §8d."""
import numpy as np
import pytest

from tests import cases

pytestmark = pytest.mark.gpu
SEED = 0xC0FFEE


@pytest.fixture(scope="module")
def hmats(matfiles):
    from montecarlocpp_b200 import hostapi
    return {k: hostapi.Material(*matfiles[k]) for k in ("grey", "silicon")}


def _solve(ctx, mat, dom, kind, nemit, maxscat, **opts):
    from montecarlocpp_b200 import hostapi
    prob = hostapi.FieldProblem(mat, dom, kind, nemit, maxscat)
    ctx.upload_material(mat.desc); ctx.upload_domain(dom.desc)
    ctx.set_options(slots=0, steps_per_launch=0, block=0, ctas_per_sm=0, tally_mode=0, **opts)
    sol, st = ctx.solve(prob.desc, seed=SEED)
    assert st["emitted"] == prob.nemit
    return prob, sol, st


@pytest.mark.parametrize("mname", ["grey", "silicon"])
def test_c2_slab_1e7_flux_continuity_and_antisymmetry(gpu_ctx, hmats, mname):
    """C2: 100 nm slab between +-0.5 K walls, 1e7 phonons, gray vs full dispersion.  At steady state the heat flux is
    the same through every cell, the temperature profile is antisymmetric about the mid-plane and monotone."""
    from montecarlocpp_b200 import hostapi
    dom = hostapi.Domain("slab", [100e-9] * 3, [100, 0, 0], 1.0)
    prob, sol, st = _solve(gpu_ctx, hmats[mname], dom, "multi", 10_000_000, 1000)
    assert st["esc"] == 0 and st["steps"] > prob.nemit
    T, qx = sol[0], sol[1]
    assert qx.min() > 0 and (qx.max() - qx.min()) / qx.mean() < 0.02
    assert np.abs(T + T[::-1]).max() < 0.03 * np.abs(T).max()
    coarse = T.reshape(10, 10).mean(1)
    assert (np.diff(coarse) < 0).all()
    assert np.abs(sol[2]).max() < 0.02 * qx.mean() and np.abs(sol[3]).max() < 0.02 * qx.mean()   # no transverse flux


def test_c3_wire_1e8_symmetry(gpu_ctx, hmats):
    """C3: 1 um x 100 nm x 100 nm wire with four diffuse walls, 32x32 cross-section tally (2-D walk), 1e8 phonons."""
    from montecarlocpp_b200 import hostapi
    dom = hostapi.Domain("wire", [1e-6, 100e-9, 100e-9], [0, 32, 32], 1.0)
    prob, sol, st = _solve(gpu_ctx, hmats["silicon"], dom, "multi", 100_000_000, 100)
    assert st["esc"] == 0 and st["cols"] == 1024
    qx = sol[1].reshape(32, 32)                       # [k (z), j (y)]
    assert qx.min() > 0
    sym = 0.25 * (qx + qx.T + qx[::-1, ::-1] + qx[::-1, ::-1].T)
    assert np.abs(qx - sym).max() < 0.05 * qx.mean()  # y<->z and mirror symmetry of the square cross-section
    assert qx[12:20, 12:20].mean() > qx[0, :].mean()  # diffuse walls suppress the flux near the surface


def test_c4_tube_1e8_no_failed_handoffs(gpu_ctx, hmats):
    """C4: TubeDomain (3 boxes, 2 Inter pairs, 3 Peri pairs), 1e8 phonons: every Inter hand-off finds its partner."""
    from montecarlocpp_b200 import hostapi
    dom = hostapi.Domain("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0)
    prob, sol, st = _solve(gpu_ctx, hmats["silicon"], dom, "multi", 100_000_000, 100)
    assert st["esc"] == 0
    assert st["cols"] == 4 * 8 + 4 * 4 + 8 * 4
    # subdomains 0 (y-arm) and 2 (z-arm) are mirror images about the tube diagonal
    q0 = sol[1][:32].reshape(8, 4)                    # sdom0 shape (1, 4, 8): [k, j]
    q2 = sol[1][48:].reshape(4, 8)                    # sdom2 shape (1, 8, 4): [k, j]
    assert np.abs(q0 - q2.T).max() < 0.05 * q0.mean()
    assert (sol[1] > 0).all()


def test_c5_bulk_128cubed_global_tally(gpu_ctx, hmats):
    """C5 (one GPU's share): 1 um bulk cube, 128^3 cells -> 67 MB field, fp64 RED straight to L2, 3-D walk."""
    from montecarlocpp_b200 import hostapi
    dom = hostapi.Domain("bulk", [1e-6] * 3, [128, 128, 128], 1.0)
    mat = hmats["grey"]
    prob, sol, st = _solve(gpu_ctx, mat, dom, "multi", 20_000_000, 20)
    assert st["esc"] == 0 and st["cols"] == 128**3
    k = sol[1].mean() / 1e6
    assert abs(k / mat.cond() - 1) < 0.01
    assert abs(sol[2].mean()) < 0.01 * sol[1].mean() and abs(sol[3].mean()) < 0.01 * sol[1].mean()
