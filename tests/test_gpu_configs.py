"""Parity of the BASELINE configurations at reduced N, against the REFERENCE ITSELF (oracle/_ref/ref_driver:
the reference's FieldProblem::solve objects, mt19937) and, for the 128^3 grid, against the oracle in Philox mode.

Bars (north star): per-cell temperature / heat-flux within 3 sigma of the batch-means error (>= 90 % of the cells inside 3 sigma and the worst of the
hundreds of cells tested at once inside the Student-t quantile of a 0.1 % family-wise false-alarm rate), domain-mean flux within 3 sigma of ITS
batch-means error and, wherever the reference sample resolves it (sigma < 0.33 %), within 1 %.  The CPU reference bounds the
sample: the GPU side runs 10x the phonons, so the error bar is the reference's.  KA1 (bulk conductivity) is asserted at 1 % for
both materials."""
import numpy as np
import pytest

from montecarlocpp_b200 import hostapi
from oracle import pyoracle as orc
from oracle import refbin
from tests import cases

pytestmark = pytest.mark.gpu
SEED = 0x5EED0000

# name: (domain kind, dim, div, dT, maxscat, reference phonons per batch)
CONFIGS = {
    "C2-slab-si": ("slab", [100e-9] * 3, [100, 0, 0], 1.0, 1000, 200_000),
    "C1-film-si": ("film", [1e-6, 100e-9, 1e-6], [0, 20, 0], 1.0, 100, 1_000_000),
    "C3-wire32-si": ("wire", [1e-6, 100e-9, 100e-9], [0, 32, 32], 1.0, 100, 1_000_000),
    "C4-tube-si": ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0, 100, 1_000_000),
    "C5-bulk16-si": ("bulk", [1e-6] * 3, [16, 16, 16], 1.0, 100, 1_000_000),
}


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("name", sorted(CONFIGS))
def test_config_profiles_within_three_sigma_of_the_reference(matfiles, name):
    kind, dim, div, dT, maxscat, nref = CONFIGS[name]
    disp, relax = matfiles["silicon"]
    B = 8
    mat = hostapi.Material(disp, relax, 300.0)
    dom = hostapi.Domain(kind, dim, div, dT)
    prob = hostapi.FieldProblem(mat, dom, "multi", 10 * nref, maxscat)
    hostapi.set_devices([0])
    g, esc = [], 0
    for b in range(B):
        sol, st = prob.solve_seeded(SEED + 900 + b)
        g.append(sol); esc += st["esc"]
    import os
    r = [refbin.drive(disp, relax, 300.0, kind, dim, div, dT, "multi", nref, maxscat, seed=5000 + 64 * b, threads=os.cpu_count() or 4) for b in range(B)]
    assert esc == 0 and sum(x["esc"] for x in r) == 0
    g, r = np.stack(g), np.stack([x["output"] for x in r])
    assert g.shape == r.shape
    vg, vr = g.var(0, ddof=1)[:2] / B, r.var(0, ddof=1)[:2] / B                        # T and q_x rows
    se = np.sqrt(vg + vr)
    z = np.abs(g.mean(0) - r.mean(0))[:2] / np.where(se > 0, se, 1.0)
    # with 8 batches per side the statistic is Student-t (Welch), not normal: the worst of several hundred cells is held to the
    # t quantile of a 0.1 % family-wise false-alarm rate (about 5 - 7 "sigma" at 8 - 14 degrees of freedom); 90 % inside 3 sigma
    from scipy import stats
    dof = np.where(se > 0, (vg + vr) ** 2 / np.maximum(vg ** 2 / (B - 1) + vr ** 2 / (B - 1), 1e-300), 2 * B - 2)
    cap = stats.t.ppf(1.0 - 0.5e-3 / z.size, np.maximum(dof, 1.0))
    assert (z < cap).all() and (z < 3.0).mean() >= 0.9, (name, z.max(), (z < 3.0).mean())
    qg, qr = g[:, 1].mean(axis=1), r[:, 1].mean(axis=1)                                # domain-mean q_x per batch
    sig = np.sqrt(qg.var(ddof=1) / B + qr.var(ddof=1) / B)
    assert abs(qg.mean() - qr.mean()) <= 3.0 * sig, (name, qg.mean(), qr.mean(), sig)
    if sig < 0.0033 * abs(qr.mean()):
        assert abs(qg.mean() / qr.mean() - 1.0) < 0.01, (name, qg.mean() / qr.mean())


@pytest.mark.parametrize("mname,nemit", [("grey", 64_000_000), ("silicon", 1_000_000_000)])
def test_bulk_conductivity_within_one_percent_both_materials(matfiles, mname, nemit):
    """KA1 at the stated bar for BOTH materials: <q_x> / |grad T| -> Material::cond() (material.cpp:160-161) within 1 %.  The
    synthetic silicon's heavy-tailed free paths need ~1e9 first flights (maxscat = 1: later flights are isotropic, zero-mean
    noise) -- about a second on the device."""
    mat = hostapi.Material(*matfiles[mname])
    dom = hostapi.Domain("bulk", [1e-6] * 3, [10, 0, 0], 1.0)
    prob = hostapi.FieldProblem(mat, dom, "flux", nemit, 1)
    hostapi.set_devices([0])
    sol, st = prob.solve_seeded(SEED + 77)
    assert st["esc"] == 0 and st["emitted"] == nemit
    k = sol[0].mean() / 1e6                       # gradT = dT / L = 1e6 K/m
    assert abs(k / mat.cond() - 1.0) < 0.01, (mname, k, mat.cond())


def test_bulk_128_cubed_matches_the_oracle_in_philox_mode(gpu_ctx, omats):
    """C5's grid (128^3 cells, the N-D walk with long flights cut into pieces over the warp, fp64 RED into the 67-MB field)
    with the Si-like material, same Philox streams on both sides: counters exact, field within 1e-9 of the row scale."""
    mat, dom = omats["silicon"], cases.bulk(div=(128, 128, 128))
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 20000, 50)
    got, gst = gpu_ctx.solve(prob.desc, seed=SEED + 5)
    ref, rst = prob.solve(rng=orc.RNG_PHILOX, seed=SEED + 5)
    assert (gst["emitted"], gst["steps"], gst["esc"]) == (rst["emitted"], rst["steps"], rst["esc"])
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-9 * scale).all(), np.abs(got - ref).max(axis=1) / scale[:, 0]
