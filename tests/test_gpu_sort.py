"""K3 with a key (north star: "periodic phonon sort and compaction by cell"): the counting sort behind
mcb_options.sort_mode.  The reference never reorders phonons (its loop `break`s, problem.cpp:411,425,434), so the parity
statements are (1) integer work, bit-exact: the sorted slots are a permutation of exactly the survivors the oracle's trace
predicts, every survivor's key is the field column of the oracle's (sdom, cell) (field.cpp:25-45, subdomain.cpp:148-159),
survivors are contiguous from slot 0 and their bins are non-decreasing; (2) a solve does not depend on the slot order:
counters identical, field to 1e-9 (the RNG stream is keyed by the particle id)."""
import numpy as np
import pytest

from oracle import pyoracle as orc
from tests import cases

pytestmark = pytest.mark.gpu
SEED = 20240607


def columns_of(dom, sdom, cell):
    """Field::init (field.cpp:25-45): subdomain k owns columns [offset_k, offset_k + prod(shape_k)); cell (i, j, k) is column
    offset + i + j * shape0 + k * shape0 * shape1."""
    d = dom.desc
    off, shp = [], []
    cols = 0
    for k in range(d.nsdom):
        s = [int(x) for x in d.sdoms[k].shape]
        n = s[0] * s[1] * s[2]
        off.append(cols if n > 0 else -1)
        shp.append(s)
        cols += n
    off, shp = np.array(off), np.array(shp)
    c = off[sdom] + cell[:, 0] + cell[:, 1] * shp[sdom, 0] + cell[:, 2] * shp[sdom, 0] * shp[sdom, 1]
    return np.where(off[sdom] < 0, 0, c)


PROBE_CASES = [("grey", "slab", 0), ("grey", "slab", 7), ("silicon", "film", 5), ("grey", "wire", 6), ("silicon", "tube", 9),
               ("silicon", "jct", 4), ("grey", "bulk64", 3), ("silicon_small", "skew", 5)]


@pytest.mark.parametrize("mname,dname,nsteps", PROBE_CASES)
@pytest.mark.parametrize("sorted_", [True, False])
def test_sort_probe_permutes_exactly_the_survivors_and_orders_them_by_cell(gpu_ctx, omats, mname, dname, nsteps, sorted_):
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 6000, 12)
    ref = prob.trace(SEED, 0, prob.nemit, nsteps)                       # the CPU oracle's per-particle state after nsteps trips
    maxscat, maxloop = prob.desc.maxscat, prob.desc.maxloop
    surv = (ref["alive"] != 0) & (ref["steps"] == nsteps) & (ref["nscat"] < maxscat) & (nsteps < maxloop)
    want_col = columns_of(dom, ref["sdom"], ref["cell"].reshape(-1, 3))
    b, c, p, cpb = gpu_ctx.sort_probe(prob.desc, SEED, 0, prob.nemit, nsteps, sorted_)
    live = int(surv.sum())
    assert 0 < live <= prob.nemit
    assert (p[:live] >= 0).all() and (p[live:] == -1).all()             # compaction: survivors contiguous from slot 0
    assert np.array_equal(np.sort(p[:live]), np.nonzero(surv)[0])       # ... and exactly the oracle's survivors, once each
    assert np.array_equal(c[:live], want_col[p[:live]])                 # key = the oracle's field column, bit-exact
    assert cpb >= 1 and np.array_equal(b[:live], c[:live] // cpb)
    if sorted_:
        assert (np.diff(b[:live]) >= 0).all()                           # sortedness


def test_sort_probe_bins_follow_the_column_ranges_of_a_fine_grid(gpu_ctx, omats):
    """More columns than bins (128^3 cells, 16384 bins): bin = column // cols_per_bin, still non-decreasing."""
    mat = omats["silicon"]
    dom = cases.bulk(div=(128, 128, 128))
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 50000, 10)
    b, c, p, cpb = gpu_ctx.sort_probe(prob.desc, SEED, 0, prob.nemit, 2, True)
    live = int((p >= 0).sum())
    assert cpb == 128 and live > 40000
    assert (p[:live] >= 0).all() and len(np.unique(p[:live])) == live
    assert np.array_equal(b[:live], c[:live] // cpb) and (np.diff(b[:live]) >= 0).all()
    assert 0 <= c[:live].min() and c[:live].max() < 128 ** 3


SOLVE_CASES = [("grey", "wire", "multi"), ("silicon", "tube", "multi"), ("silicon", "bulk64", "flux"), ("grey", "slab", "multi")]


@pytest.mark.parametrize("mname,dname,pkind", SOLVE_CASES)
def test_sorted_solve_equals_the_unsorted_one(gpu_ctx, omats, mname, dname, pkind):
    """sort_mode is scheduling only: same counters, field to 1e-9; the stats say the sort ran."""
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, pkind, 600000, 40)                     # enough phonons for several compactions of the decay phase
    zero = dict(slots=0, steps_per_launch=0, block=0, ctas_per_sm=0, tally_mode=0, emit_mode=0, decay_mode=0, compact_pct=0, sort_mode=0)
    # two loop trips per launch throughout and a 99 % threshold: the decay phase compacts after nearly every launch
    many = {**zero, "steps_per_launch": 2, "decay_mode": 1, "compact_pct": 99}
    gpu_ctx.set_options(**many)
    base, bst = gpu_ctx.solve(prob.desc, seed=SEED)
    assert bst["compactions"] >= 1 and bst["sorts"] == 0
    scale = np.abs(base).max(axis=1, keepdims=True)
    try:
        for mode in (1, 3):
            gpu_ctx.set_options(**{**many, "sort_mode": mode})
            got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
            assert (gst["steps"], gst["esc"], gst["emitted"]) == (bst["steps"], bst["esc"], bst["emitted"])
            assert gst["sorts"] >= 1 and gst["compactions"] >= gst["sorts"]
            if mode == 1:
                assert gst["sorts"] == gst["compactions"]
            assert (np.abs(got - base) <= 1e-9 * scale).all()
    finally:
        gpu_ctx.set_options(**zero)


@pytest.mark.parametrize("mname,dname,pkind,maxscat", [("grey", "slab", "multi", 300), ("silicon", "slab", "cumflux", 60), ("silicon", "tube", "multi", 40)])
def test_fused_compaction_equals_separate_passes(gpu_ctx, omats, mname, dname, pkind, maxscat):
    """The decay phase with the compaction inside k_step (default) vs separate k_compact passes (decay_mode 2) vs a fixed S with
    passes (decay_mode 1), on populations well above one tile per CTA, resident and streaming: counters identical, field to 1e-9."""
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, pkind, 700000, maxscat, size=5 if pkind.startswith("cum") else 0)
    zero = dict(slots=0, steps_per_launch=0, block=0, ctas_per_sm=0, tally_mode=0, emit_mode=0, decay_mode=0, compact_pct=0, sort_mode=0, decay_pct=0)
    try:
        gpu_ctx.set_options(**{**zero, "decay_mode": 2})
        base, bst = gpu_ctx.solve(prob.desc, seed=SEED)
        scale = np.abs(base).max(axis=1, keepdims=True)
        for opts in (dict(), dict(decay_pct=4), dict(decay_pct=40), dict(steps_per_launch=1, slots=148 * 768 * 3), dict(decay_mode=1, steps_per_launch=5)):
            gpu_ctx.set_options(**{**zero, **opts})
            got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
            assert (gst["steps"], gst["esc"], gst["emitted"]) == (bst["steps"], bst["esc"], bst["emitted"]), opts
            assert (np.abs(got - base) <= 1e-9 * scale).all(), opts
            if opts.get("decay_mode", 0) == 0:
                assert gst["compactions"] == 0                      # no separate K3 pass ran
    finally:
        gpu_ctx.set_options(**zero)
