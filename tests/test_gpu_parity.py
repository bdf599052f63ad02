"""GPU parity tests: the CUDA path, called through the C ABI (include/mcb.h), against the CPU oracle
on the same seeded inputs.  Bars: integer work bit-exact; fp64 state within 1e-11
relative in Philox mode (libm vs CUDA log/sincos and FMA contraction differ by ulps); tally fields
within 1e-9 of the field scale (atomic summation order)."""
import numpy as np
import pytest

from montecarlocpp_b200 import abi, capi
from oracle import pyoracle as orc
from tests import cases

pytestmark = pytest.mark.gpu

SEED = 0x5EED0000


def test_philox_device_words_match_random123_kat():
    # Random123 kat_vectors: philox4x32-10
    assert capi.philox_words(0, 0, 0, 0) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert capi.philox_words(0xffffffffffffffff, 0xffffffffffffffff, 0xffffffff, 0xffffffff) == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert capi.philox_words(0x299f31d0a4093822, 0x85a308d3243f6a88, 0x13198a2e, 0x03707344) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    assert capi.philox_words(SEED, 12345, 7, 1) == orc.philox_words(SEED, 12345, 7, 1)


@pytest.mark.parametrize("mname", ["grey", "silicon"])
def test_alias_tables_bit_exact(gpu_ctx, omats, mname):
    mat = omats[mname]
    gpu_ctx.upload_material(mat.desc)
    for which in (0, 1):
        a, b = gpu_ctx.alias(which), mat.alias(which)
        for x, y in zip(a, b):
            assert np.array_equal(x, y)


@pytest.mark.parametrize("dname", ["film", "skew", "tube", "wire"])
def test_cell_index_bit_exact(gpu_ctx, omats, dname):
    dom = cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, omats["grey"], dom)
    rng = np.random.default_rng(1)
    n = 200000
    nsd = dom.desc.nsdom
    sd = rng.integers(0, nsd, n).astype(np.int32)
    pos = np.zeros((n, 3))
    for s in range(nsd):
        S = dom.desc.sdoms[s]
        o = np.array(S.origin[:]); m = np.array(S.mat[:]).reshape(3, 3).T
        k = sd == s
        u = rng.uniform(-0.05, 1.05, (k.sum(), 3))
        # a third of the points exactly on cell faces (the floor() boundary)
        shape = np.array(S.shape[:], float)
        snap = rng.random(k.sum()) < 0.33
        u[snap] = np.round(u[snap] * shape) / shape
        pos[k] = o + u @ m.T
    assert np.array_equal(gpu_ctx.cell_index(pos, sd), dom.cell_index(pos, sd))


@pytest.mark.parametrize("dname,rows", [("film", 4), ("slab", 1), ("wire", 4), ("skew", 3), ("tube", 4), ("bulk", 3), ("bulk64", 4)])
def test_accumulate_matches_oracle(gpu_ctx, omats, dname, rows):
    """Field::accumulate (field.cpp:92-220): 1-D shares and the N-D crossing walk."""
    dom = cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, omats["grey"], dom)
    rng = np.random.default_rng(2)
    n = 20000
    nsd = dom.desc.nsdom
    sd = rng.integers(0, nsd, n).astype(np.int32)
    b = np.zeros((n, 3)); e = np.zeros((n, 3))
    for s in range(nsd):
        S = dom.desc.sdoms[s]
        o = np.array(S.origin[:]); m = np.array(S.mat[:]).reshape(3, 3).T
        k = sd == s
        ub = rng.uniform(0, 1, (k.sum(), 3)); ue = rng.uniform(0, 1, (k.sum(), 3))
        short = rng.random(k.sum()) < 0.5
        ue[short] = np.clip(ub[short] + rng.normal(0, 0.03, (short.sum(), 3)), 0, 1)
        axis = rng.random(k.sum()) < 0.1          # axis-aligned segments: dcoord == 0 on two axes
        ue[axis, 1:] = ub[axis, 1:]
        b[k] = o + ub @ m.T; e[k] = o + ue @ m.T
    amt = rng.normal(0, 1, (n, rows))
    got = gpu_ctx.accumulate(rows, sd, b, e, amt)
    ref = dom.accumulate(rows, sd, b, e, amt)
    scale = np.abs(ref).max()
    assert np.abs(got - ref).max() <= 1e-11 * scale
    # the tally conserves the deposited amount (field.cpp:134-154 shares sum to `amount`)
    assert np.allclose(got.sum(axis=1), amt.sum(axis=0), rtol=1e-9, atol=1e-9 * np.abs(amt).sum())


TRACE_CASES = [("grey", "slab", "multi", 50), ("silicon", "film", "multi", 40), ("grey", "bulk", "flux", 30),
               ("silicon", "jct", "temp", 60), ("grey", "tee", "multi", 60), ("silicon", "tube", "multi", 60),
               ("grey", "wire", "flux", 40), ("silicon_small", "skew", "multi", 40)]


@pytest.mark.parametrize("mname,dname,pkind,nsteps", TRACE_CASES)
def test_trace_state_parity(gpu_ctx, omats, mname, dname, pkind, nsteps):
    """KA6: per-particle (sdom, cell, nscat, alive, steps, w, p, sign) identical, fp state to ulps."""
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, pkind, 4000, 20)
    for k in (0, 1, nsteps):
        ref = prob.trace(SEED, 0, prob.nemit, k)
        got = gpu_ctx.trace(prob.desc, SEED, 0, prob.nemit, k)
        for key in ("w", "p", "sign", "alive", "sdom", "nscat", "steps", "cell"):
            assert np.array_equal(got[key], ref[key]), (key, k)
        scale = np.abs(ref["pos"]).max()
        assert np.abs(got["pos"] - ref["pos"]).max() <= 1e-11 * scale, k
        assert np.abs(got["dir"] - ref["dir"]).max() <= 1e-11, k
        assert np.allclose(got["scat_next"], ref["scat_next"], rtol=1e-10, atol=1e-11 * scale), k


SOLVE_CASES = [("grey", "slab", "multi", 0), ("grey", "slab", "temp", 0), ("silicon", "film", "multi", 0),
               ("grey", "bulk", "flux", 0), ("silicon", "jct", "multi", 0), ("grey", "tee", "multi", 0),
               ("silicon", "tube", "multi", 0), ("grey", "wire", "multi", 0), ("silicon_small", "skew", "flux", 0),
               ("grey", "film", "cumtemp", 4), ("silicon", "slab", "cumflux", 3), ("grey", "bulk64", "multi", 0),
               ("silicon", "bulk64", "flux", 0)]


@pytest.mark.parametrize("mname,dname,pkind,size", SOLVE_CASES)
def test_solve_matches_oracle_philox(gpu_ctx, omats, mname, dname, pkind, size):
    """FieldProblem::solve end to end with the shared Philox streams: counters exact, field to 1e-9."""
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    gpu_ctx.set_options(slots=0, steps_per_launch=0, tally_mode=0)
    prob = orc.Problem(mat, dom, pkind, 20000, 30, size=size)
    ref, rst = prob.solve(rng=orc.RNG_PHILOX, seed=SEED)
    got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
    assert gst["emitted"] == rst["emitted"] == prob.nemit
    assert gst["steps"] == rst["steps"]
    assert gst["esc"] == rst["esc"]
    if dname in ("slab", "wire", "skew", "bulk", "film", "bulk64"):
        assert gst["esc"] == 0                     # KA5: no escapes on single-box domains
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-9 * scale).all()


ZERO_OPTS = dict(slots=0, steps_per_launch=0, block=0, ctas_per_sm=0, tally_mode=0, emit_mode=0, decay_mode=0, compact_pct=0, sort_mode=0, decay_pct=0)


@pytest.mark.parametrize("opts", [dict(steps_per_launch=1, slots=4096), dict(steps_per_launch=7, slots=1000),
                                  dict(steps_per_launch=64, slots=0, tally_mode=2), dict(block=128, ctas_per_sm=1, slots=30000), dict(block=512, steps_per_launch=5),
                                  dict(tally_mode=3, steps_per_launch=3), dict(tally_mode=1, block=256, steps_per_launch=2),
                                  dict(slots=1001, steps_per_launch=2), dict(decay_mode=2), dict(decay_mode=1, steps_per_launch=3, compact_pct=99),
                                  dict(decay_pct=3), dict(decay_pct=60, steps_per_launch=1, slots=8192), dict(sort_mode=1, decay_mode=2, compact_pct=99)])
def test_schedule_options_do_not_change_results(gpu_ctx, omats, opts):
    """Slots / S / tally mode are scheduling only: Philox keyed by particle id makes the result
    independent of them (up to fp summation order)."""
    mat, dom = omats["silicon"], cases.film()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 30000, 25)
    gpu_ctx.set_options(**ZERO_OPTS)
    base, bst = gpu_ctx.solve(prob.desc, seed=SEED)
    gpu_ctx.set_options(**{**ZERO_OPTS, **opts})
    got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
    gpu_ctx.set_options(**ZERO_OPTS)
    assert (gst["steps"], gst["esc"], gst["emitted"]) == (bst["steps"], bst["esc"], bst["emitted"])
    scale = np.abs(base).max(axis=1, keepdims=True)
    assert (np.abs(got - base) <= 1e-9 * scale).all()


def test_particle_ranges_add_up(gpu_ctx, omats):
    """solve([0,n)) == solve([0,k)) + solve([k,n)): the per-thread partial sums of main.cpp:162-165
    and the per-GPU shards of the multi-GPU path."""
    mat, dom = omats["grey"], cases.slab()
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 40000, 50)
    full, fst = gpu_ctx.solve(prob.desc, seed=SEED)
    k = 12345
    a, ast = gpu_ctx.solve(prob.desc, seed=SEED, n_begin=0, n_end=k)
    b, bst = gpu_ctx.solve(prob.desc, seed=SEED, n_begin=k, n_end=prob.nemit)
    assert ast["steps"] + bst["steps"] == fst["steps"]
    assert ast["emitted"] == k and bst["emitted"] == prob.nemit - k
    scale = np.abs(full).max(axis=1, keepdims=True)
    assert (np.abs(a + b - full) <= 1e-9 * scale).all()
    empty, est = gpu_ctx.solve(prob.desc, seed=SEED, n_begin=7, n_end=7)
    assert est["steps"] == 0 and not empty.any()


@pytest.mark.parametrize("S", [1, 16])
def test_single_cell_ballistic_tally_does_not_overflow_the_limbs(gpu_ctx, tmp_path, S):
    """Worst case for the carry-free fixed-point histograms (mcb_device.cuh: deposit): a one-cell domain (accumFlag -1: every
    lane of every warp deposits into the SAME entry on every loop trip) with a ballistic grey material (flights span the
    domain, payloads near fx_max).  The limb widths are chosen so that neither limb can overflow between two flushes; the
    field must still match the oracle to 1e-9 and the counters exactly."""
    from montecarlocpp_b200 import materials
    disp, relax = materials.write_grey(str(tmp_path), inv_tau=1e7)            # mean free path 600 um >> domain
    mat = orc.Material(disp, relax, 300.0)
    dom = orc.Domain.create("bulk", [1e-6, 1e-6, 1e-6], [0, 0, 0], 1.0)
    cases.upload(gpu_ctx, mat, dom)
    gpu_ctx.set_options(slots=148 * 896 * 2, steps_per_launch=S, tally_mode=1)     # warp histograms, every lane busy
    prob = orc.Problem(mat, dom, "multi", 300000, 40, maxloop=400)
    ref, rst = prob.solve(rng=orc.RNG_PHILOX, seed=SEED)
    got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
    gpu_ctx.set_options(slots=0, steps_per_launch=0, tally_mode=0)
    assert gst["steps"] == rst["steps"] and gst["esc"] == rst["esc"] == 0
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-9 * scale).all(), np.abs(got - ref).max(axis=1) / scale[:, 0]


def test_statistical_parity_with_mt19937_oracle(gpu_ctx, omats):
    """North-star bar: T and q profiles within 3 sigma of batch-means error against the oracle run
    with the reference's own RNG family (mt19937), k_eff within 1 %."""
    mat, dom = omats["grey"], cases.slab(ncell=10)
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 100000, 200)
    B = 8
    g = np.stack([gpu_ctx.solve(prob.desc, seed=SEED + b)[0] for b in range(B)])
    r = np.stack([prob.solve(rng=orc.RNG_MT19937, seed=1000 + 17 * b)[0] for b in range(B)])
    mg, mr = g.mean(0), r.mean(0)
    se = np.sqrt(g.var(0, ddof=1) / B + r.var(0, ddof=1) / B)
    z = np.abs(mg - mr)[:2] / se[:2]              # T and q_x rows
    assert (z < 4.0).all() and (z < 3.0).mean() > 0.9
    qg, qr = mg[1].mean(), mr[1].mean()
    assert abs(qg / qr - 1.0) < 0.01


def test_statistical_parity_with_the_reference_itself(gpu_ctx, matfiles, omats):
    """The same bar against the REFERENCE'S OWN OBJECTS (oracle/_ref/ref_driver: FieldProblem::solve compiled from the
    reference's sources, mt19937; the slab composed from its templates): T / q_x profiles within 3 sigma of batch-means error,
    mean flux within 1 %."""
    from oracle import refbin
    if not refbin.driver_available():
        pytest.skip("oracle/_ref/ref_driver not built (make -C oracle ref in the dev container)")
    disp, relax = matfiles["grey"]
    mat, dom = omats["grey"], cases.slab(ncell=10)
    cases.upload(gpu_ctx, mat, dom)
    prob = orc.Problem(mat, dom, "multi", 100000, 200)
    B = 8
    g = np.stack([gpu_ctx.solve(prob.desc, seed=SEED + 100 + b)[0] for b in range(B)])
    r = np.stack([refbin.drive(disp, relax, 300.0, "slab", [100e-9] * 3, [10, 0, 0], 1.0, "multi", 100000, 200,
                               seed=4000 + 16 * b, threads=4)["output"] for b in range(B)])
    mg, mr = g.mean(0), r.mean(0)
    se = np.sqrt(g.var(0, ddof=1) / B + r.var(0, ddof=1) / B)
    z = np.abs(mg - mr)[:2] / se[:2]
    assert (z < 4.0).all() and (z < 3.0).mean() > 0.9
    assert abs(mg[1].mean() / mr[1].mean() - 1.0) < 0.01


@pytest.mark.parametrize("div", [[0, 0, 0, 0], [2, 2, 2, 1]], ids=["cells", "grid"])
def test_octet_domain_through_the_reference_binding(gpu_ctx, matfiles, div):
    """The reference's 42-subdomain OctetDomain (domain.cpp:576-1280; prisms, pyramids, triangular prisms, boxes, 70 Inter
    pairs, mirror-periodic pairs) is not restated anywhere in this repo.  Its own objects, flattened by the reference-side
    binding (oracle/ref_driver.cpp `flatten`), run on the device through mcb_upload_domain / mcb_solve; the result must agree
    with the reference's CPU solve of the same objects: zero escapes (every hand-off finds its partner), per-cell 4-sigma,
    domain-mean flux within 3 sigma."""
    from oracle import refbin
    if not refbin.driver_available():
        pytest.skip("oracle/_ref/ref_driver not built (make -C oracle ref in the dev container)")
    from montecarlocpp_b200 import hostapi
    disp, relax = matfiles["silicon_small"]
    dim, dT, nemit, maxscat, B = [1e-6, 1e-7, 1e-7, 1e-8], 1.0, 60000, 50, 6
    fl = refbin.flatten(disp, relax, 300.0, "octet", dim, div, dT, "multi", nemit, maxscat)
    mat = hostapi.Material(disp, relax, 300.0)
    gpu_ctx.upload_material(mat.desc)
    gpu_ctx.upload_domain(fl.domain)
    g, esc = [], 0
    for b in range(B):
        sol, st = gpu_ctx.solve(fl.problem, seed=SEED + 300 + b)
        g.append(sol); esc += st["esc"]
    r = [refbin.drive(disp, relax, 300.0, "octet", dim, div, dT, "multi", nemit, maxscat, seed=9000 + 16 * b, threads=4) for b in range(B)]
    assert esc == 0 and sum(x["esc"] for x in r) == 0
    g, r = np.stack(g), np.stack([x["output"] for x in r])
    assert g.shape == r.shape and np.isfinite(g).all()
    mg, mr = g.mean(0), r.mean(0)
    se = np.sqrt(g.var(0, ddof=1) / B + r.var(0, ddof=1) / B)
    z = np.abs(mg - mr) / np.where(se > 0, se, 1.0)
    assert (z < 4.5).all() and (z < 3.0).mean() > 0.9, z.max()
    # Domain::average (OctetDomain::WeightF weights exported by the binding): the averaged rows of the batches agree within 3.5 sigma
    ag = np.stack([fl.average(x)[:, 0] for x in g]); ar = np.stack([fl.average(x)[:, 0] for x in r])
    sa = np.sqrt(ag.var(0, ddof=1) / B + ar.var(0, ddof=1) / B)
    assert (np.abs(ag.mean(0) - ar.mean(0)) <= 3.5 * sa).all(), (ag.mean(0), ar.mean(0), sa)


@pytest.mark.parametrize("div", [[0, 0, 0, 0], [2, 2, 2, 1]], ids=["cells", "grid"])
def test_octet_domain_of_the_host_mirror(matfiles, div):
    """The PRODUCT's OctetDomain (host mirror, geometry table of octet_table.inc) through FieldProblem::solve, against the
    reference's CPU solve of its own OctetDomain: zero escapes, per-cell 4.5 sigma, Domain::average within 3.5 sigma."""
    from oracle import refbin
    if not refbin.driver_available():
        pytest.skip("oracle/_ref/ref_driver not built (make -C oracle ref in the dev container)")
    from montecarlocpp_b200 import hostapi
    disp, relax = matfiles["silicon_small"]
    dim, dT, nemit, maxscat, B = [1e-6, 1e-7, 1e-7, 1e-8], 1.0, 60000, 50, 6
    mat = hostapi.Material(disp, relax, 300.0)
    dom = hostapi.Domain("octet", dim, div, dT)
    prob = hostapi.FieldProblem(mat, dom, "multi", nemit, maxscat)
    hostapi.set_devices([0])
    g, esc = [], 0
    for b in range(B):
        sol, st = prob.solve_seeded(SEED + 700 + b)
        g.append(sol); esc += st["esc"]
    r = [refbin.drive(disp, relax, 300.0, "octet", dim, div, dT, "multi", nemit, maxscat, seed=7000 + 16 * b, threads=4) for b in range(B)]
    assert esc == 0 and sum(x["esc"] for x in r) == 0
    g, r = np.stack(g), np.stack([x["output"] for x in r])
    assert g.shape == r.shape and np.isfinite(g).all()
    se = np.sqrt(g.var(0, ddof=1) / B + r.var(0, ddof=1) / B)
    z = np.abs(g.mean(0) - r.mean(0)) / np.where(se > 0, se, 1.0)
    assert (z < 4.5).all() and (z < 3.0).mean() > 0.9, z.max()
    ag = np.stack([dom.average(x)[:, 0] for x in g]); ar = np.stack([dom.average(x)[:, 0] for x in r])
    sa = np.sqrt(ag.var(0, ddof=1) / B + ar.var(0, ddof=1) / B)
    assert (np.abs(ag.mean(0) - ar.mean(0)) <= 3.5 * sa).all(), (ag.mean(0), ar.mean(0), sa)


def test_bulk_conductivity_within_one_percent(gpu_ctx, omats):
    """KA1: <q_x>/|grad T| -> Material::cond() (material.cpp:160-161)."""
    # grey at 1.6e7 phonons; the synthetic silicon's heavy-tailed free paths need ~1e9 for the same 1 % bar:
    # tests/test_gpu_configs.py::test_bulk_conductivity_within_one_percent_both_materials
    for mname, tol in (("grey", 0.01),):
        mat, dom = omats[mname], cases.bulk()
        cases.upload(gpu_ctx, mat, dom)
        # maxscat = 1: only the first flight carries signal (later flights are isotropic, zero-mean noise)
        prob = orc.Problem(mat, dom, "flux", 16000000, 1)
        sol, st = gpu_ctx.solve(prob.desc, seed=SEED)
        k = sol[0].mean() / 1e6
        assert st["esc"] == 0
        assert abs(k / mat.cond() - 1.0) < tol, (mname, k, mat.cond())


def test_error_behaviour(omats):
    ctx = capi.Context(0)
    mat, dom = omats["grey"], cases.slab()
    prob = orc.Problem(mat, dom, "multi", 1000, 10)
    ctx.cols = 1
    with pytest.raises(capi.McbError) as e:       # solve before upload
        ctx.solve(prob.desc)
    assert e.value.code == abi.MCB_ESTATE
    cases.upload(ctx, mat, dom)
    bad = abi.ProblemDesc.from_buffer_copy(prob.desc)
    bad.rows = 3
    with pytest.raises(capi.McbError) as e:
        ctx.solve(bad)
    assert e.value.code == abi.MCB_EINVAL
    with pytest.raises(capi.McbError):
        ctx.solve(prob.desc, n_begin=5, n_end=2)
    ctx.close()


def test_long_histories_and_cum_bins_are_validated(gpu_ctx, omats):
    """(a) the reference takes any `long` maxloop (default 100 * maxscat, problem.cpp:339): the pid|step word gives
    the loop trip up to 32 bits (the Philox event counter), so maxloop >= 2^28 solves -- and gives the very field of a short
    bound no history reaches; (b) a Cum* step too small for maxscat and size would index past the field rows: rejected."""
    mat, dom = omats["grey"], cases.slab()
    cases.upload(gpu_ctx, mat, dom)
    base = orc.Problem(mat, dom, "multi", 20000, 50)
    ref, rst = gpu_ctx.solve(base.desc, seed=SEED)
    for maxloop in (1 << 28, (1 << 31) + 5, 0xFFFFFFFE):
        d = abi.ProblemDesc.from_buffer_copy(base.desc)
        d.maxloop = maxloop
        got, gst = gpu_ctx.solve(d, seed=SEED)
        assert (gst["emitted"], gst["steps"], gst["esc"]) == (rst["emitted"], rst["steps"], rst["esc"])
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(got - ref) <= 1e-9 * scale).all()
    d = abi.ProblemDesc.from_buffer_copy(base.desc)
    d.maxloop = 0xFFFFFFFF
    with pytest.raises(capi.McbError) as e:
        gpu_ctx.solve(d, seed=SEED)
    assert e.value.code == abi.MCB_EINVAL and "maxloop" in str(e.value)
    cum = orc.Problem(mat, dom, "cumtemp", 2000, 50, size=5)
    bad = abi.ProblemDesc.from_buffer_copy(cum.desc)
    bad.step = max(1, cum.desc.step // 2)
    with pytest.raises(capi.McbError) as e:
        gpu_ctx.solve(bad, seed=SEED)
    assert e.value.code == abi.MCB_EINVAL and "step" in str(e.value)


def test_host_mirror_solve_matches_oracle(matfiles, omats):
    """The C++ mirror of the reference API (Material -> FilmDomain/TubeDomain -> MultiProblem::solve) drives the
    same device path: identical counters and field as the oracle on the same Philox seed."""
    from montecarlocpp_b200 import hostapi
    for dname, dim, div, dT, odom in (("film", [1e-6, 1e-7, 1e-6], [0, 20, 0], 1.0, cases.film()),
                                      ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0, cases.tube())):
        hm, hd = hostapi.Material(*matfiles["silicon"]), hostapi.Domain(dname, dim, div, dT)
        hp = hostapi.FieldProblem(hm, hd, "multi", 20000, 30)
        got, gst = hp.solve_seeded(SEED)
        op = orc.Problem(omats["silicon"], odom, "multi", 20000, 30)
        ref, rst = op.solve(rng=orc.RNG_PHILOX, seed=SEED)
        assert (gst["emitted"], gst["steps"], gst["esc"]) == (rst["emitted"], rst["steps"], rst["esc"])
        scale = np.abs(ref).max(axis=1, keepdims=True)
        assert (np.abs(got - ref) <= 1e-9 * scale).all()
        # the reference signature solve(Rng& gen, Progress*): the seed is (gen() << 32) | gen() of mt19937(7)
        raw = np.random.RandomState(7).randint(0, 2**32, size=2, dtype=np.uint64)
        a, _ = hp.solve(mt_seed=7)
        b, _ = hp.solve_seeded(int((int(raw[0]) << 32) | int(raw[1])))
        assert np.allclose(a, b, rtol=1e-9, atol=1e-12 * np.abs(b).max())


def test_cli_driver_keeps_the_reference_stdout_blocks(matfiles, tmp_path):
    """main.cpp grammar + printSolution format (main.cpp:67-84, 146-210)."""
    import os, shutil, subprocess
    from montecarlocpp_b200 import capi
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "montecarlo")
    for f in matfiles["grey"]:
        shutil.copy(f, tmp_path)
    r = subprocess.run([exe, str(tmp_path), "grey", "300", "film", "1e-6", "1e-7", "10", "multi", "20000", "20", "0", "3"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for token in ("Material ", "FilmDomain ", "MultiProblem ", "  nemit:   20000", "  maxloop: 2000", "Solution 0", "Solution 2",
                  "  seeds: ", "Output", "Mean", "Standard Deviation", "Total time: ", "esc: 0"):
        assert token in out, token
    block = out.split("Mean\n")[1].split("\n\n")[0].strip().split("\n")
    assert len(block) == 4 and all(len(row.split()) == 10 for row in block)
    vals = np.array([[float(x) for x in row.split()] for row in block])
    assert (vals[1] > 0).all()                      # heat flows down the gradient in every cell
    bad = subprocess.run([exe, str(tmp_path), "grey", "300", "icosahedron", "1"], capture_output=True, text=True)
    assert bad.returncode != 0 and "Invalid domain" in bad.stderr
    # the octet truss (main.cpp:376-387): gridded struts -> Output (4 x 32 columns) and the Averaged block (4 x 1) of Domain::average
    r = subprocess.run([exe, str(tmp_path), "grey", "300", "octet", "1e-6", "1e-7", "1e-7", "1e-8", "2", "2", "2", "1", "1.0",
                        "multi", "40000", "20", "0", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    for token in ("OctetDomain ", "  div: [2 2 2 1]", "Output", "Averaged", "esc: 0"):
        assert token in r.stdout, token
    avg = r.stdout.split("Averaged\n")[1].split("\n\n")[0].strip().split("\n")
    assert len(avg) == 4 and all(len(row.split()) == 1 for row in avg)


TRAJ_CASES = [("grey", "tube", dict(pos=[5e-7, 6e-8, 1e-8], dir=[1, 0.3, 0.2]), 40), ("silicon", "jct", dict(pos=[1e-7, 5e-8, 2.5e-8]), 30),
              ("grey", "tee", dict(), 50), ("silicon", "film", dict(prop=(10, 1), pos=[5e-7, 5e-8, 5e-7], dir=[0, 1, 0]), 20),
              ("grey", "slab", dict(), 200), ("silicon_small", "skew", dict(), 25)]


@pytest.mark.parametrize("mname,dname,kw,maxscat", TRAJ_CASES)
def test_trajectory_matches_oracle(gpu_ctx, omats, mname, dname, kw, maxscat):
    """TrajProblem::solve (problem.cpp:226-299): the boundary trace is identical (integers), the recorded
    polyline agrees to ulps, for every constructor form (prop / pos / dir given or drawn)."""
    mat, dom = omats[mname], cases.DOMAINS[dname]()
    cases.upload(gpu_ctx, mat, dom)
    for seed in (1, 2, 3):
        ref = orc.traj(mat, dom, seed=SEED + seed, maxscat=maxscat, **kw)
        t = orc.make_traj_desc(dom, maxscat, 0, kw.get("prop"), kw.get("pos"), kw.get("dir"),
                               lambda q: orc.lib().orc_domain_locate(dom.h, q.ctypes.data_as(abi.c_double_p)))
        got = gpu_ctx.traj(t, SEED + seed)
        for k in ("step_sdom", "step_in", "step_in_kind", "step_out", "step_out_kind"):
            assert np.array_equal(got[k], ref[k]), (k, seed)
        assert got["escaped"] == ref["escaped"] and got["points"].shape == ref["points"].shape
        scale = np.abs(ref["points"]).max()
        assert np.abs(got["points"] - ref["points"]).max() <= 1e-11 * scale


def test_cli_check_mode_traces_every_boundary(matfiles, tmp_path):
    """`check` (main.cpp:104-143): 6 axis rays from every checkpoint; the printed boundary types show the wiring."""
    import os, shutil, subprocess
    from montecarlocpp_b200 import capi
    exe = os.path.join(os.path.dirname(capi.LIB_PATH), "montecarlo")
    from montecarlocpp_b200 import materials
    materials.write_grey(str(tmp_path), inv_tau=6.0)      # mean free path 1 km: every ray reaches a wall (seeds are random)
    r = subprocess.run([exe, str(tmp_path), "grey", "300", "tube", "1e-6", "5e-8", "2e-8", "8", "4", "check", "1", "0", "0", "0", "1", "0", "0", "0", "1"],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    assert out.count("TrajProblem ") == 18 and out.count("Trajectory ") >= 18 and "Combined Trajectory" in out
    assert "Escaped" not in out
    # sdom 0's top face (5) is the Inter hand-off to sdom 1; the partner face is 2 there, and so on round the L
    for token in (" 0: -1  Null ->  5 Inter", " 0: -1  Null ->  3 PeriP", " 0: -1  Null ->  2  Spec", " 1: -1  Null ->  1 Inter",
                  " 1: -1  Null ->  2 Inter", " 2: -1  Null ->  4 Inter", " 2: -1  Null ->  1  Spec", " 1:  2 Inter -> ", " Diff"):
        assert token in out, token
    r2 = subprocess.run([exe, str(tmp_path), "grey", "300", "film", "1e-6", "1e-7", "10", "traj", "5e-7", "5e-8", "5e-7", "0", "1", "0", "3", "0"],
                        capture_output=True, text=True, timeout=120)
    assert r2.returncode == 0 and " 0: -1  Null ->  4  Diff" in r2.stdout


@pytest.mark.parametrize("dname", sorted(cases.NONBOX))
def test_nonbox_cells_match_oracle(gpu_ctx, omats, dname):
    """N3 cells (subdomain.h:200-630): tri-prism / tetrahedron / prism / pyramid emission folds, Polygon and Triangle
    emitters, generic plane loop, literal cellVol: per-particle state and whole solves against the oracle."""
    mat, dom = omats["grey"], cases.NONBOX[dname]()
    cases.upload(gpu_ctx, mat, dom)
    gpu_ctx.set_options(**ZERO_OPTS)
    prob = orc.Problem(mat, dom, "multi", 20000, 25)
    for k in (0, 30):
        ref = prob.trace(SEED, 0, 4000, k)
        got = gpu_ctx.trace(prob.desc, SEED, 0, 4000, k)
        for key in ("w", "p", "sign", "alive", "sdom", "nscat", "steps", "cell"):
            assert np.array_equal(got[key], ref[key]), (key, k)
        scale = np.abs(ref["pos"]).max()
        assert np.abs(got["pos"] - ref["pos"]).max() <= 1e-11 * scale and np.abs(got["dir"] - ref["dir"]).max() <= 1e-11
    ref, rst = prob.solve(rng=orc.RNG_PHILOX, seed=SEED)
    got, gst = gpu_ctx.solve(prob.desc, seed=SEED)
    assert (gst["emitted"], gst["steps"], gst["esc"]) == (rst["emitted"], rst["steps"], rst["esc"]) and gst["esc"] == 0
    assert np.array_equal(np.isfinite(got), np.isfinite(ref))          # cells outside the simplex: 0/0 in both (latent in the reference)
    fin = np.isfinite(ref)
    scale = np.abs(ref[fin]).max()
    assert np.abs(got[fin] - ref[fin]).max() <= 1e-9 * scale


def test_host_mirror_hex_and_pyr_domains_solve(matfiles, omats):
    """HexDomain / PyrDomain through the C++ mirror (Prism / Pyramid cells, Polygon<6> periodic faces)."""
    from montecarlocpp_b200 import hostapi
    hm = hostapi.Material(*matfiles["grey"])
    for k, dim, odom in (("hex", [1e-6, 5e-8, 8e-8, 3e-8], cases.hexd()), ("pyr", [1e-7, 1e-7, 1e-7], cases.pyr())):
        hp = hostapi.FieldProblem(hm, hostapi.Domain(k, dim, [], 1e6 * dim[0]), "multi", 20000, 30)
        got, gst = hp.solve_seeded(SEED)
        ref, rst = orc.Problem(omats["grey"], odom, "multi", 20000, 30).solve(rng=orc.RNG_PHILOX, seed=SEED)
        assert (gst["emitted"], gst["steps"], gst["esc"]) == (rst["emitted"], rst["steps"], 0)
        assert np.abs(got - ref).max() <= 1e-9 * np.abs(ref).max()
