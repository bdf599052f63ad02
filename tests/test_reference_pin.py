"""Pins the oracle restatement (oracle/mc_oracle.cpp) against the REFERENCE ITSELF.

oracle/_ref/montecarlo_ref is the reference's own sources (boundary / domain / field / main / material / phonon / problem /
random / subdomain .cpp, untouched, compiled where they lie) built against the Eigen/Boost stand-ins of oracle/shim/.  Its
stdout for the cases in tests/refcases.py (seed 0, one thread) is committed under tests/golden/ref_*.json
(tests/golden/make_ref_golden.py).  With the same mt19937 stream the oracle must reproduce every printed digit:
tolerance 2e-9 of the row scale (the reference prints 10 significant digits).

CPU only; no test here reads /root/reference.
"""
import json
import os

import numpy as np
import pytest

from oracle import pyoracle as orc, refbin
from tests import refcases

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 2e-9


def _golden(name):
    return json.load(open(os.path.join(GOLD, f"ref_{name}.json")))


def _oracle_solution(case, tmp_path):
    mat_kw, (disp, relax) = refcases.write_material(case, str(tmp_path))
    mat = orc.Material(disp, relax, case["T"])
    kind, dim, div, dT = case["odom"]
    dom = orc.Domain.create(kind, dim, div, dT)
    pk, nemit, size, maxscat, maxloop = case["prob"]
    prob = orc.Problem(mat, dom, pk, nemit, maxscat, maxloop, size)
    orc.set_arg_order(True)                    # g++ evaluates the three position draws right to left
    try:
        sol, st = prob.solve(rng=orc.RNG_MT19937, seed=0, nthreads=1)
    finally:
        orc.set_arg_order(False)
    return sol, st


def _close(got, ref):
    scale = np.abs(ref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    return np.abs(got - ref) / scale


@pytest.mark.parametrize("name", sorted(refcases.CASES))
def test_oracle_reproduces_reference_golden(name, tmp_path):
    case, g = refcases.CASES[name], _golden(name)
    ref = np.array(g["output"])
    sol, st = _oracle_solution(case, tmp_path)
    assert sol.shape == ref.shape
    assert st["esc"] == g["esc"]
    err = _close(sol, ref)
    assert err.max() <= TOL, f"{name}: max rel err {err.max():.3e} at {np.unravel_index(err.argmax(), err.shape)}"


def test_arg_order_is_the_only_difference(tmp_path):
    """With the default (left-to-right) draw order the oracle differs from the g++-built reference exactly where the
    swapped emission coordinates matter (film: q_z through the specular z faces), and nowhere else."""
    case, g = refcases.CASES["film_multi_si"], _golden("film_multi_si")
    ref = np.array(g["output"])
    mat_kw, (disp, relax) = refcases.write_material(case, str(tmp_path))
    mat = orc.Material(disp, relax, case["T"])
    dom = orc.Domain.create(*case["odom"])
    pk, nemit, size, maxscat, maxloop = case["prob"]
    sol, _ = orc.Problem(mat, dom, pk, nemit, maxscat, maxloop, size).solve(rng=orc.RNG_MT19937, seed=0, nthreads=1)
    err = _close(sol, ref)
    assert err[:3].max() <= TOL          # T, q_x, q_y do not depend on which of x / z got which draw
    assert err[3].max() > 1e-3           # q_z does


@pytest.mark.skipif(not refbin.available(), reason="oracle/_ref/montecarlo_ref not built (make -C oracle ref, dev container only)")
@pytest.mark.parametrize("name", ["film_multi_si", "tube_multi_si", "bulk_cumflux_grey"])
def test_live_reference_binary_matches_golden_and_oracle(name, tmp_path):
    case, g = refcases.CASES[name], _golden(name)
    mat_kw, _ = refcases.write_material(case, str(tmp_path))
    dom, prob = refcases.ref_argv(case)
    blocks, out, _ = refbin.run(str(tmp_path), mat_kw, case["T"], dom, prob, seed=0, threads=1)
    live = blocks["Output"][0]
    assert np.array_equal(live, np.array(g["output"]))          # optimised build == checked build that made the golden
    sol, _ = _oracle_solution(case, tmp_path)
    assert _close(sol, live).max() <= TOL


@pytest.mark.skipif(not refbin.available(), reason="oracle/_ref/montecarlo_ref not built")
def test_reference_binary_seed_changes_result(tmp_path):
    """The seed override of the random_device stand-in is honoured (different seed -> different Monte Carlo estimate)."""
    case = refcases.CASES["bulk_temp_grey"]
    mat_kw, _ = refcases.write_material(case, str(tmp_path))
    dom, prob = refcases.ref_argv(case)
    a = refbin.run(str(tmp_path), mat_kw, case["T"], dom, prob, seed=0)[0]["Output"][0]
    b = refbin.run(str(tmp_path), mat_kw, case["T"], dom, prob, seed=7)[0]["Output"][0]
    assert a.shape == b.shape and not np.array_equal(a, b)


@pytest.mark.parametrize("name", sorted(refcases.REF_ONLY))
def test_octet_goldens_are_well_formed(name):
    """OctetDomain is not restated; its reference output is kept as golden data for that work."""
    g = _golden(name)
    out, avg = np.array(g["output"]), np.array(g["averaged"])
    assert out.shape[0] == 4 and np.isfinite(out).all() and g["esc"] == 0
    assert avg.shape == (4, 1) and np.isfinite(avg).all()


# ---- the reference's objects behind oracle/ref_driver.cpp: full precision (17 digits), loop-trip counts, threads, and the
#      two domains BASELINE.json's configs need that the reference does not ship (composed from its own templates)
DRIVER_CASES = {
    # name: (material, domain kw, dim, div, dT, oracle domain builder, problem, nemit, maxscat, threads)
    "slab_multi": ("silicon", "slab", [100e-9] * 3, [20, 0, 0], 1.0, "multi", 30000, 1000, 1),
    "slab_multi_3thr": ("silicon", "slab", [100e-9] * 3, [20, 0, 0], 1.0, "multi", 30000, 1000, 3),
    "wire_multi": ("silicon", "wire", [1e-6, 1e-7, 1e-7], [0, 6, 6], 1.0, "multi", 20000, 100, 2),
    "film_flux": ("grey", "film", [1e-6, 1e-7, 1e-6], [0, 12, 0], 1.0, "flux", 20000, 100, 1),
    "bulk3d_temp": ("silicon", "bulk", [2e-7] * 3, [6, 5, 4], 0.2, "temp", 20000, 60, 2),
}


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built (make -C oracle ref, dev container only)")
@pytest.mark.parametrize("name", sorted(DRIVER_CASES))
def test_oracle_matches_reference_objects_full_precision(name, tmp_path):
    from montecarlocpp_b200 import abi, materials
    mkind, dk, dim, div, dT, pk, nemit, maxscat, thr = DRIVER_CASES[name]
    disp, relax = materials.write_grey(str(tmp_path)) if mkind == "grey" else materials.write_silicon(str(tmp_path), nw=200)
    ref = refbin.drive(disp, relax, 300.0, dk, dim, div, dT, pk, nemit, maxscat, seed=5, threads=thr)
    mat = orc.Material(disp, relax, 300.0)
    S, D, ISO, P = abi.BDRY_SPEC, abi.BDRY_DIFF, abi.BDRY_ISOT, abi.BDRY_PERI
    if dk == "slab":
        dom = orc.Domain.box([0, 0, 0], dim, div, [0, 0, 0], [ISO, S, S, ISO, S, S], [dT / 2, 0, 0, -dT / 2, 0, 0])
    elif dk == "wire":
        dom = orc.Domain.box([0, 0, 0], dim, div, [-dT / dim[0], 0, 0], [P, D, D, P, D, D])
    else:
        dom = orc.Domain.create(dk, dim, div, dT)
    prob = orc.Problem(mat, dom, pk, nemit, maxscat)
    orc.set_arg_order(True)
    try:
        sol, st = prob.solve(rng=orc.RNG_MT19937, seed=5, nthreads=thr)
    finally:
        orc.set_arg_order(False)
    assert ref["threads"] == thr
    assert st["steps"] == ref["steps"], "loop-trip count differs from the reference's"          # integer work: exact
    assert st["esc"] == ref["esc"]
    err = _close(sol, ref["output"])
    assert err.max() <= 1e-11, f"{name}: {err.max():.3e}"


# ---- the reference-side binding: descriptors flattened from the reference's OWN objects (ref_driver ... flatten) must equal
#      what the host mirror (montecarlocpp_b200/host, same class names) flattens for the same constructor arguments
FLATTEN_CASES = [("bulk", [1e-6] * 3, [8, 0, 0], 1.0), ("film", [1e-6, 1e-7, 1e-6], [0, 10, 0], 1.0),
                 ("jct", [1e-7, 1e-7, 1e-7, 5e-8], [2, 3, 2, 2], 0.2), ("tee", [1e-7, 1.2e-7, 1e-7, 0.9e-7, 5e-8], [3, 2, 3, 2, 0], 0.3),
                 ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 6, 6, 3], 1.0), ("slab", [1e-7] * 3, [20, 0, 0], 1.0),
                 ("wire", [1e-6, 1e-7, 1e-7], [0, 6, 6], 1.0),
                 ("hex", [1e-6, 5e-8, 8e-8, 3e-8], [], 1.0), ("pyr", [1e-7, 1e-7, 1e-7], [], 0.1),
                 # OctetDomain: the mirror evaluates its geometry table (linear forms derived from the reference's outputs,
                 # tools/derive_octet_table.py) in another order than the reference's hand-typed expressions: rounding-level bar
                 ("octet", [1e-6, 1e-7, 1e-7, 1e-8], [2, 2, 2, 1], 1.0), ("octet", [1.1e-6, 0.9e-7, 1.3e-7, 0.9e-8], [3, 2, 4, 2], 0.7)]
FLATTEN_IDS = [c[0] + ("" if [k[0] for k in FLATTEN_CASES[:i]].count(c[0]) == 0 else "2") for i, c in enumerate(FLATTEN_CASES)]


def _struct_diff(a, b, rtol=1e-15, scale=0.0):
    bad = []
    for f, _ in a._fields_:
        x, y = getattr(a, f), getattr(b, f)
        if hasattr(x, "__len__"):
            # rounding-level bar (octet): entries that cancel to zero are compared against the size of their array
            atol = 0 if rtol <= 1e-15 else rtol * max(1e-300, float(np.abs(np.array(y[:], dtype=float)).max()))
            if not np.allclose(np.array(x[:]), np.array(y[:]), rtol=rtol, atol=atol):
                bad.append(f)
        elif isinstance(x, (int, float)) and not (x == y or abs(x - y) <= rtol * max(abs(y), scale)):
            bad.append(f)
    return bad


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("kind,dim,div,dT", FLATTEN_CASES, ids=FLATTEN_IDS)
@pytest.mark.parametrize("pkind,size", [("multi", 0), ("cumflux", 4)])
def test_host_mirror_flattens_like_the_reference_objects(kind, dim, div, dT, pkind, size, tmp_path):
    rtol, scale = (1e-11, dim[0]) if kind == "octet" else (1e-15, 0.0)      # scale: plane offsets that cancel to ~0
    from montecarlocpp_b200 import hostapi, materials
    disp, relax = materials.write_silicon(str(tmp_path), nw=64)
    fl = refbin.flatten(disp, relax, 300.0, kind, dim, div, dT, pkind, 20000, 100, size=size, outdir=str(tmp_path))
    mat = hostapi.Material(disp, relax, 300.0)
    hd = hostapi.Domain(kind, dim, div, dT)
    hp = hostapi.FieldProblem(mat, hd, pkind, 20000, 100, size=size)
    d = hd.desc
    assert (fl.nsdom, fl.nplane, fl.npair, fl.nemitter, fl.cols) == (d.nsdom, d.nplane, d.npair, d.nemitter, d.ncols)
    for i in range(d.nsdom):
        assert _struct_diff(fl.sdoms[i], d.sdoms[i], rtol) == [], f"subdomain {i}"
    for i in range(d.nplane):
        assert _struct_diff(fl.planes[i], d.planes[i], rtol, scale) == [], f"plane {i}"
    assert list(fl.pairs[:fl.npair]) == list(d.pairs[:d.npair])                      # integer work: exact
    for i in range(d.nemitter):
        assert _struct_diff(fl.emitters[i], d.emitters[i], rtol) == [], f"emitter {i}"
    assert np.allclose(np.array(fl.cell_vol[:fl.cols]), np.ctypeslib.as_array(d.cell_vol, (d.ncols,)), rtol=rtol, atol=0)
    assert _struct_diff(fl.problem, hp.desc, rtol) == []
    if fl.weights is not None:                                  # OctetDomain::average: the WeightF weights (domain.cpp:1252-1280)
        sol = np.random.default_rng(3).normal(size=(hp.rows, fl.cols))
        assert np.allclose(hd.average(sol), fl.average(sol), rtol=1e-10, atol=0)
    assert list(fl.emit_count[:fl.nemitter]) == list(hp.emit_count())                # emitPdf_: exact


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
def test_octet_domain_flattens(tmp_path):
    """The reference's 42-subdomain OctetDomain through the binding: structure checks (the GPU run is in test_gpu_parity)."""
    from montecarlocpp_b200 import abi, materials
    disp, relax = materials.write_silicon(str(tmp_path), nw=64)
    fl = refbin.flatten(disp, relax, 300.0, "octet", [1e-6, 1e-7, 1e-7, 1e-8], [2, 2, 2, 1], 1.0, "multi", 4000, 50, outdir=str(tmp_path))
    assert fl.nsdom == 42 and fl.cols == 32 and fl.nemitter > 0
    kinds = {fl.sdoms[i].cell for i in range(42)}
    assert kinds == {abi.CELL_PARALLELEPIPED, abi.CELL_TRIPRISM, abi.CELL_PRISM, abi.CELL_PYRAMID} or len(kinds) >= 3
    for q in range(fl.nplane):                                             # every hand-off / periodic partner is mutual
        p = fl.planes[q]
        for k in range(p.pair_count):
            r = fl.planes[fl.pairs[p.pair_begin + k]]
            back = [fl.pairs[r.pair_begin + j] for j in range(r.pair_count)]
            assert q in back
            assert p.kind == r.kind and p.kind in (abi.BDRY_INTER, abi.BDRY_PERI)


# ---- N4: TrajProblem / the reference's `traj` mode (main.cpp:86-103, problem.cpp:226-299).  The reference prints the boundary
#      trace of every loop trip ("<sdom>: <in> <type> -> <out> <type>") and the polyline; the oracle's mt19937 trajectory must
#      reproduce both: integers exactly, points to the printed precision.
TRAJ_CASES = {
    "tube": (["tube", 1e-6, 5e-8, 2e-8, 4, 2], ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 4, 4, 2], 1.0), [5e-7, 6e-8, 1e-8], [0.3, 0.5, 0.81], 12),
    "jct": (["jct", 1e-7, 5e-8], ("jct", [1e-7, 1e-7, 1e-7, 5e-8], [0, 0, 0, 0], 2e6 * 1e-7), [5e-8, 5e-8, 2.5e-8], [0.2, 0.9, 0.1], 12),
    "film": (["film", 1e-6, 1e-7, 10], ("film", [1e-6, 1e-7, 1e-6], [0, 10, 0], 1.0), [9.5e-7, 5e-8, 5e-7], [0.9, 0.3, -0.2], 10),
}
_KIND_NAMES = {"Spec": 0, "Diff": 1, "Inter": 2, "Isot": 3, "Peri": 4, "Null": -1}


@pytest.mark.skipif(not refbin.available(), reason="oracle/_ref/montecarlo_ref not built")
@pytest.mark.parametrize("name", sorted(TRAJ_CASES))
def test_trajectory_mode_matches_the_reference_binary(name, tmp_path):
    import re
    from montecarlocpp_b200 import materials
    dom_argv, odom, pos, dirv, maxscat = TRAJ_CASES[name]
    disp, relax = materials.write_grey(str(tmp_path), inv_tau=2e10)             # mean free path 300 nm: mostly boundary events
    _, out, _ = refbin.run(str(tmp_path), "grey", 300.0, dom_argv, ["traj"] + pos + dirv + [maxscat, 0], seed=0, threads=1)
    body = out[out.index("Trajectory 0"):]
    trace = re.findall(r"^\s*(-?\d+):\s+(-?\d+)\s+(\w+?)[PT\d]?\s+->\s+(-?\d+)\s+(\w+?)[PT\d]?\s*$", body, re.M)
    rows = [ln.split() for ln in body.splitlines() if ln.strip() and re.fullmatch(r"[\s\d.eE+\-nan]+", ln) and "e" in ln.lower()]
    pts = np.array([[float(x) for x in r] for r in rows]).T                       # N x 3
    mat = orc.Material(disp, relax, 300.0)
    dom = orc.Domain.create(*odom)
    r = orc.traj(mat, dom, 0, maxscat, 0, pos=pos, dir=dirv, rng=orc.RNG_MT19937)
    assert len(trace) == len(r["step_sdom"]) and len(trace) > 3
    for k, (sd, bin_, tin, bout, tout) in enumerate(trace):
        assert int(sd) == r["step_sdom"][k] and int(bin_) == r["step_in"][k] and int(bout) == r["step_out"][k], k
        assert _KIND_NAMES[tin] == r["step_in_kind"][k] and _KIND_NAMES[tout] == r["step_out_kind"][k], k
    assert pts.shape == r["points"].shape
    assert np.allclose(pts, r["points"], rtol=2e-9, atol=1e-18)


# ---- N3: the non-box cells.  The reference only uses them inside OctetDomain (and in HexDomain / PyrDomain, whose constructors
#      forget to call init()); ref_driver wraps one cell of the reference's own templates in a domain (and calls init() for hex / pyr),
#      which pins the emission folds (subdomain.cpp:309-320, 351-377, 401-409, 433-441), the literal cellVol (:283-349) with its 0/0
#      columns, the Triangle / Polygon<N> emitters (boundary.cpp:182-187, 243-251) and the Polygon<6> periodic pair.
NONBOX_CASES = {
    # name: (origin, edge columns, div, gradT, dT, oracle domain builder)
    "triprism": ([1e-8, 0, -2e-8], [[1e-7, 0, 0], [2e-8, 1e-7, 0], [0, 1e-8, 2e-7]], [4, 4, 2], [-1e6, 5e5, 0], 1.0),
    "tet": ([0, 0, 0], [[1e-7, 0, 0], [1e-8, 1e-7, 0], [0, 2e-8, 1e-7]], [3, 3, 3], [0, 0, 0], 1.0),
    "prism5": ([0, 0, 0], [[1e-7, 0, 0], [0, 5e-8, -2e-8], [0, 9e-8, 1e-8], [0, 7e-8, 6e-8], [0, 0, 5e-8]], [0, 0, 0], [-1e6, 0, 0], 1.0),
}


def _compare_with_driver(name, dim, div, dT, dom, pk, tmp_path, tol=1e-11):
    from montecarlocpp_b200 import materials
    disp, relax = materials.write_silicon(str(tmp_path), nw=200)
    ref = refbin.drive(disp, relax, 300.0, name, dim, div, dT, pk, 20000, 60, seed=2, threads=1)
    prob = orc.Problem(orc.Material(disp, relax, 300.0), dom, pk, 20000, 60)
    orc.set_arg_order(True)
    try:
        sol, st = prob.solve(rng=orc.RNG_MT19937, seed=2, nthreads=1)
    finally:
        orc.set_arg_order(False)
    assert st["steps"] == ref["steps"] and st["esc"] == ref["esc"] == 0
    finite = np.isfinite(ref["output"])
    assert np.array_equal(np.isfinite(sol), finite)                   # cells outside a simplex: 0/0 in both (cellVol == 0)
    scale = np.abs(np.where(finite, ref["output"], 0.0)).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.0
    assert (np.abs(np.where(finite, sol - ref["output"], 0.0)) <= tol * scale).all()


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("pk", ["multi", "temp"])
@pytest.mark.parametrize("name", sorted(NONBOX_CASES))
def test_nonbox_cells_match_the_reference_templates(name, pk, tmp_path):
    from tests import cases
    o, cols, div, g, dT = NONBOX_CASES[name]
    dim = list(o) + [x for c in cols for x in c] + list(g)
    dom = cases.prism5(0) if name == "prism5" else cases.NONBOX[name]()
    _compare_with_driver(name, dim, div, dT, dom, pk, tmp_path)


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("name,dim,dT", [("hex", [1e-6, 5e-8, 8e-8, 3e-8], 1.0), ("pyr", [1e-7, 1e-7, 1e-7], 0.1)])
def test_hex_and_pyr_domains_match_the_reference(name, dim, dT, tmp_path):
    _compare_with_driver(name, dim, [], dT, orc.Domain.create(name, dim, [], dT), "multi", tmp_path)


@pytest.mark.skipif(not refbin.driver_available(), reason="oracle/_ref/ref_driver not built")
@pytest.mark.parametrize("name", sorted(refcases.REF_ONLY))
def test_octet_average_weights_reproduce_the_reference_averaged_block(name, tmp_path):
    """Domain::average for the octet truss: the WeightF weights exported by the binding, applied to the reference's printed
    Output, give the reference's printed Averaged block (domain.cpp:1252-1280)."""
    case, g = refcases.REF_ONLY[name], _golden(name)
    _, (disp, relax) = refcases.write_material(case, str(tmp_path))
    d = case["dom"]
    fl = refbin.flatten(disp, relax, case["T"], "octet", d[1:5], d[5:9], d[9], "multi", case["prob"][1], case["prob"][3], outdir=str(tmp_path))
    out, avg = np.array(g["output"]), np.array(g["averaged"])
    assert fl.weights is not None and fl.weights.shape == (out.shape[1],) and (fl.weights >= 0).all() and fl.weights.sum() > 0
    got = fl.average(out)
    assert got.shape == avg.shape
    assert np.allclose(got, avg, rtol=5e-9, atol=0)
