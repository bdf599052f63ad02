"""Cases run through the REFERENCE ITSELF (oracle/_ref/montecarlo_ref, the reference's sources compiled against the
Eigen/Boost stand-ins in oracle/shim/) to pin the oracle restatement.  Each case is the reference's own argv
(main.cpp:216-235) plus the equivalent oracle constructor arguments (the vectors main.cpp:285-362 builds from argv).

material: ("silicon" | "grey", nw) -> files written by montecarlocpp_b200.materials in the reference's on-disk format.
"""

CASES = {
    # name: dict(material, T, dom=(ref argv), odom=(kind, dim, div, dT), prob=(kind, nemit, size, maxscat, maxloop))
    "film_multi_si": dict(material=("silicon", 200), T=300.0, dom=["film", 1e-6, 1e-7, 10],
                          odom=("film", [1e-6, 1e-7, 1e-6], [0, 10, 0], 1e6 * 1e-6), prob=("multi", 20000, 0, 100, 0)),
    "bulk_temp_grey": dict(material=("grey", 0), T=300.0, dom=["bulk", 1e-6, 16],
                           odom=("bulk", [1e-6] * 3, [16, 0, 0], 1e6 * 1e-6), prob=("temp", 20000, 0, 50, 0)),
    "bulk_flux_si": dict(material=("silicon", 200), T=250.0, dom=["bulk", 2e-7, 8],
                         odom=("bulk", [2e-7] * 3, [8, 0, 0], 1e6 * 2e-7), prob=("flux", 20000, 0, 100, 0)),
    "jct_multi_si": dict(material=("silicon", 200), T=300.0, dom=["jct", 1e-7, 5e-8],
                         odom=("jct", [1e-7, 1e-7, 1e-7, 5e-8], [0, 0, 0, 0], 2e6 * 1e-7), prob=("multi", 20000, 0, 100, 0)),
    "tee_multi_si": dict(material=("silicon", 200), T=300.0, dom=["tee", 1e-7, 5e-8, 3],
                         odom=("tee", [1e-7, 1e-7, 1e-7, 1e-7, 5e-8], [3, 3, 3, 3, 0], 3e6 * 1e-7), prob=("multi", 20000, 0, 100, 0)),
    "tube_multi_si": dict(material=("silicon", 200), T=300.0, dom=["tube", 1e-6, 5e-8, 2e-8, 6, 3],
                          odom=("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 6, 6, 3], 1e6 * 1e-6), prob=("multi", 20000, 0, 100, 0)),
    "film_cumtemp_si": dict(material=("silicon", 200), T=300.0, dom=["film", 1e-6, 2e-7, 8],
                            odom=("film", [1e-6, 2e-7, 1e-6], [0, 8, 0], 1e6 * 1e-6), prob=("cumtemp", 10000, 5, 40, 0)),
    "bulk_cumflux_grey": dict(material=("grey", 0), T=300.0, dom=["bulk", 5e-7, 4],
                              odom=("bulk", [5e-7] * 3, [4, 0, 0], 1e6 * 5e-7), prob=("cumflux", 10000, 4, 30, 0)),
    "film_multi_maxloop": dict(material=("silicon", 200), T=300.0, dom=["film", 1e-6, 5e-8, 5],
                               odom=("film", [1e-6, 5e-8, 1e-6], [0, 5, 0], 1e6 * 1e-6), prob=("multi", 10000, 0, 1000, 25)),
}
# Run through the reference only (no oracle restatement of the 42-subdomain octet truss exists): kept as golden output
# for the day OctetDomain is built.
REF_ONLY = {
    "octet_multi_si": dict(material=("silicon", 200), T=300.0, dom=["octet", 1e-6, 1e-7, 1e-7, 1e-8, 0, 0, 0, 0, 1.0],
                           prob=("multi", 4000, 0, 50, 0)),
    "octet_grid_multi_si": dict(material=("silicon", 200), T=300.0, dom=["octet", 1e-6, 1e-7, 1e-7, 1e-8, 2, 2, 2, 1, 1.0],
                                prob=("multi", 4000, 0, 50, 0)),
}


def ref_argv(case):
    kind, nemit, size, maxscat, maxloop = case["prob"]
    p = [kind, nemit] + ([size] if kind in ("cumtemp", "cumflux") else []) + [maxscat, maxloop, 1]
    return case["dom"], p


def write_material(case, dirname):
    """-> (reference material keyword, (disp, relax))"""
    from montecarlocpp_b200 import materials
    kind, nw = case["material"]
    if kind == "grey":
        return "grey", materials.write_grey(dirname)
    return "silicon", materials.write_silicon(dirname, nw=nw)
