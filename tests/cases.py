"""Shared problem zoo for the parity tests: every shipped box domain of the reference
(domain.cpp: Bulk, Film, Jct, Tee, Tube) plus the two box variants BASELINE.json's configs need
(isothermal-wall slab, diffuse wire)."""
import numpy as np

from montecarlocpp_b200 import abi
from oracle import pyoracle as orc

S, D, I, ISO, P = abi.BDRY_SPEC, abi.BDRY_DIFF, abi.BDRY_INTER, abi.BDRY_ISOT, abi.BDRY_PERI


def slab(L=100e-9, W=100e-9, ncell=20, dT=1.0):
    """C2: Parallelepiped<IsotP,Spec,Spec> between walls at +-dT/2."""
    return orc.Domain.box([0, 0, 0], [L, W, W], [ncell, 0, 0], [0, 0, 0], [ISO, S, S, ISO, S, S],
                          [dT / 2, 0, 0, -dT / 2, 0, 0])


def wire(L=1e-6, W=100e-9, n=8):
    """C3: Parallelepiped<PeriP,Diff,Diff> with a 2-D tally grid over the cross-section."""
    return orc.Domain.box([0, 0, 0], [L, W, W], [0, n, n], [-1e6, 0, 0], [P, D, D, P, D, D])


def skew(ncell=(3, 4, 5)):
    """A sheared parallelepiped (non-diagonal mat_) with a 3-D tally grid and mixed walls."""
    mat = np.array([[1e-7, 2e-8, 1e-8], [0, 1.2e-7, 3e-8], [0, 0, 0.9e-7]])
    return orc.Domain.box([1e-8, -2e-8, 3e-8], mat, list(ncell), [-2e6, 1e6, 5e5], [D, S, D, S, D, S])


def bulk(L=1e-6, div=(10, 0, 0)):
    return orc.Domain.create("bulk", [L, L, L], list(div), 1e6 * L)


def bulk64(L=1e-6):
    """Fine 3-D grid (>= 64 cells along x): takes the warp-cooperative N-D walk of the CUDA kernel."""
    return orc.Domain.create("bulk", [L, L, L], [64, 6, 5], 1e6 * L)


def film(L=1e-6, t=100e-9, ncell=20):
    return orc.Domain.create("film", [L, t, L], [0, ncell, 0], 1e6 * L)


def jct(a=1e-7, h=5e-8, div=(2, 3, 2, 2)):
    return orc.Domain.create("jct", [a, a, a, h], list(div), 2e6 * a)


def tee(a=1e-7, h=5e-8, div=(2, 2, 2, 2, 0)):
    return orc.Domain.create("tee", [a, a, a, a, h], list(div), 3e6 * a)


def tube(L=1e-6, a=5e-8, t=2e-8, div=(0, 8, 8, 4)):
    return orc.Domain.create("tube", [L, a, a, t], list(div), 1e6 * L)


def hexd(L=1e-6, a=5e-8, b=8e-8, c=3e-8):
    """HexDomain (domain.h:142): one hexagonal Prism, periodic along x, six specular sides."""
    return orc.Domain.create("hex", [L, a, b, c], [], 1e6 * L)


def pyr(a=1e-7):
    """PyrDomain (domain.h:165): one square Pyramid, all faces specular."""
    return orc.Domain.create("pyr", [a, a, a], [], 1e6 * a)


def triprism(div=(4, 4, 2)):
    """One gridded TriangularPrism (subdomain.h:206) with mixed diffuse / specular faces and a volumetric source."""
    return orc.Domain.cell(abi.CELL_TRIPRISM, [1e-8, 0, -2e-8], [[1e-7, 0, 0], [2e-8, 1e-7, 0], [0, 1e-8, 2e-7]], list(div),
                           [-1e6, 5e5, 0], [D, S, D, S, D])


def tet(div=(3, 3, 3)):
    """One gridded Tetrahedron (subdomain.h:289) with an isothermal emitting face (Triangle shape)."""
    return orc.Domain.cell(abi.CELL_TETRAHEDRON, [0, 0, 0], [[1e-7, 0, 0], [1e-8, 1e-7, 0], [0, 2e-8, 1e-7]], list(div),
                           [0, 0, 0], [ISO, S, D, D], [1.0, 0, 0, 0])


def prism5(div=0):
    """A 5-column Prism (pentagonal base) with isothermal Polygon faces top and bottom (Polygon emitters)."""
    cols = [[1e-7, 0, 0], [0, 5e-8, -2e-8], [0, 9e-8, 1e-8], [0, 7e-8, 6e-8], [0, 0, 5e-8]]
    return orc.Domain.cell(abi.CELL_PRISM, [0, 0, 0], cols, [div, 0, 0], [-1e6, 0, 0], [ISO, ISO, D, S, D, S, D], [0.5, -0.5, 0, 0, 0, 0, 0])


NONBOX = {"hex": hexd, "pyr": pyr, "triprism": triprism, "tet": tet, "prism5": prism5}
DOMAINS = {"bulk64": bulk64, "slab": slab, "wire": wire, "skew": skew, "bulk": bulk, "film": film, "jct": jct, "tee": tee, "tube": tube}


def upload(ctx, mat, dom):
    ctx.upload_material(mat.desc)
    ctx.upload_domain(dom.desc)
    return ctx
