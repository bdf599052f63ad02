"""ctypes binding of libmcbhost.so — the C++ host mirror of the reference's Material / Domain /
FieldProblem API (montecarlocpp_b200/host).  Used by bench.py and the tests so that they construct the
same objects, and call the same `FieldProblem::solve`, a C++ user of the reference would.
"""
import ctypes as C
import os

import numpy as np

from . import abi, capi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcbhost.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        capi.lib()          # libmcb.so first (fails loudly when it is missing)
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run __graft_entry__.build()")
        L = C.CDLL(LIB_PATH)
        vp, dp, lp = C.c_void_p, abi.c_double_p, abi.c_int64_p
        L.mcbh_last_error.restype = C.c_char_p
        L.mcbh_material_create.restype = vp
        L.mcbh_material_create.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
        L.mcbh_material_free.argtypes = [vp]
        L.mcbh_material_desc.argtypes = [vp, C.POINTER(abi.MaterialDesc)]
        L.mcbh_material_cond.restype = C.c_double
        L.mcbh_material_cond.argtypes = [vp]
        L.mcbh_domain_create.restype = vp
        L.mcbh_domain_create.argtypes = [C.c_char_p, dp, C.c_int, lp, C.c_int, C.c_double]
        L.mcbh_domain_free.argtypes = [vp]
        L.mcbh_domain_desc.argtypes = [vp, C.POINTER(abi.DomainDesc)]
        L.mcbh_domain_cols.restype = C.c_int64
        L.mcbh_domain_cols.argtypes = [vp]
        L.mcbh_domain_average.restype = C.c_int64
        L.mcbh_domain_average.argtypes = [vp, dp, C.c_int64, dp]
        L.mcbh_problem_create.restype = vp
        L.mcbh_problem_create.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.mcbh_problem_free.argtypes = [vp]
        L.mcbh_problem_desc.argtypes = [vp, C.POINTER(abi.ProblemDesc)]
        L.mcbh_problem_solve.argtypes = [vp, C.c_int, C.c_uint32, dp, C.POINTER(abi.Stats)]
        L.mcbh_problem_solve_seeded.argtypes = [vp, C.c_int, C.c_uint64, C.c_int64, C.c_int64, dp, C.POINTER(abi.Stats)]
        L.mcbh_problem_solve_omp.argtypes = [vp, C.c_uint32, C.c_int, dp, C.POINTER(abi.Stats), lp]
        L.mcbh_set_devices.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.mcbh_simplex_cell_vol.restype = C.c_double
        L.mcbh_simplex_cell_vol.argtypes = [C.c_int, lp, lp, C.c_double]
        _lib = L
    return _lib


def _err():
    return lib().mcbh_last_error().decode()


class Material:
    """Material(disp, relax, temp) — material.h:23-68."""

    def __init__(self, disp, relax, temp=300.0):
        self.h = lib().mcbh_material_create(disp.encode(), relax.encode(), float(temp))
        if not self.h:
            raise RuntimeError(_err())
        self.desc = abi.MaterialDesc()
        lib().mcbh_material_desc(self.h, C.byref(self.desc))

    def cond(self):
        return lib().mcbh_material_cond(self.h)

    def table(self, name):
        return np.ctypeslib.as_array(getattr(self.desc, name), shape=(self.desc.nw * self.desc.np,)).copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().mcbh_material_free(self.h); self.h = None


class Domain:
    """BulkDomain / FilmDomain / JctDomain / TeeDomain / TubeDomain (domain.h) + SlabDomain / WireDomain."""

    def __init__(self, kind, dim, div, dT):
        dim = np.ascontiguousarray(dim, np.float64); div = np.ascontiguousarray(div, np.int64)
        self.h = lib().mcbh_domain_create(kind.encode(), dim.ctypes.data_as(abi.c_double_p), len(dim),
                                          div.ctypes.data_as(abi.c_int64_p), len(div), float(dT))
        if not self.h:
            raise RuntimeError(_err())
        self.desc = abi.DomainDesc()
        lib().mcbh_domain_desc(self.h, C.byref(self.desc))
        self.cols = lib().mcbh_domain_cols(self.h)

    def average(self, sol):
        """Domain::average (domain.cpp:78-81; OctetDomain: rows x 1 weighted mean, domain.cpp:1252-1257)."""
        sol = np.asarray(sol, np.float64)
        flat = np.ascontiguousarray(sol.T).ravel()               # column-major like ArrayXXd
        out = np.zeros(flat.size)
        n = lib().mcbh_domain_average(self.h, flat.ctypes.data_as(abi.c_double_p), sol.shape[0], out.ctypes.data_as(abi.c_double_p))
        if n < 0:
            raise RuntimeError(_err())
        return out[:sol.shape[0] * n].reshape(n, sol.shape[0]).T.copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().mcbh_domain_free(self.h); self.h = None


class FieldProblem:
    """TempProblem / FluxProblem / MultiProblem / CumTempProblem / CumFluxProblem (problem.h:121-239)."""

    def __init__(self, mat, dom, kind, nemit, maxscat, maxloop=0, size=0):
        self.mat, self.dom = mat, dom
        k = abi.PROB_KINDS[kind] if isinstance(kind, str) else kind
        self.h = lib().mcbh_problem_create(mat.h, dom.h, k, nemit, size, maxscat, maxloop)
        if not self.h:
            raise RuntimeError(_err())
        self.desc = abi.ProblemDesc()
        lib().mcbh_problem_desc(self.h, C.byref(self.desc))

    @property
    def rows(self): return self.desc.rows
    @property
    def nemit(self): return self.desc.nemit

    def emit_count(self):
        return np.ctypeslib.as_array(self.desc.emit_count, shape=(self.dom.desc.nemitter,)).copy()

    def solve(self, mt_seed=0, device=0):
        """FieldProblem::solve(Rng& gen, Progress*) with gen = mt19937(mt_seed): host buffers in and out."""
        out = np.zeros(self.rows * self.dom.cols); st = abi.Stats()
        if lib().mcbh_problem_solve(self.h, device, mt_seed, out.ctypes.data_as(abi.c_double_p), C.byref(st)) != 0:
            raise RuntimeError(_err())
        return out.reshape(self.dom.cols, self.rows).T.copy(), st.asdict()

    def solve_seeded(self, seed, n_begin=0, n_end=None, device=0):
        n_end = self.nemit if n_end is None else n_end
        out = np.zeros(self.rows * self.dom.cols); st = abi.Stats()
        if lib().mcbh_problem_solve_seeded(self.h, device, seed, n_begin, n_end, out.ctypes.data_as(abi.c_double_p), C.byref(st)) != 0:
            raise RuntimeError(_err())
        return out.reshape(self.dom.cols, self.rows).T.copy(), st.asdict()

    def solve_omp(self, mt_seed=0, nthreads=1):
        """The reference's calling pattern (main.cpp:155-166): solve() from every thread of an OpenMP region, partials summed.
        Returns (field, stats, Progress::count())."""
        out = np.zeros(self.rows * self.dom.cols); st = abi.Stats(); cnt = C.c_int64()
        if lib().mcbh_problem_solve_omp(self.h, mt_seed, nthreads, out.ctypes.data_as(abi.c_double_p), C.byref(st), C.byref(cnt)) != 0:
            raise RuntimeError(_err())
        return out.reshape(self.dom.cols, self.rows).T.copy(), st.asdict(), cnt.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().mcbh_problem_free(self.h); self.h = None


def set_devices(ordinals=()):
    """Process-wide device selection of FieldProblem::solve (empty: every visible sm_100 device).  With several devices the
    particle range is sharded over them and the raw tallies are summed with NCCL (mcb_allreduce)."""
    arr = (C.c_int * max(1, len(ordinals)))(*ordinals)
    n = lib().mcbh_set_devices(arr, len(ordinals))
    if n < 0:
        raise RuntimeError(_err())
    return n


def device_count():
    n = C.c_int()
    capi.lib().mcb_device_count(C.byref(n))
    return n.value


def simplex_cell_vol(cell, index, shape, vol):
    """TriangularPrismImpl::cellVol / TetrahedronImpl::cellVol of the host mirror."""
    i = np.ascontiguousarray(index, np.int64); s = np.ascontiguousarray(shape, np.int64)
    return lib().mcbh_simplex_cell_vol(cell, i.ctypes.data_as(abi.c_int64_p), s.ctypes.data_as(abi.c_int64_p), float(vol))
