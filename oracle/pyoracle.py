"""ctypes binding of the CPU oracle (oracle/libmc_oracle.so).  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from montecarlocpp_b200 import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmc_oracle.so")
RNG_MT19937, RNG_PHILOX = 0, 1


def build(force=False):
    src = os.path.join(_HERE, "mc_oracle.cpp")
    stale = (not os.path.exists(_LIB)) or (
        os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(_LIB))
    if (force or stale) and os.path.exists("/usr/bin/g++"):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        vp, dp, ip, lp = C.c_void_p, abi.c_double_p, abi.c_int32_p, abi.c_int64_p
        L.orc_last_error.restype = C.c_char_p
        L.orc_material_load.restype = vp
        L.orc_material_load.argtypes = [C.c_char_p, C.c_char_p, C.c_double]
        L.orc_material_free.argtypes = [vp]
        L.orc_material_desc.argtypes = [vp, C.POINTER(abi.MaterialDesc)]
        L.orc_material_cond.restype = C.c_double
        L.orc_material_cond.argtypes = [vp]
        L.orc_material_alias.argtypes = [vp, C.c_int, dp, ip, dp, ip]
        L.orc_domain_create.restype = vp
        L.orc_domain_create.argtypes = [C.c_char_p, dp, C.c_int, lp, C.c_int, C.c_double]
        L.orc_domain_box.restype = vp
        L.orc_domain_box.argtypes = [dp, dp, lp, dp, ip, dp]
        L.orc_domain_cell.restype = vp
        L.orc_domain_cell.argtypes = [C.c_int, dp, dp, C.c_int, lp, dp, ip, dp]
        L.orc_domain_free.argtypes = [vp]
        L.orc_domain_desc.argtypes = [vp, C.POINTER(abi.DomainDesc)]
        L.orc_domain_cols.restype = C.c_int64
        L.orc_domain_cols.argtypes = [vp]
        L.orc_domain_cell_vol.argtypes = [vp, dp]
        L.orc_problem_create.restype = vp
        L.orc_problem_create.argtypes = [vp, vp, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int64]
        L.orc_problem_free.argtypes = [vp]
        L.orc_problem_desc.argtypes = [vp, C.POINTER(abi.ProblemDesc)]
        for f in (L.orc_solve, L.orc_solve_raw):
            f.argtypes = [vp, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int, dp, C.POINTER(abi.Stats)]
        L.orc_finalize.argtypes = [vp, dp]
        L.orc_trace.argtypes = [vp, C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(abi.TraceOut)]
        L.orc_cell_index.argtypes = [vp, C.c_int64, dp, ip, lp]
        L.orc_traj.argtypes = [vp, vp, C.POINTER(abi.TrajDesc), C.c_uint64, C.POINTER(abi.TrajOut)]
        L.orc_traj_rng.argtypes = [vp, vp, C.POINTER(abi.TrajDesc), C.c_int, C.c_uint64, C.POINTER(abi.TrajOut)]
        L.orc_domain_locate.argtypes = [vp, dp]
        L.orc_accumulate.argtypes = [vp, C.c_int32, C.c_int64, ip, dp, dp, dp, dp]
        L.orc_philox_words.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        L.orc_mt_draws.argtypes = [C.c_uint32, C.c_int, C.c_int64, C.c_int64, dp]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_arg_order.argtypes = [C.c_int]
        _lib = L
    return _lib


def _err():
    return lib().orc_last_error().decode()


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


class Material:
    def __init__(self, disp, relax, temp=300.0):
        self.h = lib().orc_material_load(disp.encode(), relax.encode(), float(temp))
        if not self.h:
            raise RuntimeError(_err())
        self.desc = abi.MaterialDesc()
        lib().orc_material_desc(self.h, C.byref(self.desc))

    @property
    def nw(self): return self.desc.nw
    @property
    def np_(self): return self.desc.np
    def cond(self): return lib().orc_material_cond(self.h)
    def table(self, name):
        n = self.desc.nw * self.desc.np
        return np.ctypeslib.as_array(getattr(self.desc, name), shape=(n,)).copy()
    def alias(self, which):
        nw, npol = self.desc.nw, self.desc.np
        wprob, walias = np.zeros(nw), np.zeros(nw, np.int32)
        pprob, palias = np.zeros(nw * npol), np.zeros(nw * npol, np.int32)
        lib().orc_material_alias(self.h, which, _dp(wprob), walias.ctypes.data_as(abi.c_int32_p),
                                 _dp(pprob), palias.ctypes.data_as(abi.c_int32_p))
        return wprob, walias, pprob, palias
    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_material_free(self.h); self.h = None


class Domain:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError(_err())
        self.h = handle
        self.desc = abi.DomainDesc()
        lib().orc_domain_desc(self.h, C.byref(self.desc))
        self.cols = lib().orc_domain_cols(self.h)

    @classmethod
    def create(cls, kind, dim, div, dT):
        dim = np.ascontiguousarray(dim, np.float64)
        div = np.ascontiguousarray(div, np.int64)
        return cls(lib().orc_domain_create(kind.encode(), _dp(dim), len(dim),
                                           div.ctypes.data_as(abi.c_int64_p), len(div), float(dT)))

    @classmethod
    def box(cls, origin, mat, div, grad_t, kinds, T=(0,) * 6):
        origin = np.ascontiguousarray(origin, np.float64)
        mat = np.asarray(mat, np.float64)
        if mat.shape == (3,):
            mat = np.diag(mat)
        matc = np.ascontiguousarray(mat.T.reshape(-1))        # column-major
        div = np.ascontiguousarray(div, np.int64)
        g = np.ascontiguousarray(grad_t, np.float64)
        k = np.ascontiguousarray(kinds, np.int32)
        t = np.ascontiguousarray(T, np.float64)
        return cls(lib().orc_domain_box(_dp(origin), _dp(matc), div.ctypes.data_as(abi.c_int64_p), _dp(g),
                                        k.ctypes.data_as(abi.c_int32_p), _dp(t)))

    @classmethod
    def cell(cls, cell, origin, cols, div, grad_t, kinds, T=None):
        """One non-box cell: cols = list of 3-vectors (mat columns)."""
        origin = np.ascontiguousarray(origin, np.float64)
        c = np.ascontiguousarray(np.asarray(cols, np.float64).reshape(-1))
        div = np.ascontiguousarray(div, np.int64); g = np.ascontiguousarray(grad_t, np.float64)
        k = np.ascontiguousarray(kinds, np.int32)
        t = np.ascontiguousarray(T if T is not None else [0.0] * len(k), np.float64)
        return cls(lib().orc_domain_cell(cell, _dp(origin), _dp(c), len(c) // 3, div.ctypes.data_as(abi.c_int64_p), _dp(g),
                                         k.ctypes.data_as(abi.c_int32_p), _dp(t)))

    def cell_vol(self):
        v = np.zeros(self.cols)
        lib().orc_domain_cell_vol(self.h, _dp(v))
        return v

    def cell_index(self, pos, sdom):
        pos = np.ascontiguousarray(pos, np.float64); sdom = np.ascontiguousarray(sdom, np.int32)
        out = np.zeros((len(sdom), 3), np.int64)
        lib().orc_cell_index(self.h, len(sdom), _dp(pos), sdom.ctypes.data_as(abi.c_int32_p),
                             out.ctypes.data_as(abi.c_int64_p))
        return out

    def accumulate(self, rows, sdom, bpos, epos, amount):
        sdom = np.ascontiguousarray(sdom, np.int32)
        bpos = np.ascontiguousarray(bpos, np.float64); epos = np.ascontiguousarray(epos, np.float64)
        amount = np.ascontiguousarray(amount, np.float64)
        field = np.zeros(rows * self.cols)
        lib().orc_accumulate(self.h, rows, len(sdom), sdom.ctypes.data_as(abi.c_int32_p), _dp(bpos), _dp(epos),
                             _dp(amount), _dp(field))
        return field.reshape(self.cols, rows).T.copy()

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_domain_free(self.h); self.h = None


class Problem:
    def __init__(self, mat, dom, kind, nemit, maxscat, maxloop=0, size=0):
        self.mat, self.dom = mat, dom           # keep alive
        k = abi.PROB_KINDS[kind] if isinstance(kind, str) else kind
        self.h = lib().orc_problem_create(mat.h, dom.h, k, nemit, size, maxscat, maxloop)
        if not self.h:
            raise RuntimeError(_err())
        self.desc = abi.ProblemDesc()
        lib().orc_problem_desc(self.h, C.byref(self.desc))

    @property
    def rows(self): return self.desc.rows
    @property
    def nemit(self): return self.desc.nemit
    def emit_count(self):
        return np.ctypeslib.as_array(self.desc.emit_count, shape=(self.dom.desc.nemitter,)).copy()

    def solve(self, rng=RNG_PHILOX, seed=0, n_begin=0, n_end=None, nthreads=0, raw=False):
        n_end = self.nemit if n_end is None else n_end
        out = np.zeros(self.rows * self.dom.cols)
        st = abi.Stats()
        fn = lib().orc_solve_raw if raw else lib().orc_solve
        rc = fn(self.h, rng, seed, n_begin, n_end, nthreads, _dp(out), C.byref(st))
        if rc != 0:
            raise RuntimeError(_err())
        return out.reshape(self.dom.cols, self.rows).T.copy(), st.asdict()

    def finalize(self, raw):
        f = np.array(raw.T.reshape(-1), dtype=np.float64, copy=True)     # never alias the caller's array
        lib().orc_finalize(self.h, _dp(f))
        return f.reshape(self.dom.cols, self.rows).T.copy()

    def trace(self, seed, n_begin, n_end, nsteps):
        bufs, out = abi.trace_buffers(n_end - n_begin)
        rc = lib().orc_trace(self.h, seed, n_begin, n_end, nsteps, C.byref(out))
        if rc != 0:
            raise RuntimeError(_err())
        return bufs

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_problem_free(self.h); self.h = None


def traj(mat, dom, seed, maxscat, maxloop=0, prop=None, pos=None, dir=None, rng=RNG_PHILOX):
    """TrajProblem(mat, dom, [prop], [pos], [dir], maxscat, maxloop).solve() -> dict of records.
    rng=RNG_MT19937: one sequential mt19937(seed), the reference's own stream (main.cpp:86-103)."""
    t = make_traj_desc(dom, maxscat, maxloop, prop, pos, dir, lambda q: lib().orc_domain_locate(dom.h, _dp(q)))
    bufs, out = abi.traj_buffers(t.maxloop)
    if lib().orc_traj_rng(mat.h, dom.h, C.byref(t), rng, seed, C.byref(out)) != 0:
        raise RuntimeError(_err())
    return abi.traj_result(bufs, out)


def make_traj_desc(dom, maxscat, maxloop, prop, pos, dir, locate):
    t = abi.TrajDesc()
    t.maxscat, t.maxloop = maxscat, (maxloop if maxloop else 100 * maxscat)
    t.sdom = -1
    if prop is not None:
        t.has_prop, t.w, t.p = 1, prop[0], prop[1]
    if pos is not None:
        q = np.ascontiguousarray(pos, np.float64)
        t.has_pos = 1; t.pos[:] = list(q); t.sdom = locate(q)
        if dir is not None:
            t.has_dir = 1; t.dir[:] = list(dir)
    return t


def philox_words(seed, particle, event, block):
    out = (C.c_uint32 * 4)()
    lib().orc_philox_words(seed, particle, event, block, out)
    return [int(x) for x in out]


def mt_draws(seed, which, n, m=0):
    out = np.zeros(n)
    lib().orc_mt_draws(seed, which, m, n, _dp(out))
    return out


def max_threads():
    return lib().orc_max_threads()


def set_arg_order(right_to_left):
    """Order of the three position draws at subdomain.cpp:279/:312/:354 (unspecified in C++): False = x,y,z (default, the
    GPU path's order), True = z,y,x (what g++ does for the reference binary oracle/_ref/montecarlo_ref)."""
    lib().orc_set_arg_order(1 if right_to_left else 0)
