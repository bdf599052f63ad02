"""ctypes binding of the product library libmcb.so (C ABI of include/mcb.h).

Fails loudly when the CUDA library is missing or no sm_100 device is present: there is no
CPU fallback anywhere in the product path.
"""
import ctypes as C
import os

import numpy as np

from . import abi

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MCB_LIBMCB") or os.path.join(_HERE, "libmcb.so")     # MCB_LIBMCB: A/B a differently built library

SYMBOLS = [
    "mcb_create", "mcb_destroy", "mcb_last_error", "mcb_abi_version", "mcb_build_info", "mcb_set_options", "mcb_get_options",
    "mcb_upload_material", "mcb_upload_domain", "mcb_field_cols", "mcb_solve", "mcb_solve_raw_dev",
    "mcb_finalize_dev", "mcb_stream", "mcb_trace", "mcb_cell_index", "mcb_accumulate", "mcb_get_alias",
    "mcb_philox_words", "mcb_traj", "mcb_device_count", "mcb_solve_raw", "mcb_allreduce", "mcb_finalize", "mcb_sort_probe",
]

_lib = None


class McbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mcb error {code}: {msg}")
        self.code = code


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        vp, dp, ip, lp = C.c_void_p, abi.c_double_p, abi.c_int32_p, abi.c_int64_p
        L.mcb_create.argtypes = [C.c_int, C.POINTER(vp)]
        L.mcb_destroy.argtypes = [vp]; L.mcb_destroy.restype = None
        L.mcb_last_error.argtypes = [vp]; L.mcb_last_error.restype = C.c_char_p
        L.mcb_set_options.argtypes = [vp, C.POINTER(abi.Options)]
        L.mcb_get_options.argtypes = [vp, C.POINTER(abi.Options)]
        L.mcb_upload_material.argtypes = [vp, C.POINTER(abi.MaterialDesc)]
        L.mcb_upload_domain.argtypes = [vp, C.POINTER(abi.DomainDesc)]
        L.mcb_field_cols.argtypes = [vp, lp]
        L.mcb_solve.argtypes = [vp, C.POINTER(abi.ProblemDesc), C.c_uint64, C.c_int64, C.c_int64, dp, C.POINTER(abi.Stats)]
        L.mcb_solve_raw_dev.argtypes = [vp, C.POINTER(abi.ProblemDesc), C.c_uint64, C.c_int64, C.c_int64, vp, C.POINTER(abi.Stats)]
        L.mcb_finalize_dev.argtypes = [vp, C.POINTER(abi.ProblemDesc), vp]
        L.mcb_stream.argtypes = [vp, C.POINTER(vp)]
        L.mcb_trace.argtypes = [vp, C.POINTER(abi.ProblemDesc), C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.POINTER(abi.TraceOut)]
        L.mcb_traj.argtypes = [vp, C.POINTER(abi.TrajDesc), C.c_uint64, C.POINTER(abi.TrajOut)]
        L.mcb_sort_probe.argtypes = [vp, C.POINTER(abi.ProblemDesc), C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int32, lp, lp, lp, lp]
        L.mcb_cell_index.argtypes = [vp, C.c_int64, dp, ip, lp]
        L.mcb_accumulate.argtypes = [vp, C.c_int32, C.c_int64, ip, dp, dp, dp, dp]
        L.mcb_get_alias.argtypes = [vp, C.c_int, dp, ip, dp, ip]
        L.mcb_philox_words.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(abi.c_double_p)


class Context:
    """One device context: upload Material + Domain tables, then solve FieldProblems on them."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        rc = lib().mcb_create(device, C.byref(self.h))
        if rc != 0:
            raise McbError(rc, lib().mcb_last_error(None).decode())
        self._keep = []
        self.cols = None
        self.np_ = None

    def _check(self, rc):
        if rc != 0:
            raise McbError(rc, lib().mcb_last_error(self.h).decode())

    def close(self):
        if self.h:
            lib().mcb_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        o = abi.Options()
        self._check(lib().mcb_get_options(self.h, C.byref(o)))
        for k, v in kw.items():
            setattr(o, k, v)
        self._check(lib().mcb_set_options(self.h, C.byref(o)))

    def upload_material(self, desc):
        self._check(lib().mcb_upload_material(self.h, C.byref(desc)))
        self.nw, self.np_ = desc.nw, desc.np

    def upload_domain(self, desc):
        self._check(lib().mcb_upload_domain(self.h, C.byref(desc)))
        cols = C.c_int64()
        self._check(lib().mcb_field_cols(self.h, C.byref(cols)))
        self.cols = cols.value

    def solve(self, prob_desc, seed=0, n_begin=0, n_end=None):
        n_end = prob_desc.nemit if n_end is None else n_end
        out = np.zeros(prob_desc.rows * self.cols)
        st = abi.Stats()
        self._check(lib().mcb_solve(self.h, C.byref(prob_desc), seed, n_begin, n_end, _dp(out), C.byref(st)))
        return out.reshape(self.cols, prob_desc.rows).T.copy(), st.asdict()

    def solve_raw_dev(self, prob_desc, dev_ptr, seed=0, n_begin=0, n_end=None):
        n_end = prob_desc.nemit if n_end is None else n_end
        st = abi.Stats()
        self._check(lib().mcb_solve_raw_dev(self.h, C.byref(prob_desc), seed, n_begin, n_end, C.c_void_p(dev_ptr), C.byref(st)))
        return st.asdict()

    def finalize_dev(self, prob_desc, dev_ptr):
        self._check(lib().mcb_finalize_dev(self.h, C.byref(prob_desc), C.c_void_p(dev_ptr)))

    def stream(self):
        s = C.c_void_p()
        self._check(lib().mcb_stream(self.h, C.byref(s)))
        return s.value

    def trace(self, prob_desc, seed, n_begin, n_end, nsteps):
        bufs, out = abi.trace_buffers(n_end - n_begin)
        self._check(lib().mcb_trace(self.h, C.byref(prob_desc), seed, n_begin, n_end, nsteps, C.byref(out)))
        return bufs

    def sort_probe(self, prob_desc, seed, n_begin, n_end, nsteps, sorted_=True):
        """K3 probe: (bin, column, particle id) per slot after compaction / the counting sort, and columns per bin."""
        n = n_end - n_begin
        b, c, p = (np.zeros(n, np.int64) for _ in range(3))
        cpb = C.c_int64(0)
        self._check(lib().mcb_sort_probe(self.h, C.byref(prob_desc), seed, n_begin, n_end, nsteps, 1 if sorted_ else 0,
                                         b.ctypes.data_as(abi.c_int64_p), c.ctypes.data_as(abi.c_int64_p),
                                         p.ctypes.data_as(abi.c_int64_p), C.byref(cpb)))
        return b, c, p, cpb.value

    def traj(self, traj_desc, seed):
        bufs, out = abi.traj_buffers(traj_desc.maxloop)
        self._check(lib().mcb_traj(self.h, C.byref(traj_desc), seed, C.byref(out)))
        return abi.traj_result(bufs, out)

    def cell_index(self, pos, sdom):
        pos = np.ascontiguousarray(pos, np.float64); sdom = np.ascontiguousarray(sdom, np.int32)
        out = np.zeros((len(sdom), 3), np.int64)
        self._check(lib().mcb_cell_index(self.h, len(sdom), _dp(pos), sdom.ctypes.data_as(abi.c_int32_p),
                                         out.ctypes.data_as(abi.c_int64_p)))
        return out

    def accumulate(self, rows, sdom, bpos, epos, amount):
        sdom = np.ascontiguousarray(sdom, np.int32)
        bpos = np.ascontiguousarray(bpos, np.float64); epos = np.ascontiguousarray(epos, np.float64)
        amount = np.ascontiguousarray(amount, np.float64)
        field = np.zeros(rows * self.cols)
        self._check(lib().mcb_accumulate(self.h, rows, len(sdom), sdom.ctypes.data_as(abi.c_int32_p), _dp(bpos), _dp(epos),
                                         _dp(amount), _dp(field)))
        return field.reshape(self.cols, rows).T.copy()

    def alias(self, which):
        nw, npol = self.nw, self.np_
        wprob, walias = np.zeros(nw), np.zeros(nw, np.int32)
        pprob, palias = np.zeros(nw * npol), np.zeros(nw * npol, np.int32)
        self._check(lib().mcb_get_alias(self.h, which, _dp(wprob), walias.ctypes.data_as(abi.c_int32_p), _dp(pprob),
                                        palias.ctypes.data_as(abi.c_int32_p)))
        return wprob, walias, pprob, palias


def philox_words(seed, particle, event, block):
    out = (C.c_uint32 * 4)()
    rc = lib().mcb_philox_words(seed, particle, event, block, out)
    if rc != 0:
        raise McbError(rc, lib().mcb_last_error(None).decode())
    return [int(x) for x in out]
