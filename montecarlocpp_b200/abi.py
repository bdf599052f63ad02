"""ctypes mirror of include/mcb.h (the C ABI structs).

Plumbing only: the structs are laid out exactly as the header declares them, so the
same Python objects can be handed to libmcb.so (the product) and, in tests, filled by
the CPU oracle.
"""
import ctypes as C

MCB_ABI_VERSION = 2
MCB_OK, MCB_EINVAL, MCB_ENODEVICE, MCB_ECUDA, MCB_ESTATE, MCB_ELIMIT = 0, -1, -2, -3, -4, -5
BDRY_SPEC, BDRY_DIFF, BDRY_INTER, BDRY_ISOT, BDRY_PERI = range(5)
SHAPE_NONE, SHAPE_PARALLELOGRAM, SHAPE_TRIANGLE, SHAPE_POLYGON = range(4)
CELL_PARALLELEPIPED, CELL_TRIPRISM, CELL_TETRAHEDRON, CELL_PRISM, CELL_PYRAMID = range(5)
EMIT_SDOM, EMIT_BDRY = 0, 1
PROB_TEMP, PROB_FLUX, PROB_MULTI, PROB_CUMTEMP, PROB_CUMFLUX = range(5)
PROB_KINDS = {"temp": PROB_TEMP, "flux": PROB_FLUX, "multi": PROB_MULTI,
              "cumtemp": PROB_CUMTEMP, "cumflux": PROB_CUMFLUX}
MAX_VERTS, MAX_BASE = 8, 9

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class MaterialDesc(C.Structure):
    _fields_ = [("nw", C.c_int64), ("np", C.c_int64), ("temp", C.c_double),
                ("vel", c_double_p), ("tau", c_double_p),
                ("flux_pdf", c_double_p), ("scat_pdf", c_double_p),
                ("energy_sum", C.c_double), ("flux_sum", C.c_double), ("scat_sum", C.c_double)]


class PlaneDesc(C.Structure):
    _fields_ = [("normal", C.c_double * 3), ("offset", C.c_double),
                ("kind", C.c_int32), ("sdom", C.c_int32),
                ("pair_begin", C.c_int32), ("pair_count", C.c_int32),
                ("rot", C.c_double * 9), ("peri_rot", C.c_double * 9),
                ("peri_transl", C.c_double * 3), ("T", C.c_double),
                ("origin", C.c_double * 3), ("shape", C.c_int32), ("nvert", C.c_int32),
                ("verts", C.c_double * (3 * MAX_VERTS))]


class SdomDesc(C.Structure):
    _fields_ = [("origin", C.c_double * 3), ("mat", C.c_double * 9), ("inv", C.c_double * 9),
                ("div", C.c_int64 * 3), ("shape", C.c_int64 * 3), ("max", C.c_int64 * 3),
                ("accum", C.c_int32), ("cell", C.c_int32),
                ("eps", C.c_double), ("vol", C.c_double),
                ("grad_t", C.c_double * 3), ("emit_rot", C.c_double * 9),
                ("plane_begin", C.c_int32), ("plane_count", C.c_int32),
                ("nbase", C.c_int32), ("pad_", C.c_int32),
                ("base", C.c_double * (3 * MAX_BASE))]


class EmitterDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("index", C.c_int32), ("weight", C.c_double)]


class DomainDesc(C.Structure):
    _fields_ = [("nsdom", C.c_int32), ("sdoms", C.POINTER(SdomDesc)),
                ("nplane", C.c_int32), ("planes", C.POINTER(PlaneDesc)),
                ("npair", C.c_int32), ("pairs", c_int32_p),
                ("nemitter", C.c_int32), ("emitters", C.POINTER(EmitterDesc)),
                ("ncols", C.c_int64), ("cell_vol", c_double_p)]


class ProblemDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("rows", C.c_int32),
                ("size", C.c_int64), ("step", C.c_int64), ("nemit", C.c_int64),
                ("maxscat", C.c_int64), ("maxloop", C.c_int64),
                ("power", C.c_double), ("emit_count", c_int64_p)]


class Stats(C.Structure):
    _fields_ = [("emitted", C.c_int64), ("steps", C.c_int64), ("esc", C.c_int64),
                ("launches", C.c_int64), ("cols", C.c_int64),
                ("device_ms", C.c_double), ("step_ms", C.c_double),
                ("step_launches", C.c_int64), ("slot_steps", C.c_int64), ("state_stores", C.c_int64),
                ("steady_launches", C.c_int64), ("steady_steps", C.c_int64), ("steady_stores", C.c_int64),
                ("steady_ms", C.c_double), ("compactions", C.c_int64), ("sorts", C.c_int64),
                ("tail_ms", C.c_double), ("tail_steps", C.c_int64)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Options(C.Structure):
    _fields_ = [("slots", C.c_int64), ("steps_per_launch", C.c_int32), ("block", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("tally_mode", C.c_int32), ("decay_mode", C.c_int32),
                ("emit_mode", C.c_int32), ("compact_pct", C.c_int32), ("sort_mode", C.c_int32), ("decay_pct", C.c_int32), ("pad_", C.c_int32)]


class TraceOut(C.Structure):
    _fields_ = [("pos", c_double_p), ("dir", c_double_p), ("scat_next", c_double_p),
                ("w", c_int64_p), ("p", c_int64_p), ("sign", c_int32_p), ("alive", c_int32_p),
                ("sdom", c_int32_p), ("nscat", c_int64_p), ("steps", c_int64_p), ("cell", c_int32_p)]


class TrajDesc(C.Structure):
    _fields_ = [("has_prop", C.c_int32), ("has_pos", C.c_int32), ("has_dir", C.c_int32), ("sdom", C.c_int32),
                ("w", C.c_int64), ("p", C.c_int64), ("pos", C.c_double * 3), ("dir", C.c_double * 3),
                ("maxscat", C.c_int64), ("maxloop", C.c_int64)]


class TrajOut(C.Structure):
    _fields_ = [("max_points", C.c_int64), ("points", c_double_p), ("max_steps", C.c_int64),
                ("step_sdom", c_int32_p), ("step_in", c_int32_p), ("step_in_kind", c_int32_p),
                ("step_out", c_int32_p), ("step_out_kind", c_int32_p),
                ("npoints", C.c_int64), ("nsteps", C.c_int64), ("escaped", C.c_int32), ("pad_", C.c_int32)]


def traj_buffers(maxloop):
    """Buffers for one trajectory of up to `maxloop` loop trips and the TrajOut pointing at them."""
    import numpy as np
    bufs = {"points": np.zeros((2 * maxloop + 1, 3), np.float64)}
    for k in ("step_sdom", "step_in", "step_in_kind", "step_out", "step_out_kind"):
        bufs[k] = np.full(max(maxloop, 1), -9, np.int32)
    out = TrajOut()
    out.max_points, out.max_steps = 2 * maxloop + 1, max(maxloop, 1)
    out.points = bufs["points"].ctypes.data_as(c_double_p)
    for k in ("step_sdom", "step_in", "step_in_kind", "step_out", "step_out_kind"):
        setattr(out, k, bufs[k].ctypes.data_as(c_int32_p))
    return bufs, out


def traj_result(bufs, out):
    n, m = out.npoints, out.nsteps
    r = {"points": bufs["points"][:n].copy(), "escaped": bool(out.escaped)}
    for k in ("step_sdom", "step_in", "step_in_kind", "step_out", "step_out_kind"):
        r[k] = bufs[k][:m].copy()
    return r


def trace_buffers(n):
    """Allocate numpy buffers for an n-particle trace and the TraceOut that points at them."""
    import numpy as np
    bufs = {
        "pos": np.zeros((n, 3), np.float64), "dir": np.zeros((n, 3), np.float64),
        "scat_next": np.zeros(n, np.float64), "w": np.zeros(n, np.int64), "p": np.zeros(n, np.int64),
        "sign": np.zeros(n, np.int32), "alive": np.zeros(n, np.int32), "sdom": np.zeros(n, np.int32),
        "nscat": np.zeros(n, np.int64), "steps": np.zeros(n, np.int64), "cell": np.zeros((n, 3), np.int32),
    }
    out = TraceOut()
    for name, ctype in TraceOut._fields_:
        setattr(out, name, bufs[name].ctypes.data_as(ctype))
    return bufs, out
