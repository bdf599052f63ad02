"""Runs the REFERENCE ITSELF: oracle/_ref/montecarlo_ref, the reference's own sources compiled by oracle/Makefile against
the Eigen/Boost stand-ins under oracle/shim/.  TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's cpu_baseline /
--impl reference legs may use it.  Nothing here reads /root/reference at run time (the binary is prebuilt).

The binary keeps the reference's positional grammar (main.cpp:216-235):
    <dir> <material> <T> <domain> <dims...> <divs...> <problem> <nemit> [size] <maxscat> <maxloop> <nsim>
and its stdout blocks ("Output", "Averaged", "Mean", "Standard Deviation").  MCREF_SEED makes the seeds deterministic
(oracle/shim/boost/random/random_device.hpp): thread t of solve k gets seed MCREF_SEED + (calls so far).
"""
import os
import re
import subprocess
import time

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
BIN = os.path.join(_HERE, "_ref", "montecarlo_ref")
BIN_CHK = os.path.join(_HERE, "_ref", "montecarlo_ref_chk")


def available(checked=False):
    return os.access(BIN_CHK if checked else BIN, os.X_OK)


def _blocks(text):
    """{title: 2-D array} for every 'Title\\n<rows of numbers>\\n\\n' block of the reference's stdout."""
    out, lines, i = {}, text.splitlines(), 0
    num = re.compile(r"^\s*[-+]?(\d|nan|inf)", re.I)
    while i < len(lines):
        t = lines[i].strip()
        if t in ("Output", "Averaged", "Mean", "Standard Deviation", "Combined Trajectory"):
            rows, j = [], i + 1
            while j < len(lines) and lines[j].strip() and num.match(lines[j]):
                rows.append([float(x) for x in lines[j].split()])
                j += 1
            out.setdefault(t, []).append(np.array(rows))
            i = j
        else:
            i += 1
    return out


def run(matdir, material, T, domain_args, problem_args, seed=0, threads=1, checked=False, timeout=600):
    """One run of the reference binary.  Returns (blocks, stdout, seconds)."""
    exe = BIN_CHK if checked else BIN
    if not os.access(exe, os.X_OK):
        raise RuntimeError(f"{exe} is missing: run `make -C oracle ref` where /root/reference exists")
    env = dict(os.environ, OMP_NUM_THREADS=str(threads))
    if seed is not None:
        env["MCREF_SEED"] = str(seed)
    else:
        env.pop("MCREF_SEED", None)
    argv = [exe, matdir, material, repr(float(T))] + [str(a) for a in domain_args] + [str(a) for a in problem_args]
    t0 = time.perf_counter()
    r = subprocess.run(argv, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"reference binary failed ({r.returncode}): {r.stderr[-400:]}")
    return _blocks(r.stdout), r.stdout, dt


DRIVER = os.path.join(_HERE, "_ref", "ref_driver")


def driver_available():
    return os.access(DRIVER, os.X_OK)


def drive(disp, relax, T, domain, dim, div, dT, problem, nemit, maxscat, maxloop=0, size=0, seed=0, threads=1, timeout=3600):
    """FieldProblem::solve of the reference's own objects through oracle/ref_driver.cpp (domains: bulk film jct tee tube octet
    + slab / wire composed from the reference's templates).  Thread t uses mt19937(seed + t).
    Returns dict(output (rows x cols), steps, seconds (solve only), esc, threads)."""
    if not driver_available():
        raise RuntimeError(f"{DRIVER} is missing: run `make -C oracle ref` where /root/reference exists")
    matdir = os.path.dirname(os.path.abspath(disp))
    argv = [DRIVER, matdir, os.path.basename(disp), os.path.basename(relax), repr(float(T)), domain, str(len(dim))]
    argv += [repr(float(x)) for x in dim] + [str(len(div))] + [str(int(x)) for x in div] + [repr(float(dT))]
    argv += [problem, str(int(nemit)), str(int(size)), str(int(maxscat)), str(int(maxloop)), str(int(seed))]
    env = dict(os.environ)
    if threads:
        env["OMP_NUM_THREADS"] = str(threads)
    r = subprocess.run(argv, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=timeout)
    if r.returncode != 0:
        raise RuntimeError(f"ref_driver failed ({r.returncode}): {r.stderr[-400:]}")
    lines = r.stdout.splitlines()
    head = {}
    i = 0
    while lines[i].strip() != "Output":
        k, v = lines[i].split()
        head[k] = float(v) if k == "seconds" else int(v)
        i += 1
    rows = [[float(x) for x in ln.split()] for ln in lines[i + 1:i + 1 + head["rows"]]]
    out = np.array(rows).reshape(head["rows"], head["cols"])
    return dict(output=out, steps=head["steps"], seconds=head["seconds"], esc=head["esc"], threads=head["threads"])


class Flattened:
    """C-ABI descriptors (include/mcb.h) of the REFERENCE'S OWN Domain / FieldProblem objects, written by
    `ref_driver ... flatten <file>` (oracle/ref_driver.cpp: the reference-side half of the drop-in binding)."""

    def __init__(self, path):
        import ctypes as C
        from montecarlocpp_b200 import abi
        raw = open(path, "rb").read()
        hdr = np.frombuffer(raw, np.int32, 8)
        assert hdr[0] == 0x4642434D, "not a flatten file"
        nsdom, nplane, npair, nemit, ncols, ssz, psz = (int(x) for x in hdr[1:8])
        assert ssz == C.sizeof(abi.SdomDesc) and psz == C.sizeof(abi.PlaneDesc), "descriptor layout mismatch with include/mcb.h"
        off = 32

        def take(ctype, n):
            nonlocal off
            arr = (ctype * max(n, 1)).from_buffer_copy(raw[off:off + C.sizeof(ctype) * n].ljust(C.sizeof(ctype) * max(n, 1), b"\0"))
            off += C.sizeof(ctype) * n
            return arr
        self.sdoms, self.planes = take(abi.SdomDesc, nsdom), take(abi.PlaneDesc, nplane)
        self.pairs, self.emitters = take(C.c_int32, npair), take(abi.EmitterDesc, nemit)
        self.cell_vol = take(C.c_double, ncols)
        self.problem = abi.ProblemDesc.from_buffer_copy(raw[off:off + C.sizeof(abi.ProblemDesc)]); off += C.sizeof(abi.ProblemDesc)
        self.emit_count = take(C.c_int64, nemit)
        self.problem.emit_count = C.cast(self.emit_count, abi.c_int64_p)
        nweights = int(np.frombuffer(raw[off:off + 4], np.int32)[0]); off += 4
        self.weights = np.frombuffer(raw[off:off + 8 * nweights], np.float64).copy() if nweights else None
        d = abi.DomainDesc()
        d.nsdom, d.sdoms = nsdom, C.cast(self.sdoms, C.POINTER(abi.SdomDesc))
        d.nplane, d.planes = nplane, C.cast(self.planes, C.POINTER(abi.PlaneDesc))
        d.npair, d.pairs = npair, C.cast(self.pairs, abi.c_int32_p)
        d.nemitter, d.emitters = nemit, C.cast(self.emitters, C.POINTER(abi.EmitterDesc))
        d.ncols, d.cell_vol = ncols, C.cast(self.cell_vol, abi.c_double_p)
        self.domain = d
        self.cols = ncols
        self.nsdom, self.nplane, self.npair, self.nemitter = nsdom, nplane, npair, nemit


    def average(self, sol):
        """Domain::average of a rows x cols solution: OctetDomain's weighted mean over its averaging subdomains
        (domain.cpp:1252-1257, rows x 1), the identity for every other domain (domain.cpp:78-81)."""
        if self.weights is None:
            return sol
        return (sol * self.weights[None, :]).sum(axis=1, keepdims=True) / self.weights.sum()


def flatten(disp, relax, T, domain, dim, div, dT, problem, nemit, maxscat, maxloop=0, size=0, outdir=None):
    """Descriptors of the reference's own objects for (domain, problem): see Flattened."""
    import tempfile
    if not driver_available():
        raise RuntimeError(f"{DRIVER} is missing: run `make -C oracle ref` where /root/reference exists")
    matdir = os.path.dirname(os.path.abspath(disp))
    path = os.path.join(outdir or tempfile.mkdtemp(prefix="mcflat_"), f"{domain}.mcbf")
    argv = [DRIVER, matdir, os.path.basename(disp), os.path.basename(relax), repr(float(T)), domain, str(len(dim))]
    argv += [repr(float(x)) for x in dim] + [str(len(div))] + [str(int(x)) for x in div] + [repr(float(dT))]
    argv += [problem, str(int(nemit)), str(int(size)), str(int(maxscat)), str(int(maxloop)), "0", "flatten", path]
    r = subprocess.run(argv, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    if r.returncode != 0:
        raise RuntimeError(f"ref_driver flatten failed ({r.returncode}): {r.stderr[-400:]}")
    return Flattened(path)
