"""GPU tests of the drop-in boundary as the REFERENCE calls it:
  * FieldProblem::solve entered from every thread of an OpenMP region (main.cpp:155-166) sums to ONE solve, for the host
    mirror (libmcbhost.so) and for the reference's own main() linked against the CUDA path (oracle/_ref/montecarlo_gpu:
    the reference's untouched objects + oracle/ref_gpu_solve.cpp);
  * that binary agrees with the reference's CPU binary statistically (batch means, 3 sigma);
  * the C++ multi-GPU path (one context per device, NCCL all-reduce of the raw tallies inside libmcb.so) equals the 1-GPU field.
"""
import os
import subprocess

import numpy as np
import pytest

from montecarlocpp_b200 import hostapi
from oracle import refbin

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
GPU_BIN = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "montecarlo_gpu")


@pytest.fixture(scope="module")
def film(matfiles):
    mat = hostapi.Material(*matfiles["silicon"])
    dom = hostapi.Domain("film", [1e-6, 100e-9, 1e-6], [0, 20, 0], 1.0)
    return mat, dom, hostapi.FieldProblem(mat, dom, "multi", 300_000, 100)


@pytest.mark.parametrize("threads", [1, 3, 16])
def test_mirror_solve_called_like_the_reference_is_one_solve(film, threads):
    """T threads enter FieldProblem::solve; the summed partials equal a single-threaded solve (no T-fold over-count)."""
    mat, dom, prob = film
    hostapi.set_devices([0])
    ref, rst = prob.solve(mt_seed=11)
    got, gst, count = prob.solve_omp(mt_seed=11, nthreads=threads)
    assert gst["steps"] == rst["steps"] and gst["emitted"] == rst["emitted"] == prob.nemit and gst["esc"] == rst["esc"] == 0
    assert count == prob.nemit                                   # Progress advanced once per particle, not T times
    scale = np.abs(ref).max(axis=1, keepdims=True)
    assert (np.abs(got - ref) <= 1e-9 * scale).all()


def _run_gpu_binary(matdir, args, threads, mcb_seed, mcref_seed=5):
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), MCREF_SEED=str(mcref_seed), MCB_SEED=str(mcb_seed), MCB_VERBOSE="1")
    r = subprocess.run([GPU_BIN, matdir] + [str(a) for a in args], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-600:]
    stats = [ln for ln in r.stderr.splitlines() if ln.startswith("mcb: ")]
    return refbin._blocks(r.stdout), r.stdout, stats


@pytest.mark.skipif(not os.access(GPU_BIN, os.X_OK), reason="oracle/_ref/montecarlo_gpu not built (needs /root/reference at build time)")
def test_reference_main_linked_against_the_cuda_path(matfiles, film):
    """The reference's own main() / solveField / printSolution with FieldProblem::solve supplied by the binding TU:
    same field and loop-trip count for 1, 3 and 16 OpenMP threads, and equal to the host mirror's solve with that seed."""
    matdir = os.path.dirname(matfiles["silicon"][0])
    args = ["silicon", 300, "film", 1e-6, 1e-7, 20, "multi", 300000, 100, 0, 1]
    outs = {}
    for t in (1, 3, 16):
        blocks, text, stats = _run_gpu_binary(matdir, args, t, mcb_seed=777)
        assert len(stats) == 1, stats                            # ONE device solve per solveField, whatever the thread count
        assert "esc: 0" in text and f"seeds: " in text
        outs[t] = (np.array(blocks["Output"][0]), stats[0])
    steps = {t: int(s.split("steps ")[1].split()[0]) for t, (_, s) in outs.items()}
    assert steps[1] == steps[3] == steps[16]
    scale = np.abs(outs[1][0]).max(axis=1, keepdims=True)
    for t in (3, 16):
        assert (np.abs(outs[t][0] - outs[1][0]) <= 2e-9 * scale).all()       # printed with 10 digits
    mat, dom, prob = film
    hostapi.set_devices([0])
    ours, st = prob.solve_seeded(777)
    assert st["steps"] == steps[1]
    assert (np.abs(outs[1][0] - ours) <= 2e-9 * scale).all()


@pytest.mark.skipif(not (os.access(GPU_BIN, os.X_OK) and refbin.available()), reason="reference binaries not built")
def test_gpu_binary_agrees_with_the_cpu_reference_statistically(matfiles):
    """montecarlo_gpu vs montecarlo_ref, same argv, 8 repetitions each (the reference's own nsim statistics):
    per-cell |mean difference| <= 4 sigma of the batch means, domain-mean flux within 3 sigma."""
    matdir = os.path.dirname(matfiles["silicon"][0])
    args = ["silicon", 300, "film", 1e-6, 1e-7, 20, "multi", 200000, 100, 0, 8]
    env = dict(os.environ, OMP_NUM_THREADS="4", MCREF_SEED="21")      # no MCB_SEED: every repetition takes its seed from thread 0's engine
    env.pop("MCB_SEED", None)
    r = subprocess.run([GPU_BIN, matdir] + [str(a) for a in args], env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-600:]
    gb = refbin._blocks(r.stdout)
    rb, _, _ = refbin.run(matdir, "silicon", 300, ["film", 1e-6, 1e-7, 20], ["multi", 200000, 100, 0, 8], seed=99, threads=os.cpu_count() or 4)
    gm, gs = np.array(gb["Mean"][0]), np.array(gb["Standard Deviation"][0])
    rm, rs = np.array(rb["Mean"][0]), np.array(rb["Standard Deviation"][0])
    B = 8
    for row in (0, 1):                                           # temperature and q_x
        sig = np.sqrt(gs[row] ** 2 / B + rs[row] ** 2 / B)
        assert (np.abs(gm[row] - rm[row]) <= 4.0 * sig + 1e-30).all(), (row, np.abs(gm[row] - rm[row]) / sig)
    q_g, q_r = gm[1].mean(), rm[1].mean()
    sq = np.sqrt((gs[1] ** 2).sum() / B + (rs[1] ** 2).sum() / B) / gm.shape[1]
    assert abs(q_g - q_r) <= 3.0 * sq, (q_g, q_r, sq)


def test_cpp_multi_gpu_solve_equals_one_gpu(film):
    """FieldProblem::solve sharded over two devices (one mcb context + host thread per device, ncclAllReduce of the raw tallies
    inside libmcb.so, one finalize) against the one-device solve of the same seed."""
    if hostapi.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    mat, dom, prob = film
    hostapi.set_devices([0])
    one, st1 = prob.solve_seeded(4242)
    assert hostapi.set_devices([0, 1]) == 2
    try:
        two, st2 = prob.solve_seeded(4242)
    finally:
        hostapi.set_devices([0])
    assert st2["steps"] == st1["steps"] and st2["emitted"] == st1["emitted"] and st2["esc"] == st1["esc"]
    scale = np.abs(one).max(axis=1, keepdims=True)
    assert (np.abs(two - one) <= 1e-9 * scale).all()
