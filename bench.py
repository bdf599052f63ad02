#!/usr/bin/env python
"""bench.py — phonon-steps/s of the FieldProblem::solve hot path on N B200s (one process per GPU).

A "step" of this benchmark is ONE pass of the hot path over one batch of synthetic phonons: a full
`FieldProblem::solve` of BASELINE.json configs[1] ("1D cross-plane Si thin film 100 nm at 300 K, 1e7
phonons", pinned in SURVEY.md §8d as C2: slab between isothermal walls at +-0.5 K, 100 cells, Multi
tally, maxscat 1000, Si-like full dispersion nw=1000 x np=3).  The metric is phonon-steps/s where one
phonon-step is one trip of the loop body problem.cpp:401-435.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode resident|streaming]

Arms
  ours       the CUDA path through the C++ host mirror + C ABI (no oracle on this path).
  reference  the reference's own CPU implementation (oracle/_ref/ref_driver: the reference's objects, compiled from its
             sources against the Eigen/Boost stand-ins under oracle/shim; the oracle port if that binary is absent) on all
             host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_ALG = 128.0            # algorithmic bytes per phonon-step (64 B state read + 64 B write)
WORKLOADS = {
    # name: (domain kind, dim, div, dT, problem, maxscat, material)
    "C2-slab100nm-si": ("slab", [100e-9, 100e-9, 100e-9], [100, 0, 0], 1.0, "multi", 1000, "silicon"),
    "C2-slab100nm-grey": ("slab", [100e-9, 100e-9, 100e-9], [100, 0, 0], 1.0, "multi", 1000, "grey"),
    "C1-film100nm-si": ("film", [1e-6, 100e-9, 1e-6], [0, 20, 0], 1.0, "multi", 100, "silicon"),
    "C3-wire32x32-si": ("wire", [1e-6, 100e-9, 100e-9], [0, 32, 32], 1.0, "multi", 100, "silicon"),
    "C4-tube-si": ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0, "multi", 100, "silicon"),
    "C5-bulk128-si": ("bulk", [1e-6, 1e-6, 1e-6], [128, 128, 128], 1.0, "multi", 100, "silicon"),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2-slab100nm-si", choices=sorted(WORKLOADS))
    ap.add_argument("--nemit", type=int, default=10_000_000, help="phonons per GPU per step")
    ap.add_argument("--mode", default="streaming", choices=["resident", "streaming"],
                    help="streaming: one loop trip per state load/store while the population is full (S=1; the mode the "
                         "HBM roofline is quoted on); resident: S=16 loop trips per load/store (max phonon-steps/s)")
    ap.add_argument("--slots", type=int, default=0, help="resident phonon slots per GPU (0 = library default)")
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="phonons in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def material_files(kind):
    from montecarlocpp_b200 import materials
    d = tempfile.mkdtemp(prefix="mcb_mat_")
    return materials.write_grey(d) if kind == "grey" else materials.write_silicon(d, nw=1000)


def _ref_domain_args(workload):
    """(domain keyword, dim, div, dT) in the form the reference's Domain constructors take (main.cpp:285-362)."""
    kind, dim, div, dT, pkind, maxscat, mkind = WORKLOADS[workload]
    return kind, list(dim), list(div), dT


def cpu_reference_rate(workload, nemit_total, sample, seed, threads=0, prefer="reference"):
    """The reference's CPU implementation of the path on a bounded sample (the first `sample` phonons' worth: a problem of
    the same domain / material / maxscat with nemit = sample) on the host cores.
      kind "reference": the REFERENCE ITSELF -- oracle/_ref/ref_driver, the reference's own objects (FieldProblem::solve,
                        problem.cpp:370-445, OpenMP as in main.cpp:155) built against the Eigen/Boost stand-ins; its time is
                        the solve alone, its step count is the reference's own loop-trip count;
      kind "port":      the oracle restatement (when the prebuilt reference binary is absent).
    Only this leg (and tests / smoke) may touch oracle/."""
    from oracle import pyoracle as orc
    kind, dim, div, dT, pkind, maxscat, mkind = WORKLOADS[workload]
    disp, relax = material_files(mkind)
    cores = orc.max_threads() if threads == 0 else threads
    n = min(sample, nemit_total)
    try:
        from oracle import refbin
        have_ref = refbin.driver_available() and prefer == "reference"
    except Exception:
        have_ref = False
    if have_ref:
        dk, ddim, ddiv, ddT = _ref_domain_args(workload)
        r = refbin.drive(disp, relax, 300.0, dk, ddim, ddiv, ddT, pkind, n, maxscat, seed=seed, threads=cores)
        return r["steps"] / r["seconds"], r["seconds"], r["steps"], r["threads"], n, "reference"
    mat = orc.Material(disp, relax)
    if kind == "slab":
        from montecarlocpp_b200 import abi
        dom = orc.Domain.box([0, 0, 0], dim, div, [0, 0, 0], [abi.BDRY_ISOT, abi.BDRY_SPEC, abi.BDRY_SPEC] * 2,
                             [dT / 2, 0, 0, -dT / 2, 0, 0])
    elif kind == "wire":
        from montecarlocpp_b200 import abi
        dom = orc.Domain.box([0, 0, 0], dim, div, [-dT / dim[0], 0, 0], [abi.BDRY_PERI, abi.BDRY_DIFF, abi.BDRY_DIFF] * 2)
    else:
        dom = orc.Domain.create(kind, dim, div, dT)
    prob = orc.Problem(mat, dom, pkind, n, maxscat)
    t0 = time.perf_counter()
    _, st = prob.solve(rng=orc.RNG_MT19937, seed=seed, n_begin=0, n_end=prob.nemit, nthreads=threads)
    dt = time.perf_counter() - t0
    return st["steps"] / dt, dt, st["steps"], cores, prob.nemit, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rates, t_all = [], []
    cores, kind, n = None, "port", 0
    for i in range(args.warmup + args.steps):
        rate, dt, steps, cores, n, kind = cpu_reference_rate(args.workload, args.nemit * args.gpus, args.cpu_sample, 1000 + i)
        if i >= args.warmup:
            rates.append(rate); t_all.append(dt)
    total_rate = sum(rates) / len(rates)
    what = ("the reference's own FieldProblem::solve (oracle/_ref/ref_driver: reference objects built against the Eigen/Boost "
            "stand-ins), mt19937 per thread, OpenMP static") if kind == "reference" else "oracle port, mt19937, OpenMP static"
    sample = f"{n} phonons of {args.workload} per step (of {args.nemit * args.gpus}), {what}, {cores} threads"
    line = {
        "impl": "reference", "metric": "phonon_steps_per_s", "value": total_rate, "unit": "phonon-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_all) / len(t_all),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nemit_per_gpu": args.nemit,
                   "note": "CPU: the reference's algorithm on the host cores, bounded sample per step"},
        "cpu_baseline": {"value": total_rate, "unit": "phonon-steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": total_rate, "unit": "phonon-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def run_ours(args):
    import torch
    import torch.distributed as dist
    from montecarlocpp_b200 import capi, hostapi, sharding

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    kind, dim, div, dT, pkind, maxscat, mkind = WORKLOADS[args.workload]
    disp, relax = material_files(mkind)
    mat = hostapi.Material(disp, relax, 300.0)
    dom = hostapi.Domain(kind, dim, div, dT)
    n_total = args.nemit * world                       # weak scaling: per-GPU work fixed
    prob = hostapi.FieldProblem(mat, dom, pkind, n_total, maxscat)
    n_begin, n_end = sharding.shard_range(prob.nemit, world, rank)

    ctx = capi.Context(local)
    ctx.upload_material(mat.desc)
    ctx.upload_domain(dom.desc)
    S = 1 if args.mode == "streaming" else 16
    ctx.set_options(steps_per_launch=S, slots=args.slots)
    raw = torch.zeros(prob.rows * dom.cols, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()

    def one_step(seed):
        raw.zero_()
        torch.cuda.synchronize()
        st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=seed, n_begin=n_begin, n_end=n_end)
        sharding.allreduce_raw_field(raw)              # the per-solve field reduction (main.cpp:162-165) over NCCL
        ctx.finalize_dev(prob.desc, raw.data_ptr())
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        one_step(100 + i)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    tot = {"steps": 0, "launches": 0, "step_ms": 0.0, "step_launches": 0, "state_stores": 0, "device_ms": 0.0, "esc": 0,
           "steady_launches": 0, "steady_steps": 0, "steady_stores": 0, "steady_ms": 0.0}
    for i in range(args.steps):
        st = one_step(1000 + i)
        for k in tot:
            tot[k] += st[k]
        tot["launches"] += 1                            # finalize kernel
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end arm: the call a user makes, with HOST buffers: tables uploaded, solve, field read back
    def e2e_step(seed):
        ctx.upload_material(mat.desc); ctx.upload_domain(dom.desc)
        if world == 1:
            sol, st = ctx.solve(prob.desc, seed=seed, n_begin=n_begin, n_end=n_end)      # host field out
            return st
        raw.zero_(); torch.cuda.synchronize()
        st = ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=seed, n_begin=n_begin, n_end=n_end)
        sharding.allreduce_raw_field(raw)
        ctx.finalize_dev(prob.desc, raw.data_ptr())
        raw.cpu()                                                                        # field D2H on every rank
        return st
    e2e_step(7)
    barrier()
    t1 = time.perf_counter()
    e2e_steps = 0
    ke = max(1, min(args.steps, 5))
    for i in range(ke):
        e2e_steps += e2e_step(2000 + i)["steps"]
    barrier()
    e2e_elapsed = time.perf_counter() - t1

    # ---- the other schedule, for reference: resident mode (S = 16 loop trips per state round trip, 8 tiles per thread)
    other = None
    if args.mode == "streaming":
        ctx.set_options(steps_per_launch=16, slots=148 * 768 * 8)
        one_step(50)
        barrier()
        t2 = time.perf_counter()
        osteps = 0
        for i in range(3):
            osteps += one_step(3000 + i)["steps"]
        barrier()
        other = (osteps, time.perf_counter() - t2)
        ctx.set_options(steps_per_launch=S, slots=args.slots)

    vals = torch.tensor([elapsed, e2e_elapsed, float(tot["steps"]), float(e2e_steps), float(tot["launches"]), float(tot["device_ms"])],
                        dtype=torch.float64, device="cuda")
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed, e2e_elapsed, device_ms = mx[0].item(), mx[1].item(), mx[5].item()
        all_steps, all_e2e_steps, all_launches = sm[2].item(), sm[3].item(), sm[4].item()
    else:
        all_steps, all_e2e_steps, all_launches = float(tot["steps"]), float(e2e_steps), float(tot["launches"])
        device_ms = float(tot["device_ms"])

    if rank == 0:
        nw, npol = mat.desc.nw, mat.desc.np
        h2d = nw * npol * (8 * 3 + 1) + nw * 10 + nw * npol * 12 + nw * 12 + 2048     # tables + alias + geometry (approx, bytes)
        d2h = prob.rows * dom.cols * 8
        # Roofline of the dominant kernel, k_step (all other launches are < 1 % of the step).  It is quoted on the
        # STEADY-phase launches (population full, S loop trips per state round trip): algorithmic bytes =
        # B_alg x state round trips, counted on the device; duration = CUDA events on the library's stream.
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        sdy_s = tot["steady_ms"] * 1e-3
        achieved = tot["steady_stores"] * B_ALG / sdy_s / 1e9 if sdy_s > 0 else 0.0
        dec_ms = tot["step_ms"] - tot["steady_ms"]
        dec_steps = tot["steps"] - tot["steady_steps"]
        slots = args.slots if args.slots else 148 * 896 * 32          # library default: 32 tiles per CTA (896 threads for 1-D tallies)
        # DRAM traffic of that kernel from the committed ncu --set full capture (profiles/ncu_traffic.json), per launch
        traffic = None
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
            key = {"C2-slab100nm-si": "k_step_steady_S1_C2", "C1-film100nm-si": "k_step_film_S1_C1"}.get(args.workload)
            if key and args.mode == "streaming" and not args.slots:
                traffic = tj[key]["dram_bytes_per_launch"]
        except Exception:
            pass
        line = {
            "metric": "phonon_steps_per_s", "value": all_steps / elapsed, "unit": "phonon-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * elapsed / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "timing": {"value_from": "host clock around the K steps, barrier + cuda synchronize on both sides, max over ranks",
                       "device_ms_per_step": device_ms / args.steps,
                       "device_value": all_steps / (device_ms * 1e-3) if device_ms > 0 else None,
                       "device_note": "CUDA events on the library's stream around each solve (mcb_stats.device_ms), summed over "
                                      "the K steps, max over ranks; excludes the all-reduce and the host gaps between solves"},
            "config": {"workload": args.workload, "nemit_per_gpu": args.nemit, "maxscat": maxscat, "problem": pkind,
                       "material": f"{mkind} nw={nw} np={npol}", "mode": args.mode, "steps_per_launch": S,
                       "l2": f"resident state {slots} slots x 72 B = {slots * 72 / 1e6:.0f} MB > 126 MB L2 (no flush needed)",
                       "parallelism": f"phonons sharded over {world} GPU(s); one fp64 all-reduce of the {prob.rows}x{dom.cols} tally per solve"},
            "clocks": clocks,
            "e2e": {"value": all_e2e_steps / e2e_elapsed, "unit": "phonon-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "calls": "mcb_upload_material + mcb_upload_domain + mcb_solve (host buffers)"},
            "gpu_launches": int(all_launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_unit": "bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum)",
                         "algorithmic_bytes_per_launch": tot["steady_stores"] * B_ALG / max(1, tot["steady_launches"]),
                         "kernel": f"k_step, steady-phase launches (S={S})",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                         "bytes_per_phonon_step": tot["steady_stores"] * B_ALG / max(1, tot["steady_steps"]),
                         "kernel_ms_per_launch": tot["steady_ms"] / max(1, tot["steady_launches"]),
                         "kernel_share_of_step": sdy_s / elapsed,
                         "phonon_steps_per_s_in_kernel": tot["steady_steps"] / sdy_s if sdy_s > 0 else 0.0,
                         "note": "B_alg = 128 B per state round trip (SURVEY 8d; the shipped layout moves 144 B, see profiles/); "
                                 "the kernel is issue-bound (fp64 geometry + Philox + tally), not HBM-bound"},
            "phases": {"steady": {"launches": tot["steady_launches"], "ms": tot["steady_ms"], "phonon_steps": tot["steady_steps"],
                                  "state_round_trips": tot["steady_stores"]},
                       "decay": {"launches": tot["step_launches"] - tot["steady_launches"], "ms": dec_ms, "phonon_steps": dec_steps,
                                 "note": "after the last emission launches are no longer full: S>=16, compaction, then run to completion"},
                       "k_step_share_of_step": tot["step_ms"] * 1e-3 / elapsed},
            "phonon_steps_per_solve": tot["steps"] / args.steps, "esc": tot["esc"],
        }
        if other is not None:
            line["resident_mode"] = {"value_per_gpu": other[0] / other[1], "unit": "phonon-steps/s", "steps_per_launch": 16,
                                     "slots": 148 * 768 * 8, "note": "rank-0 rate of the max-throughput schedule (state kept in "
                                     "registers for 16 loop trips per HBM round trip); not the mode the roofline is quoted on"}
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only
            rate, dt, steps, cores, n, kind = cpu_reference_rate(args.workload, n_total, args.cpu_sample, 4242)
            line["cpu_baseline"] = {"value": rate, "unit": "phonon-steps/s", "cores": cores, "kind": kind,
                                    "sample": f"{n} phonons of the same problem ({steps} phonon-steps, {dt:.1f} s solve), "
                                              + ("the reference's own objects via oracle/_ref/ref_driver" if kind == "reference"
                                                 else "oracle port") + ", OpenMP"}
            if kind == "reference":                         # the restatement beside it, for scale
                prate, pdt, psteps, pcores, pn, _ = cpu_reference_rate(args.workload, n_total, args.cpu_sample, 4242, prefer="port")
                line["cpu_baseline"]["oracle_port_value"] = prate
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line: anything libraries print there (NCCL's version banner, torchrun notices) is
    # sent to stderr by pointing fd 1 at fd 2 for the duration of the run; the JSON line goes to the real stdout
    sys.stdout.flush()
    _real = os.dup(1)
    os.dup2(2, 1)
    _buf = []
    _print = print

    def print(*args, **kw):                                  # noqa: A001  (the two arms print their line through this)
        _buf.append(" ".join(str(x) for x in args))
    rc = run_reference(a) if a.impl == "reference" else run_ours(a)
    sys.stdout.flush()
    os.dup2(_real, 1)
    for ln in _buf:
        os.write(1, (ln + "\n").encode())
    sys.exit(rc)
