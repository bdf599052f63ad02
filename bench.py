#!/usr/bin/env python
"""bench.py — phonon-steps/s of the FieldProblem::solve hot path on N B200s (one process per GPU).

A "step" of this benchmark is ONE pass of the hot path over one batch of synthetic phonons: a full
`FieldProblem::solve` of BASELINE.json configs[1] ("1D cross-plane Si thin film 100 nm at 300 K, 1e7
phonons", pinned in SURVEY.md §8d as C2: slab between isothermal walls at +-0.5 K, 100 cells, Multi
tally, maxscat 1000, Si-like full dispersion nw=1000 x np=3).  The metric is phonon-steps/s where one
phonon-step is one trip of the loop body problem.cpp:401-435.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mode resident|streaming]

The one JSON line carries the headline (C2) and, under "configs", the other BASELINE configurations that fit one GPU
(C1 film, C3 wire 32x32 with 1e8 phonons, C4 tube with 1e8, the 1.25e8-phonon C5 slice: bulk 128^3) measured in the same
run with a few solves each; at N > 1 the C5 record is the sharded 1e9-phonon problem with its 67-MB all-reduce timed.

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
SLOT_BYTES = 72          # the shipped layout: 7 doubles + meta + pid|step per slot
WORKLOADS = {
    # name: (domain kind, dim, div, dT, problem, maxscat, material, default phonons per GPU)
    "C2-slab100nm-si": ("slab", [100e-9, 100e-9, 100e-9], [100, 0, 0], 1.0, "multi", 1000, "silicon", 10_000_000),
    "C2-slab100nm-grey": ("slab", [100e-9, 100e-9, 100e-9], [100, 0, 0], 1.0, "multi", 1000, "grey", 10_000_000),
    "C1-film100nm-si": ("film", [1e-6, 100e-9, 1e-6], [0, 20, 0], 1.0, "multi", 100, "silicon", 10_000_000),
    "C3-wire32x32-si": ("wire", [1e-6, 100e-9, 100e-9], [0, 32, 32], 1.0, "multi", 100, "silicon", 100_000_000),
    "C4-tube-si": ("tube", [1e-6, 5e-8, 5e-8, 2e-8], [0, 8, 8, 4], 1.0, "multi", 100, "silicon", 100_000_000),
    "C5-bulk128-si": ("bulk", [1e-6, 1e-6, 1e-6], [128, 128, 128], 1.0, "multi", 100, "silicon", 125_000_000),
}
HEADLINE = "C2-slab100nm-si"
SIDE_CONFIGS = ["C1-film100nm-si", "C3-wire32x32-si", "C4-tube-si", "C5-bulk128-si"]      # N = 1: measured beside the headline
# committed ncu --set full captures of the steady-phase k_step launch of each workload: profiles/ncu_traffic.json (keyed by workload,
# written by tools/ncu_traffic.py from the captures of tools/gpu_r2_final.sh)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=HEADLINE, choices=sorted(WORKLOADS))
    ap.add_argument("--nemit", type=int, default=0, help="phonons per GPU per step (0 = the workload's BASELINE size)")
    ap.add_argument("--mode", default="streaming", choices=["resident", "streaming"],
                    help="streaming: one loop trip per state load/store while the population is full (S=1; the mode the "
                         "HBM roofline is quoted on); resident: S=16 loop trips per load/store (max phonon-steps/s)")
    ap.add_argument("--slots", type=int, default=0, help="resident phonon slots per GPU (0 = library default)")
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="phonons in the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-side-configs", action="store_true", help="headline workload only")
    ap.add_argument("--side-steps", type=int, default=3, help="timed solves per side configuration (after one warm-up)")
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


_MATS = {}


def material_files(kind):
    from montecarlocpp_b200 import materials
    if kind not in _MATS:
        d = tempfile.mkdtemp(prefix="mcb_mat_")
        _MATS[kind] = materials.write_grey(d) if kind == "grey" else materials.write_silicon(d, nw=1000)
    return _MATS[kind]


def host_threads():
    """All host cores, whatever OMP_NUM_THREADS says: torchrun exports OMP_NUM_THREADS=1 to its workers, which made the
    round-1 reference arm run on ONE core at N > 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference_rate(workload, sample, seed, threads=0, prefer="reference"):
    """The reference's CPU implementation of the path on a bounded sample (a problem of the same domain / material /
    maxscat with nemit = sample) on the host cores.
      kind "reference": the REFERENCE ITSELF -- oracle/_ref/ref_driver, the reference's own objects (FieldProblem::solve,
                        problem.cpp:370-445, OpenMP as in main.cpp:155) built against the Eigen/Boost stand-ins; its time is
                        the solve alone, its step count is the reference's own loop-trip count;
      kind "port":      the oracle restatement (when the prebuilt reference binary is absent).
    Only this leg (and tests / smoke) may touch oracle/."""
    from oracle import pyoracle as orc
    kind, dim, div, dT, pkind, maxscat, mkind, _ = WORKLOADS[workload]
    disp, relax = material_files(mkind)
    cores = threads or host_threads()
    n = sample
    try:
        from oracle import refbin
        have_ref = refbin.driver_available() and prefer == "reference"
    except Exception:
        have_ref = False
    if have_ref:
        r = refbin.drive(disp, relax, 300.0, kind, list(dim), list(div), dT, pkind, n, maxscat, seed=seed, threads=cores)
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
    _, st = prob.solve(rng=orc.RNG_MT19937, seed=seed, n_begin=0, n_end=prob.nemit, nthreads=cores)
    dt = time.perf_counter() - t0
    return st["steps"] / dt, dt, st["steps"], cores, prob.nemit, "port"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nemit = args.nemit or WORKLOADS[args.workload][7]
    rates, t_all = [], []
    cores, kind, n = None, "port", 0
    sample = min(args.cpu_sample, nemit * args.gpus)
    for i in range(args.warmup + args.steps):
        rate, dt, steps, cores, n, kind = cpu_reference_rate(args.workload, sample, 1000 + i)
        if dt > 15.0:                                    # keep the whole arm within a few minutes on a slow host
            sample = max(10_000, int(sample * 10.0 / dt))
        if i >= args.warmup:
            rates.append(rate); t_all.append(dt)
    total_rate = sum(rates) / len(rates)
    what = ("the reference's own FieldProblem::solve (oracle/_ref/ref_driver: reference objects built against the Eigen/Boost "
            "stand-ins, not real Eigen), mt19937 per thread, OpenMP static") if kind == "reference" else "oracle port, mt19937, OpenMP static"
    sample_s = (f"{n} phonons of {args.workload} per step (the full problem has {nemit * args.gpus}; a rate, so the sample size "
                f"only bounds the run time), {what}, {cores} threads (all host cores; OMP_NUM_THREADS from torchrun is ignored)")
    line = {
        "impl": "reference", "metric": "phonon_steps_per_s", "value": total_rate, "unit": "phonon-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(t_all) / len(t_all),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": args.workload, "nemit_per_gpu": nemit, "sample_phonons_per_step": n,
                   "note": "CPU: the reference's algorithm on the host cores, bounded sample per step"},
        "cpu_baseline": {"value": total_rate, "unit": "phonon-steps/s", "cores": cores, "kind": kind, "sample": sample_s},
        "e2e": {"value": total_rate, "unit": "phonon-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


class Runner:
    """One workload on this rank's GPU: tables resident, raw tally on the device, all-reduce + finalize on the library's stream."""

    def __init__(self, ctx, workload, nemit_per_gpu, world, rank, S, slots):
        import torch
        from montecarlocpp_b200 import hostapi, sharding
        self.torch, self.ctx, self.world, self.rank, self.workload = torch, ctx, world, rank, workload
        kind, dim, div, dT, pkind, maxscat, mkind, _ = WORKLOADS[workload]
        self.maxscat, self.pkind, self.mkind = maxscat, pkind, mkind
        disp, relax = material_files(mkind)
        self.mat = hostapi.Material(disp, relax, 300.0)
        self.dom = hostapi.Domain(kind, dim, div, dT)
        self.nemit_per_gpu = nemit_per_gpu
        self.prob = hostapi.FieldProblem(self.mat, self.dom, pkind, nemit_per_gpu * world, maxscat)   # weak scaling: per-GPU work fixed
        self.n_begin, self.n_end = sharding.shard_range(self.prob.nemit, world, rank)
        ctx.upload_material(self.mat.desc)
        ctx.upload_domain(self.dom.desc)
        ctx.set_options(steps_per_launch=S, slots=slots)
        self.S, self.slots = S, slots
        self.raw = torch.zeros(self.prob.rows * self.dom.cols, dtype=torch.float64, device="cuda")
        # the library's own stream: the zeroing, the all-reduce and the finalize are issued on it, so the NCCL kernel is
        # ordered after the solve's last flush and before k_finalize (a collective on torch's stream was not)
        self.lib_stream = torch.cuda.ExternalStream(ctx.stream())
        self.ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        self.coll_ms = self.zero_ms = self.fin_ms = 0.0
        self.arrive = []
        torch.cuda.synchronize()

    def step(self, seed, timed=True):
        torch, ctx = self.torch, self.ctx
        import torch.distributed as dist
        with torch.cuda.stream(self.lib_stream):
            self.ev[0].record()
            self.raw.zero_()
            self.ev[1].record()
        st = ctx.solve_raw_dev(self.prob.desc, self.raw.data_ptr(), seed=seed, n_begin=self.n_begin, n_end=self.n_end)
        t_arrive = time.perf_counter()
        with torch.cuda.stream(self.lib_stream):
            self.ev[2].record()
            if self.world > 1:
                dist.all_reduce(self.raw, op=dist.ReduceOp.SUM)      # the per-solve field reduction (main.cpp:162-165) over NCCL
            self.ev[3].record()
        ctx.finalize_dev(self.prob.desc, self.raw.data_ptr())          # k_finalize on the same stream, then a stream sync
        with torch.cuda.stream(self.lib_stream):
            self.ev[4].record()
        self.ev[4].synchronize()
        if timed:
            self.fin_ms += self.ev[3].elapsed_time(self.ev[4])
            self.zero_ms += self.ev[0].elapsed_time(self.ev[1])
            self.coll_ms += self.ev[2].elapsed_time(self.ev[3])
            self.arrive.append(t_arrive)
        return st

    def e2e_step(self, seed):
        """The call a user makes, with HOST buffers: tables uploaded, solve, field read back."""
        ctx = self.ctx
        ctx.upload_material(self.mat.desc); ctx.upload_domain(self.dom.desc)
        if self.world == 1:
            sol, st = ctx.solve(self.prob.desc, seed=seed, n_begin=self.n_begin, n_end=self.n_end)      # host field out
            return st
        st = self.step(seed, timed=False)
        self.raw.cpu()                                                                            # field D2H on every rank
        return st


KEYS = ["steps", "launches", "step_ms", "step_launches", "state_stores", "device_ms", "esc",
        "steady_launches", "steady_steps", "steady_stores", "steady_ms"]


def barrier(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def measure(runner, steps, warmup, world, seed0, e2e_steps, peaks, sampler=None):
    """W warm-up solves, then K timed ones between barrier + cuda synchronize; max over ranks; returns the record."""
    import torch
    import torch.distributed as dist
    for i in range(warmup):
        runner.step(seed0 + 100 + i, timed=False)
    if sampler:
        sampler.start()
    barrier(world)
    t0 = time.perf_counter()
    tot = {k: 0 for k in KEYS}
    for i in range(steps):
        st = runner.step(seed0 + 1000 + i)
        for k in KEYS:
            tot[k] += st[k]
        tot["launches"] += 2                            # the zeroing and the finalize kernel
    barrier(world)
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop() if sampler else None
    e2e_elapsed, e2e_n = 0.0, 0
    if e2e_steps > 0:
        runner.e2e_step(seed0 + 7)
        barrier(world)
        t1 = time.perf_counter()
        for i in range(e2e_steps):
            e2e_n += runner.e2e_step(seed0 + 2000 + i)["steps"]
        barrier(world)
        e2e_elapsed = time.perf_counter() - t1
    arrive = runner.arrive[-steps:]
    vals = torch.tensor([elapsed, e2e_elapsed, float(tot["steps"]), float(e2e_n), float(tot["launches"]), float(tot["device_ms"]),
                         runner.coll_ms, runner.fin_ms, runner.zero_ms, float(tot["state_stores"])], dtype=torch.float64, device="cuda")
    skew_ms = 0.0
    if world > 1:
        mx = vals.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = vals.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        elapsed, e2e_elapsed, device_ms = mx[0].item(), mx[1].item(), mx[5].item()
        all_steps, all_e2e, all_launches, all_stores = sm[2].item(), sm[3].item(), sm[4].item(), sm[9].item()
        coll_ms, fin_ms, zero_ms = mx[6].item(), mx[7].item(), mx[8].item()
        at = torch.tensor(arrive, dtype=torch.float64, device="cuda")
        lo = at.clone(); dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        hi = at.clone(); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        skew_ms = 1e3 * float((hi - lo).mean().item())      # same-host ranks: perf_counter is one clock
    else:
        all_steps, all_e2e, all_launches, all_stores = float(tot["steps"]), float(e2e_n), float(tot["launches"]), float(tot["state_stores"])
        device_ms, coll_ms, fin_ms, zero_ms = float(tot["device_ms"]), runner.coll_ms, runner.fin_ms, runner.zero_ms
    peak = float(peaks.get("hbm_gbs", 6650.0))
    sdy_s = tot["steady_ms"] * 1e-3
    achieved = tot["steady_stores"] * B_ALG / sdy_s / 1e9 if sdy_s > 0 else 0.0
    rec = {
        "value": all_steps / elapsed, "unit": "phonon-steps/s", "ms_per_step": 1e3 * elapsed / steps, "steps": steps,
        "nemit_per_gpu": runner.n_end - runner.n_begin, "phonon_steps_per_solve": tot["steps"] / steps, "esc": tot["esc"],
        "device_ms_per_step": device_ms / steps,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "kernel": f"k_step, steady-phase launches (S={runner.S}: population full, one state round trip per loop trip; the "
                               "launch also refills the slots that end inactive -- K1 is fused into it)",
                     "algorithmic_bytes_per_launch": tot["steady_stores"] * B_ALG / max(1, tot["steady_launches"]),
                     "kernel_ms_per_launch": tot["steady_ms"] / max(1, tot["steady_launches"]),
                     "kernel_share_of_step": sdy_s / elapsed,
                     "phonon_steps_per_s_in_kernel": tot["steady_steps"] / sdy_s if sdy_s > 0 else 0.0,
                     "bytes_per_phonon_step": tot["steady_stores"] * B_ALG / max(1, tot["steady_steps"])},
        # the whole timed step, every k_step launch of it (steady + decay): state round trips x B_alg over the wall time
        "whole_step": {"state_round_trips": all_stores, "hbm_gbs": all_stores * B_ALG / elapsed / 1e9,
                       "hbm_frac": all_stores * B_ALG / elapsed / 1e9 / (peak * world),
                       "k_step_share_of_step": tot["step_ms"] * 1e-3 / elapsed,
                       "note": "time-weighted: after the last emission every launch runs several loop trips per state round trip (S from the "
                               "measured termination rate, ~12 % of the live phonons per launch) and stores its survivors densely "
                               "(compaction fused into k_step), so the decay phase moves few bytes per phonon-step by construction"},
        "phases": {"steady": {"launches": tot["steady_launches"], "ms": tot["steady_ms"], "phonon_steps": tot["steady_steps"]},
                   "decay": {"launches": tot["step_launches"] - tot["steady_launches"], "ms": tot["step_ms"] - tot["steady_ms"],
                             "phonon_steps": tot["steps"] - tot["steady_steps"]}},
        "gpu_launches": int(all_launches),
    }
    if e2e_steps > 0:
        rec["e2e_value"] = all_e2e / e2e_elapsed
    if world > 1:
        rec["collective"] = {"allreduce_bytes": runner.raw.numel() * 8, "collective_ms_per_step": coll_ms / steps,
                             "finalize_ms_per_step": fin_ms / steps, "zero_ms_per_step": zero_ms / steps,
                             "rank_skew_ms_per_step": skew_ms,
                             "note": "max over ranks; collective_ms = CUDA events on the library's stream around the NCCL all-reduce "
                                     "(it includes waiting for the slowest rank's solve: rank_skew_ms is the spread of the arrival times)"}
    return rec, clocks


def ncu_capture(workload, mode, slots):
    """DRAM traffic / issue-slot utilisation of the steady k_step launch from the committed ncu --set full capture."""
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))["kernels"]
        k = tj[workload]
        if mode != "streaming" or slots:
            return None, None
        return k.get("dram_bytes_per_slot"), k.get("issue_active_pct")
    except Exception:
        return None, None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from montecarlocpp_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass

    S = 1 if args.mode == "streaming" else 16
    ctx = capi.Context(local)
    nemit = args.nemit or WORKLOADS[args.workload][7]
    head = Runner(ctx, args.workload, nemit, world, rank, S, args.slots)
    sampler = ClockSampler(local) if rank == 0 else None
    rec, clocks = measure(head, args.steps, args.warmup, world, 0, max(1, min(args.steps, 5)), peaks, sampler)

    # ---- the other schedule, for reference: the library's DEFAULT (no options: every phonon resident when the state fits in
    # memory, S = 16 loop trips per state round trip) -- what mcb_solve does for a caller who sets nothing
    other = None
    if args.mode == "streaming":
        ctx.set_options(steps_per_launch=0, slots=0)
        head.S = 16
        head.step(50, timed=False)
        barrier(world)
        t2 = time.perf_counter()
        osteps = 0
        for i in range(3):
            osteps += head.step(3000 + i, timed=False)["steps"]
        barrier(world)
        other = (osteps, time.perf_counter() - t2)
        head.S = S

    # ---- the other BASELINE configurations, a few solves each (N = 1: all that fit one GPU; N > 1: the sharded C5)
    side = {}
    names = [] if args.no_side_configs or args.workload != HEADLINE else (SIDE_CONFIGS if world == 1 else ["C5-bulk128-si"])
    for name in names:
        r = Runner(ctx, name, WORKLOADS[name][7], world, rank, S, 0)
        srec, _ = measure(r, args.side_steps, 1, world, 10_000, 1 if world == 1 else 0, peaks)
        per_slot, issue = ncu_capture(name, args.mode, 0)       # ncu DRAM bytes per visited slot x the slots one bench launch visits
        srec["roofline"]["traffic"] = per_slot * srec["roofline"]["algorithmic_bytes_per_launch"] / B_ALG if per_slot else None
        srec["roofline"]["issue_active_pct_ncu"] = issue
        srec["config"] = {"workload": name, "nemit_per_gpu": WORKLOADS[name][7], "maxscat": WORKLOADS[name][5],
                          "field": f"{r.prob.rows}x{r.dom.cols}"}
        if world > 1:
            # SURVEY 8d C5 stress variant: the tally all-reduced many times per solve instead of once.  The algorithm has no
            # per-step coupling, so the solve is cut into CHUNKS particle ranges per rank and every chunk's raw 67-MB tally is
            # all-reduced on its own (on the library's stream) and added up: what frequent collectives would cost.
            CHUNKS = 16
            tmp = torch.zeros_like(r.raw)
            span = (r.n_end - r.n_begin + CHUNKS - 1) // CHUNKS
            barrier(world)
            t4 = time.perf_counter()
            ssteps = 0
            with torch.cuda.stream(r.lib_stream):
                r.raw.zero_()
            for k in range(CHUNKS):
                with torch.cuda.stream(r.lib_stream):
                    tmp.zero_()
                b0 = r.n_begin + k * span
                ssteps += ctx.solve_raw_dev(r.prob.desc, tmp.data_ptr(), seed=77, n_begin=b0, n_end=min(r.n_end, b0 + span))["steps"]
                with torch.cuda.stream(r.lib_stream):
                    dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
                    r.raw.add_(tmp)
            ctx.finalize_dev(r.prob.desc, r.raw.data_ptr())
            barrier(world)
            dt4 = time.perf_counter() - t4
            tot4 = torch.tensor([float(ssteps)], dtype=torch.float64, device="cuda")
            dist.all_reduce(tot4, op=dist.ReduceOp.SUM)
            srec["collective_stress"] = {"allreduces_per_solve": CHUNKS, "allreduce_bytes": r.raw.numel() * 8, "value": tot4.item() / dt4,
                                         "ms_per_solve": 1e3 * dt4,
                                         "note": "one solve cut into 16 particle chunks per rank, the raw tally all-reduced after every chunk"}
            del tmp
        if args.mode == "streaming":            # the max-throughput schedule beside it (S = 16 loop trips per state round trip)
            ctx.set_options(steps_per_launch=0, slots=0)        # the library's default schedule
            r.S = 16
            r.step(60, timed=False)
            barrier(world)
            t3 = time.perf_counter()
            rsteps = sum(r.step(4000 + i, timed=False)["steps"] for i in range(2))
            barrier(world)
            srec["resident_mode_value_per_gpu"] = rsteps / (time.perf_counter() - t3)
            ctx.set_options(steps_per_launch=S, slots=0)
        side[name] = srec
        del r

    if rank == 0:
        kind, dim, div, dT, pkind, maxscat, mkind, _ = WORKLOADS[args.workload]
        nw, npol = head.mat.desc.nw, head.mat.desc.np
        h2d = nw * npol * (8 * 3 + 1) + nw * 10 + nw * npol * 12 + nw * 12 + 2048     # tables + alias + geometry (approx, bytes)
        d2h = head.prob.rows * head.dom.cols * 8
        slots = args.slots if args.slots else 148 * 768 * 32          # library default for an explicit S: 32 ... 96 tiles per warp, ~ nemit / 3 (768-thread CTAs for 1-D tallies)
        per_slot, issue = ncu_capture(args.workload, args.mode, args.slots)
        roof = rec["roofline"]
        traffic = per_slot * roof["algorithmic_bytes_per_launch"] / B_ALG if per_slot else None
        roof.update({"traffic": traffic, "traffic_unit": "bytes per launch: ncu dram__bytes_read.sum + dram__bytes_write.sum per visited slot of the committed "
                                                         "capture of this kernel (profiles/ncu_traffic.json -- ncu cannot run inside the bench) x the "
                                                         "slots one launch of this run visits",
                     "issue_active_pct_ncu": issue,
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                     "note": f"B_alg = 128 B per state round trip (SURVEY 8d; the shipped layout moves {2 * SLOT_BYTES} B, see profiles/); "
                             "the kernel is issue-bound (fp64 geometry + Philox + tally), not HBM-bound"})
        line = {
            "metric": "phonon_steps_per_s", "value": rec["value"], "unit": "phonon-steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "timing": {"value_from": "host clock around the K steps, barrier + cuda synchronize on both sides, max over ranks",
                       "device_ms_per_step": rec["device_ms_per_step"],
                       "device_note": "CUDA events on the library's stream around each solve (mcb_stats.device_ms), summed over "
                                      "the K steps, max over ranks; excludes the all-reduce and the host gaps between solves"},
            "config": {"workload": args.workload, "nemit_per_gpu": nemit, "maxscat": maxscat, "problem": pkind,
                       "material": f"{mkind} nw={nw} np={npol}", "mode": args.mode, "steps_per_launch": S,
                       "l2": f"resident state {slots} slots x {SLOT_BYTES} B = {slots * SLOT_BYTES / 1e6:.0f} MB > 126 MB L2 (no flush needed)",
                       "parallelism": f"phonons sharded over {world} GPU(s); one fp64 all-reduce of the {head.prob.rows}x{head.dom.cols} tally per solve"},
            "clocks": clocks,
            "e2e": {"value": rec["e2e_value"], "unit": "phonon-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "calls": "mcb_upload_material + mcb_upload_domain + mcb_solve (host buffers)"},
            "gpu_launches": rec["gpu_launches"],
            "roofline": roof, "whole_step": rec["whole_step"], "phases": rec["phases"],
            "phonon_steps_per_solve": rec["phonon_steps_per_solve"], "esc": rec["esc"],
        }
        if "collective" in rec:
            line["collective"] = rec["collective"]
        if side:
            line["configs"] = side
        if other is not None:
            line["resident_mode"] = {"value_per_gpu": other[0] / other[1], "unit": "phonon-steps/s", "steps_per_launch": "2 .. 64, from the measured termination rate",
                                     "slots": "library default: every phonon of the solve resident when the two state buffers fit in half "
                                              "of the free device memory",
                                     "note": "rank-0 rate of the library's default schedule (mcb_options all zero: state kept in registers for "
                                             "several loop trips per HBM round trip, the population only decays and every launch compacts its "
                                             "survivors); not the mode the roofline is quoted on (SURVEY 8d rule iv: B_alg / S by construction)"}
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only
            rate, dt, steps, cores, n, kind_ = cpu_reference_rate(args.workload, min(args.cpu_sample, nemit), 4242)
            line["cpu_baseline"] = {"value": rate, "unit": "phonon-steps/s", "cores": cores, "kind": kind_,
                                    "sample": f"{n} phonons of the same problem (of {nemit}; {steps} phonon-steps, {dt:.1f} s solve), "
                                              + ("the reference's own objects via oracle/_ref/ref_driver (built against the Eigen/Boost "
                                                 "stand-ins of oracle/shim, not real Eigen)" if kind_ == "reference" else "oracle port")
                                              + f", OpenMP, {cores} threads"}
            if kind_ == "reference":                         # the restatement beside it, for scale
                prate = cpu_reference_rate(args.workload, min(args.cpu_sample, nemit), 4242, prefer="port")[0]
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
