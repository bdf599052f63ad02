"""CPU baselines of SURVEY 8(d) on the host cores of the box it runs on: the reference itself (oracle/_ref/ref_driver) and the
oracle port, at 1 thread and at all cores, on C1 (film) and a slice of C2 (slab).  Writes a markdown table to stdout.
TEST INFRASTRUCTURE (uses oracle/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from oracle import pyoracle as orc

rows = []
for wl, sample in (("C1-film100nm-si", 200_000), ("C2-slab100nm-si", 200_000)):
    for thr in (1, orc.max_threads()):
        for prefer in ("reference", "port"):
            rate, dt, steps, cores, n, kind = bench.cpu_reference_rate(wl, 10_000_000, sample, 7, threads=thr, prefer=prefer)
            rows.append((wl, kind, cores, n, steps, dt, rate))
print("| workload | implementation | threads | phonons | phonon-steps | solve s | phonon-steps/s |\n|---|---|---:|---:|---:|---:|---:|")
for r in rows:
    print(f"| {r[0]} | {'the reference itself (oracle/_ref/ref_driver)' if r[1] == 'reference' else 'oracle port'} | {r[2]} | {r[3]} | {r[4]} | {r[5]:.2f} | {r[6]:.3e} |")
