#!/bin/bash
# round-2 GPU pass B: parity suite on HEAD, A/B of library variants (abv/libmcb_<tag>.so; a tag ending in b<N> runs with N-thread CTAs), ncu counters per variant
# usage: tools/gpu_r2_b.sh "<workloads>" tag1 tag2 ...
mkdir -p gpurun_out
wl="$1"; shift
if [ -z "$SKIP_TESTS" ]; then
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
fi
args=""; for t in "$@"; do args="$args $t=abv/libmcb_$t.so"; done
bash tools/ab_libs.sh "$wl" $args 2>&1 | tee gpurun_out/ab_b.log
for t in "$@"; do
  for w in $wl; do
  AB_BLOCK=768 MCB_LIBMCB="$(pwd)/abv/libmcb_$t.so" timeout 300 ncu --metrics smsp__inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,launch__grid_size,launch__block_size \
    --clock-control none -k regex:k_step -s 20 -c 1 --csv --log-file gpurun_out/inst_${t}_$w.csv python tools/ab_run.py $w > gpurun_out/ncu_inst_$t.log 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/inst_${t}_$w.csv")) if len(r)>14 and r[0].isdigit()]
d={r[12]:r[14] for r in rows}
blk=int(d.get("launch__block_size","0").replace(",","")); grid=int(d.get("launch__grid_size","0").replace(",",""))
inst=float(d["smsp__inst_executed.sum"].replace(",","")); ns=float(d["gpu__time_duration.sum"].replace(",",""))
print("$t $w: block %d grid %d  time %.1f us  inst %.3e  issue %.1f%%  stalls long_sb %s barrier %s wait %s short_sb %s" % (blk, grid, ns/1e3, inst, float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]), d["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"], d["smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]))
PY
  done
done 2>&1 | tee gpurun_out/inst_b.log
