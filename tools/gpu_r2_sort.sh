#!/bin/bash
# K3 sort: smoke + the sort parity tests, then whole solves with sort_mode 0 / 1 / 4 (library default schedule and streaming)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_gpu_sort.py -x -q > gpurun_out/pytest_sort.log 2>&1; tail -5 gpurun_out/pytest_sort.log
for s in 0 1 4; do
  for sched in "0 0" "1 0"; do
    set -- $sched
    AB_TAG="sort$s-S$1" AB_SORT=$s AB_S=$1 AB_TILES=$2 timeout 400 python tools/ab_whole.py bulk wire tube slab 2>&1 | grep -v "^$"
  done
done | tee gpurun_out/ab_sort.log
