#!/bin/bash
# ncu --set full of one steady k_step launch (S=1) of the workload(s) given: tools/gpu_r2_prof.sh slab [film ...]
mkdir -p gpurun_out
for wl in "$@"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 1 -f -o gpurun_out/prof_r2_$wl \
    python tools/ab_run.py $wl > gpurun_out/ncu_full_$wl.log 2>&1
tail -2 gpurun_out/ncu_full_$wl.log
done
ls -la gpurun_out/*.ncu-rep
