#!/bin/bash
# ncu --set full of one DECAY-phase k_step launch of C2 (library default schedule: launch 13 of the solve, S = 12, ~2.6 M live phonons,
# compacting), the phase that carries ~75 % of C2's time
mkdir -p gpurun_out
sed -i 's/^ctx.solve_raw_dev(prob.desc, raw.data_ptr(), seed=0).*$//' tools/launch_log.py
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 12 -c 1 -f -o gpurun_out/prof_r2_decay_slab \
    python tools/launch_log.py slab 0 > gpurun_out/ncu_decay_slab.log 2>&1
tail -2 gpurun_out/ncu_decay_slab.log
