#!/bin/bash
# round-2 evidence pass on ONE box: parity suite, both bench arms, ncu launch list of the bench command, one ncu --set full
# capture of the steady k_step launch per BASELINE configuration, compute-sanitizer (memcheck + racecheck) on tools/sanitize.py
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-260 gpurun_out/bench_ref.json
timeout 1200 python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc $?"; cut -c1-400 gpurun_out/bench_ours.json; tail -3 gpurun_out/bench_ours.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-side-configs > gpurun_out/ncu_launch.log 2>&1
for w in slab film wire tube bulk; do
  blk=768; case $w in wire|tube|bulk) blk=640;; esac      # the library's default slot count: 148 x CTA threads x 32 tiles (tools/ab_run.py)
  AB_NEMIT=8000000 AB_BLOCK=$blk timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 1 -f -o gpurun_out/prof_r2_final_$w \
      python tools/ab_run.py $w > gpurun_out/ncu_full_final_$w.log 2>&1
  tail -1 gpurun_out/ncu_full_final_$w.log
done
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/memcheck.log 2>&1; tail -2 gpurun_out/memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 2000 python tools/sanitize.py > gpurun_out/racecheck.log 2>&1; tail -2 gpurun_out/racecheck.log
ls -la gpurun_out | tail -15
