#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), optionally the ncu launch list + one ncu --set full capture (NCU=1).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; cut -c1-300 gpurun_out/bench_ref.json
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; cut -c1-200 gpurun_out/bench_ours.json
if [ -n "$NCU" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 1 -f -o gpurun_out/prof_steady \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
fi
ls -la gpurun_out | head -30
