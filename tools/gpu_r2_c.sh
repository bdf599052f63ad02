#!/bin/bash
# round-2 GPU pass C: the whole -m gpu suite (incl. the drop-in tests), then a short bench of both arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "bench rc $?"; cut -c1-1500 gpurun_out/bench_ours.json; tail -5 gpurun_out/bench_ours.err
