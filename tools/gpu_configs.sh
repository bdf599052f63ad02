#!/bin/bash
# Full-size bench lines of BASELINE.json configs C1, C3, C4 on one GPU (C2 is the default bench; C5 runs on 8 GPUs: tools/gpu_scale.sh)
mkdir -p gpurun_out
python bench.py --workload C1-film100nm-si --nemit 10000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c1.json 2> gpurun_out/bench_c1.err; cut -c1-200 gpurun_out/bench_c1.json
python bench.py --workload C3-wire32x32-si --nemit 100000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; cut -c1-200 gpurun_out/bench_c3.json
python bench.py --workload C4-tube-si --nemit 100000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; cut -c1-200 gpurun_out/bench_c4.json
python bench.py --workload C5-bulk128-si --nemit 125000000 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5_1gpu.json 2> gpurun_out/bench_c5_1gpu.err; cut -c1-200 gpurun_out/bench_c5_1gpu.json
