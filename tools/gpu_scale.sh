#!/bin/bash
# Weak-scaling bench at N GPUs of one box (the driver's launch line): tools/gpu_scale.sh N
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
tail -c 400 gpurun_out/bench_${N}gpu.err; cut -c1-260 gpurun_out/bench_${N}gpu.json
