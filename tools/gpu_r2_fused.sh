#!/bin/bash
# fused compaction: parity suites first, then the decay-schedule A/B (decay_mode 2 = separate compaction passes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_sort.py tests/test_gpu_dropin.py -x -q -m gpu > gpurun_out/pytest_fused.log 2>&1; tail -4 gpurun_out/pytest_fused.log
AB_S=1 timeout 500 python tools/ab_decay.py "${AB_VARIANTS:-2:0 0:12}" slab film wire tube bulk 2>&1 | grep -v "^$" | tee gpurun_out/ab_decay2.log
AB_S=0 timeout 500 python tools/ab_decay.py "${AB_VARIANTS:-2:0 0:12}" slab film wire tube bulk 2>&1 | grep -v "^$" | tee -a gpurun_out/ab_decay2.log
