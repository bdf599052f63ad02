#!/bin/bash
# round-2 GPU pass A: parity suite, A/B of the round-1 library vs HEAD, instruction count of one steady k_step launch
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
bash tools/ab_libs.sh "slab film" r1=abv/libmcb_r1.so head=montecarlocpp_b200/libmcb.so 2>&1 | tee gpurun_out/ab_a.log
timeout 300 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum \
    --clock-control none -k regex:k_step -s 20 -c 1 --csv --log-file gpurun_out/inst_head.csv \
    python tools/ab_run.py slab > gpurun_out/ncu_inst.log 2>&1
grep -E "k_step|inst_executed|duration|issue_active|dram" gpurun_out/inst_head.csv | cut -c1-300 | tail -8
