#!/bin/bash
# A/B for the N-D tally kernels (C3 wire 32x32, C5 bulk 128^3): tools/ab_nd.sh "<EXTRA flags>"
cd montecarlocpp_b200/csrc && make clean >/dev/null && make EXTRA="$1" 2>&1 | grep -E "rror" || true
grep -E "k_stepILi4ELi2ELb1ELb0" -A2 ptxas.log | grep -E "Used|spill" | tr '\n' ' '; echo
cd ../..
for w in C3-wire32x32-si C5-bulk128-si; do for m in streaming resident; do python bench.py --workload $w --steps 3 --warmup 2 --mode $m --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('  ', d['config']['workload'], d['config']['mode'], 'value=%.3e'%d['value'], 'ms=%.1f'%d['ms_per_step'], 'share=%.2f'%r['kernel_share_of_step'])"; done; done
