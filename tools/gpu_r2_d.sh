#!/bin/bash
# round-2 GPU pass D: parity suite on HEAD, whole-solve A/B of library variants at the bench sizes, optional ncu --set full per workload
# usage: tools/gpu_r2_d.sh "<workloads>" "<ncu workloads>" tag1 tag2 ...      (variants: abv/libmcb_<tag>.so)
mkdir -p gpurun_out
wl="$1"; shift; nwl="$1"; shift
if [ -z "$SKIP_TESTS" ]; then
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -12 gpurun_out/pytest_gpu.log
fi
for round in 1 2; do for t in "$@"; do
  AB_TAG=$t MCB_LIBMCB="$(pwd)/abv/libmcb_$t.so" timeout 600 python tools/ab_whole.py $wl 2>&1 | grep -v "^$"
done; done | tee gpurun_out/ab_d.log
last="${@: -1}"
for w in $nwl; do
  MCB_LIBMCB="$(pwd)/abv/libmcb_$last.so" timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_step -s 20 -c 1 -f -o gpurun_out/prof_r2_${last}_$w \
      python tools/ab_run.py $w > gpurun_out/ncu_full_${last}_$w.log 2>&1
  tail -3 gpurun_out/ncu_full_${last}_$w.log
done
