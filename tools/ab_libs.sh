#!/bin/bash
# A/B prebuilt library variants on ONE box, interleaved: tools/ab_libs.sh "workloads" tag1=lib1.so tag2=lib2.so ...
wl="$1"; shift
for round in 1 2; do
  for kv in "$@"; do
    AB_TAG="${kv%%=*}" AB_BLOCK=768 MCB_LIBMCB="$(pwd)/${kv#*=}" timeout 300 python tools/ab_run.py $wl 2>&1 | grep -v "^$"
  done
done
