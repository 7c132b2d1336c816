#!/bin/bash
# A/B the window geometry of k_fused_fast on the GPU box: rebuilds libogpu.so per config and times the 1000-series query
for cfg in "32 4 4" "64 8 8" "64 4 8" "128 16 8"; do
  set -- $cfg
  (cd opengemini_b200/csrc && touch api.cu && make -s EXTRA="-DOG_IL_NW=${1}u -DOG_IL_K=${2}u -DOG_IL_BATCH=${3}u" >/dev/null 2>&1)
  echo "NW=$1 K=$2 BATCH=$3"; python tools/prof_query.py 1000 1000000 hi 3 0 | tail -2 | cut -c1-90; python tools/prof_query.py 1000 1000000 lo 3 0 | tail -1
done
(cd opengemini_b200/csrc && touch api.cu && make -s >/dev/null 2>&1)
