#!/bin/bash
# A/B build-time knobs of k_fused_fast on the GPU box: rebuilds libogpu.so per config and times the 1000-series query
for cfg in "-DOG_FAST_MINB=1" "-DOG_FAST_MINB=9" "-DOG_FAST_MINB=10"; do
  (cd opengemini_b200/csrc && touch api.cu && make -s EXTRA="$cfg" >/dev/null 2>&1)
  echo "$cfg"; python tools/prof_query.py 1000 1000000 hi 3 0 | tail -2 | cut -c1-90; python tools/prof_query.py 1000 1000000 lo 3 0 | tail -1
done
(cd opengemini_b200/csrc && touch api.cu && make -s >/dev/null 2>&1)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v9.csv python tools/prof_query.py 1000 1000000 hi 2 0 > gpurun_out/l.log 2>&1
grep "k_merge_all\|k_fix_edges" gpurun_out/launches_v9.csv | tail -2 | cut -d, -f5,12- | cut -c1-120
