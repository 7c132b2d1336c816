#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
N=${1:-1000}; shift
for v in default "$@"; do
  if [ $v = default ]; then unset OGPU_LIB; else export OGPU_LIB=$PWD/opengemini_b200/variants/libogpu_$v.so; fi
  timeout 300 python bench.py --series $N --no-e2e --no-cpu --steps 5 > gpurun_out/sweep_$v.json 2> gpurun_out/sweep_$v.err
  python - "$v" <<'PY'
import json,sys
v=sys.argv[1]
try:
    j=json.load(open(f"gpurun_out/sweep_{v}.json")); print(v, "ms/step", round(j["ms_per_step"],3), "kernel_ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3))
except Exception as e: print(v, "ERR", e, open(f"gpurun_out/sweep_{v}.err").read()[-300:])
PY
done
