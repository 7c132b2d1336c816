#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for v in mb10 mb12; do
  if [ $v = default ]; then unset OGPU_LIB; else export OGPU_LIB=$PWD/opengemini_b200/variants/libogpu_$v.so; fi
  timeout 600 python bench.py --workload mixed --steps 5 > gpurun_out/r2_mixed_$v.json 2> gpurun_out/r2_mixed_$v.err; tail -n 2 gpurun_out/r2_mixed_$v.err
  python - $v <<'PY'
import json,sys
j=json.load(open(f'gpurun_out/r2_mixed_{sys.argv[1]}.json'))
print('mixed', sys.argv[1], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], j['verify'], 'path', j['path'])
PY
done
unset OGPU_LIB
bash tools/r2_sweep.sh 10000 k14 k16
