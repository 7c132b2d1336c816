#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strings_desc.py tests/test_tssp_container.py -m gpu -q --timeout=300 -k "column_at_a_time or where or strings or count_on or parsed or group_modes or nan" 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest14.log 2>&1; tail -n 4 gpurun_out/r2_pytest14.log
for v in default mb8 mb12; do
  if [ $v = default ]; then unset OGPU_LIB; else export OGPU_LIB=$PWD/opengemini_b200/variants/libogpu_$v.so; fi
  timeout 600 python bench.py --workload mixed --steps 5 > gpurun_out/r2_mixed_$v.json 2> gpurun_out/r2_mixed_$v.err; tail -n 2 gpurun_out/r2_mixed_$v.err
  python - $v <<'PY'
import json,sys
j=json.load(open(f'gpurun_out/r2_mixed_{sys.argv[1]}.json'))
print('mixed', sys.argv[1], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], j['verify'], 'path', j['path'])
PY
done
unset OGPU_LIB
timeout 600 python bench.py --workload mixed --nulls 50 --steps 5 2> /dev/null | python -c "import json,sys; j=json.load(sys.stdin); print('nulls50', j['value'], j['verify'])"
