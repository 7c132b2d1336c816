#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout=300 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest.log 2>&1; tail -n 6 gpurun_out/r2_pytest.log
timeout 1200 python bench.py > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.err; tail -n 3 gpurun_out/r2_bench_default.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_default.json'))
print('value', j['value'], 'ms/step', j['ms_per_step'], 'frac', j['roofline']['frac'])
print('e2e', j['e2e']['value'], j['e2e']['phase_ms_per_step'], 'resident', j['e2e']['resident'])
print('cpu', j['cpu_baseline']['value'], j['cpu_baseline']['cores'], j['cpu_baseline']['decoded_MBps_per_thread'])
print('verify', j['verify'])
PY
