#!/bin/bash
cd "$(dirname "$0")/.."
( timeout 900 python -m pytest tests -m gpu -q --timeout=240 2>&1 | tail -n 5 ) > gpurun_out/r2_pytest.log 2>&1; tail -n 3 gpurun_out/r2_pytest.log
bash tools/r2_sweep.sh 2000 "$@"
unset OGPU_LIB
timeout 900 python bench.py --no-e2e --no-cpu --steps 5 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
python -c "
import json; j=json.load(open('gpurun_out/r2_bench_full.json')); print('FULL ms/step', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'frac', round(j['roofline']['frac'],3), 'value', j['value'])"
