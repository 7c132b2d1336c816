#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fused_il -c 1 -f -o gpurun_out/r2_fused_1000 python bench.py --series ${1:-1000} --no-e2e --no-cpu --no-verify --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_full.err
timeout 300 python bench.py --series 2000 --dist lo --no-e2e --no-cpu --steps 5 > gpurun_out/r2_bench_lo_2000.json 2> gpurun_out/r2_bench_lo_2000.err
python -c "
import json; j=json.load(open('gpurun_out/r2_bench_lo_2000.json')); print('G-lo 2000 ms/step', round(j['ms_per_step'],3), 'kernel_ms', round(j['roofline']['kernel_ms'],3), 'frac', round(j['roofline']['frac'],3), 'value', j['value'], j['config']['compressed_bytes_per_value'])"
