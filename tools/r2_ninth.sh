#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strings_desc.py tests/test_tssp_container.py -m gpu -q --timeout=300 -k "column_at_a_time or where or strings or descending or parsed or count_on" 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest9.log 2>&1; tail -n 8 gpurun_out/r2_pytest9.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_cols -c 1 -f -o gpurun_out/r2_cols python bench.py --workload mixed --series 10000 --no-verify --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_cols_full.err
tail -n 2 gpurun_out/r2_ncu_cols_full.err
