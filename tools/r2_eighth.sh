#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strings_desc.py tests/test_tssp_container.py tests/test_golden.py -m gpu -q -x --timeout=300 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest8.log 2>&1; tail -n 8 gpurun_out/r2_pytest8.log
for nulls in 0 50; do
timeout 600 python bench.py --workload mixed --nulls $nulls --steps 5 > gpurun_out/r2_bench_mixed_$nulls.json 2> gpurun_out/r2_bench_mixed_$nulls.err; tail -n 2 gpurun_out/r2_bench_mixed_$nulls.err
python - $nulls <<'PY'
import json,sys
j=json.load(open(f'gpurun_out/r2_bench_mixed_{sys.argv[1]}.json'))
print('mixed nulls', sys.argv[1], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], 'frac', j['roofline']['frac'], 'B/row', j['config']['compressed_bytes_per_row'], j['verify'], 'path', j['path'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_fused_cols" -c 2 --csv --log-file gpurun_out/r2_ncu_cols.csv python bench.py --workload mixed --steps 1 --warmup 1 --no-verify > /dev/null 2> gpurun_out/r2_ncu_cols.err
grep -v "^==" gpurun_out/r2_ncu_cols.csv | cut -d, -f5,13-15 | tail -n 9
