#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_downsample.py tests/test_gpu_cursor_cpp.py -m gpu -q --timeout=300 2>&1 | tail -n 12 ) > gpurun_out/r2_pytest.log 2>&1; tail -n 5 gpurun_out/r2_pytest.log
for nulls in 0 50; do
timeout 600 python bench.py --workload mixed --nulls $nulls --steps 5 > gpurun_out/r2_bench_mixed_$nulls.json 2> gpurun_out/r2_bench_mixed_$nulls.err; tail -n 2 gpurun_out/r2_bench_mixed_$nulls.err
python - $nulls <<'PY'
import json,sys
j=json.load(open(f'gpurun_out/r2_bench_mixed_{sys.argv[1]}.json'))
print('mixed nulls', sys.argv[1], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], 'frac', j['roofline']['frac'], 'B/row', j['config']['compressed_bytes_per_row'], j['verify'], 'path', j['path'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct --clock-control none -k regex:"k_fused_multi" -c 2 --csv --log-file gpurun_out/r2_ncu_mixed.csv python bench.py --workload mixed --steps 1 --warmup 1 --no-verify > /dev/null 2> gpurun_out/r2_ncu_mixed.err
grep -v "^==" gpurun_out/r2_ncu_mixed.csv | cut -d, -f5,13-15 | tail -n 16
