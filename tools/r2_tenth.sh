#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout=300 -k "column_at_a_time or where" 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest10.log 2>&1; tail -n 4 gpurun_out/r2_pytest10.log
for v in default mb4 mb6; do
  if [ $v = default ]; then unset OGPU_LIB; else export OGPU_LIB=$PWD/opengemini_b200/variants/libogpu_$v.so; fi
  timeout 600 python bench.py --workload mixed --steps 5 > gpurun_out/r2_mixed_$v.json 2> gpurun_out/r2_mixed_$v.err; tail -n 2 gpurun_out/r2_mixed_$v.err
  python - $v <<'PY'
import json,sys
j=json.load(open(f'gpurun_out/r2_mixed_{sys.argv[1]}.json'))
print('mixed', sys.argv[1], 'value', j['value'], 'ms/step', j['ms_per_step'], 'kernel_ms', j['roofline']['kernel_ms'], j['verify'], 'path', j['path'])
PY
done
unset OGPU_LIB
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_fused_cols" -c 1 --csv --log-file gpurun_out/r2_ncu_cols.csv python bench.py --workload mixed --steps 1 --warmup 1 --no-verify > /dev/null 2> gpurun_out/r2_ncu_cols.err
grep -v "^==" gpurun_out/r2_ncu_cols.csv | cut -d, -f13-15 | tail -n 9
