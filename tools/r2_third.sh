#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=240 2>&1 | tail -n 80 ) > gpurun_out/r2_pytest.log 2>&1
timeout 600 python bench.py --series 1000 --no-e2e --no-cpu --steps 5 > gpurun_out/r2_bench_1000.json 2> gpurun_out/r2_bench_1000.err
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_fused_il|k_fix_edges|k_merge" -c 5 --csv --log-file gpurun_out/r2_ncu_1000.csv python bench.py --series 1000 --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_1000.err
timeout 900 python bench.py --no-e2e --no-cpu --steps 5 > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err
tail -n 12 gpurun_out/r2_pytest.log
python - <<'PY'
import json
for f in ("gpurun_out/r2_bench_1000.json","gpurun_out/r2_bench_full.json"):
    try:
        j=json.load(open(f)); print(f, "ms/step", round(j["ms_per_step"],3), "kernel_ms", round(j["roofline"]["kernel_ms"],3), "frac", round(j["roofline"]["frac"],3), "value", j["value"])
    except Exception as e: print(f, "ERR", e)
PY
