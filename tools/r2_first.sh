#!/bin/bash
# GPU pass: staged smoke (short timeouts), then parity tests with per-test timeouts, a 1000-series bench, an ncu look
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for st in nofast strict fold; do
  timeout 120 python tools/r2_debug.py $st > gpurun_out/r2_dbg_$st.log 2>&1; echo "stage $st rc=$?" | tee -a gpurun_out/r2_dbg_$st.log
done
grep -q "stage fold OK" gpurun_out/r2_dbg_fold.log || { for f in gpurun_out/r2_dbg_*.log; do tail -n 5 $f; done; exit 1; }
timeout 120 python tools/r2_debug.py fold 1000 100000 > gpurun_out/r2_dbg_fold_big.log 2>&1; echo "big rc=$?" >> gpurun_out/r2_dbg_fold_big.log
( time timeout 1500 python -m pytest tests -m gpu -x -q --timeout=240 2>&1 | tail -40 ) > gpurun_out/r2_pytest.log 2>&1
timeout 600 python bench.py --series 1000 --no-e2e --no-cpu --steps 5 > gpurun_out/r2_bench_1000.json 2> gpurun_out/r2_bench_1000.err
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"k_fused_il|k_fix_edges|k_merge" -c 12 --csv --log-file gpurun_out/r2_ncu_1000.csv python bench.py --series 1000 --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_1000.err
for f in gpurun_out/r2_dbg_*.log; do tail -n 3 $f; done; tail -n 8 gpurun_out/r2_pytest.log
head -c 1200 gpurun_out/r2_bench_1000.json
