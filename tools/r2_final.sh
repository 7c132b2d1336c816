#!/bin/bash
# final measurement pass of round 2 (one B200): tests, bench lines, ncu launch list, ncu --set full of the two dominant kernels
# (the .ncu-rep files of full-size runs exceed what gpurun copies back: they are summarised on the box and removed)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -n 25 ) > $O/r02_pytest_gpu.log 2>&1; tail -n 3 $O/r02_pytest_gpu.log
timeout 900 python bench.py > $O/r02_bench_full.json 2> $O/r02_bench_full.err; tail -n 3 $O/r02_bench_full.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-verify > $O/r02_launches_bench.json 2> $O/r02_launches.err
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fused_il -c 1 -f -o /tmp/r02_il python bench.py --no-e2e --no-cpu --no-verify --steps 1 --warmup 1 > /dev/null 2> $O/r02_ncu_il.err; tail -n 1 $O/r02_ncu_il.err
python tools/ncu_summary.py /tmp/r02_il.ncu-rep 1.2 > $O/r02_ncu_k_fused_il_full.txt 2>&1
python tools/ncu_lines.py /tmp/r02_il.ncu-rep 312500000 1.0 > $O/r02_ncu_k_fused_il_lines.txt 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_cols -c 1 -f -o /tmp/r02_cols python bench.py --workload mixed --no-verify --steps 1 --warmup 1 > /dev/null 2> $O/r02_ncu_cols.err; tail -n 1 $O/r02_ncu_cols.err
python tools/ncu_summary.py /tmp/r02_cols.ncu-rep 1.5 > $O/r02_ncu_k_fused_cols_full.txt 2>&1
python tools/ncu_lines.py /tmp/r02_cols.ncu-rep 31250000 1.5 > $O/r02_ncu_k_fused_cols_lines.txt 2>&1
ls -la /tmp/*.ncu-rep
timeout 300 python bench.py --dist lo --no-e2e --no-cpu > $O/r02_bench_glo.json 2> $O/r02_bench_glo.err
for nulls in 0 50; do timeout 600 python bench.py --workload mixed --nulls $nulls > $O/r02_bench_mixed_$nulls.json 2> $O/r02_bench_mixed_$nulls.err; done
timeout 600 python bench.py --workload downsample > $O/r02_bench_downsample.json 2> $O/r02_bench_downsample.err; tail -n 2 $O/r02_bench_downsample.err; head -c 400 $O/r02_bench_downsample.json; echo
du -sh $O
