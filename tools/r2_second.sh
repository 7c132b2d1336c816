#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q --timeout=240 2>&1 | tail -60 ) > gpurun_out/r2_pytest.log 2>&1
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fused_il -c 1 -f -o gpurun_out/r2_fused_1000 python bench.py --series 1000 --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_full.err
tail -n 30 gpurun_out/r2_pytest.log
