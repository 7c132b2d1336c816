#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_fused_il -c 1 -f -o gpurun_out/r2_fused_1000 python bench.py --series ${1:-1000} --no-e2e --no-cpu --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_full.err
ls -la gpurun_out/r2_fused_1000.ncu-rep
