#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_strings_desc.py -m gpu -q --timeout=120 2>&1 | tail -n 3; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 4
bash tools/r2_sweep.sh 10000 u2 u1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_multi -c 1 -f -o gpurun_out/r2_multi python bench.py --workload mixed --series 10000 --no-verify --steps 1 --warmup 1 > /dev/null 2> gpurun_out/r2_ncu_multi.err
tail -n 2 gpurun_out/r2_ncu_multi.err
