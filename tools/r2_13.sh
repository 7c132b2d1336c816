#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_scale.py -m gpu -q -x --timeout=300 2>&1 | tail -n 15 ) > gpurun_out/r2_pytest13.log 2>&1; tail -n 4 gpurun_out/r2_pytest13.log
bash tools/r2_sweep.sh 10000 k10 k8 u3 u6
