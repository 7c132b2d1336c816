#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi -L | head -4
( timeout 600 python -m pytest tests/test_gpu_merge.py -m gpu -q --timeout=300 2>&1 | tail -n 8 ) > gpurun_out/r2_pytest_2gpu.log 2>&1; tail -n 4 gpurun_out/r2_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err; tail -n 3 gpurun_out/r02_bench_2gpu.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_bench_2gpu.json'))
print('2gpu value', j['value'], 'ms/step', j['ms_per_step'], 'merge_ms', j['config'].get('merge_ms_per_step'), 'verify', j['verify'])
print('e2e', j['e2e'] and j['e2e']['value'])
PY
