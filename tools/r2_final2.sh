#!/bin/bash
# re-capture of the configs[2] artefacts after the last k_fused_cols change + a last full test run
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
O=gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q --timeout=600 2>&1 | tail -n 25 ) > $O/r02_pytest_gpu.log 2>&1; tail -n 3 $O/r02_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -n 2
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_fused_cols -c 1 -f -o /tmp/r02_cols python bench.py --workload mixed --no-verify --steps 1 --warmup 1 > /dev/null 2> $O/r02_ncu_cols.err; tail -n 1 $O/r02_ncu_cols.err
python tools/ncu_summary.py /tmp/r02_cols.ncu-rep 1.5 > $O/r02_ncu_k_fused_cols_full.txt 2>&1
python tools/ncu_lines.py /tmp/r02_cols.ncu-rep 31250000 1.5 > $O/r02_ncu_k_fused_cols_lines.txt 2>&1
python - <<'PY'
import json, re
t = json.load(open('profiles/traffic.json'))
s = open('gpurun_out/r02_ncu_k_fused_cols_full.txt').read()
def g(name):
    m = re.search(name + r' = ([0-9.]+) (\w+)', s); v = float(m.group(1)); return int(v * {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1}[m.group(2)])
rd, wr = g('dram__bytes_read.sum'), g('dram__bytes_write.sum')
t['k_fused_cols'].update(dram_bytes_per_launch=rd + wr, dram_read=rd, dram_write=wr)
json.dump(t, open('profiles/traffic.json', 'w'), indent=1); open('profiles/traffic.json', 'a').write('\n')
import shutil; shutil.copy('profiles/traffic.json', 'gpurun_out/traffic.json')
print('cols traffic', rd, wr)
PY
for nulls in 0 50; do timeout 600 python bench.py --workload mixed --nulls $nulls > $O/r02_bench_mixed_$nulls.json 2> $O/r02_bench_mixed_$nulls.err; head -c 160 $O/r02_bench_mixed_$nulls.json; echo; done
