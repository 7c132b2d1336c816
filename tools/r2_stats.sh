#!/bin/bash
cd "$(dirname "$0")/.."
OGPU_IL_STATS=1 timeout 300 python bench.py --series 1000 --no-e2e --no-cpu --steps 1 --warmup 3 2>&1 >/dev/null | tail -n 3
