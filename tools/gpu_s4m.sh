#!/bin/bash
TAG=${1:-s4m}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "compositor or config3" > gpurun_out/pytest_$TAG.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/pytest_$TAG.log | head -20
{
timeout 120 python bench.py --workload cfg3 --steps 20
PE_CFG3_PER_FRAME=1 timeout 120 python bench.py --workload cfg3 --steps 20
} > gpurun_out/bench_$TAG.log 2>&1
grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_$TAG.log
