#!/bin/bash
TAG=${1:-s4g}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "planar or crossfade or cfg or config or march or golden" > gpurun_out/pytest_$TAG.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/pytest_$TAG.log | head -20
{
timeout 120 python bench.py --workload cfg5 --steps 20
PE_CFG5_PER_CLIP=1 timeout 120 python bench.py --workload cfg5 --steps 20
} > gpurun_out/bench_$TAG.log 2>&1
grep -o '"value": [0-9.]*\|"workload": "[^"]*"\|"frac": [0-9.]*' gpurun_out/bench_$TAG.log
