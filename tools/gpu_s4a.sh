#!/bin/bash
# session-4 iteration: crossfade batch parity + cfg5 per-clip vs batch, cfg2 with / without 128-bit staging
TAG=${1:-s4a}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "crossfade or resize or cfg or config" > gpurun_out/pytest_$TAG.log 2>&1; tail -3 gpurun_out/pytest_$TAG.log
{
PE_CFG5_PER_CLIP=1 timeout 120 python bench.py --workload cfg5 --steps 20
timeout 120 python bench.py --workload cfg5 --steps 20
timeout 120 python bench.py --workload cfg2 --steps 20
PE_RESIZE_NOVEC=1 timeout 120 python bench.py --workload cfg2 --steps 20
timeout 120 python bench.py --workload cfg2 --steps 20
PE_RESIZE_NOVEC=1 timeout 120 python bench.py --workload cfg2 --steps 20
} > gpurun_out/bench_$TAG.log 2>&1
grep -o '"value": [0-9.]*\|"workload": "[^"]*"\|"frac": [0-9.]*' gpurun_out/bench_$TAG.log
