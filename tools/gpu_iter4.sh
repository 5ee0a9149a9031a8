#!/bin/bash
# quick GPU iteration: a pytest selection (-k "$2") and, optionally, bench workloads ("$3")
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "$2" > gpurun_out/pytest_$TAG.log 2>&1; grep -E "^E  |passed|failed|Error" gpurun_out/pytest_$TAG.log | head -30
for w in $3; do timeout 120 python bench.py --workload $w --steps 20 2>&1 | tail -1 | grep -o '"value": [0-9.]*\|"frac": [0-9.]*' | paste - - ; done | tee gpurun_out/bench_$TAG.log
