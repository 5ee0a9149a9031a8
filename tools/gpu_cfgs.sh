#!/bin/bash
# kernel-level numbers of the secondary BASELINE configs + their parity tests
TAG=${1:-cfgs}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "config or cfg or planar_yuv or resize or crossfade" > gpurun_out/pytest_$TAG.log 2>&1; tail -2 gpurun_out/pytest_$TAG.log
rm -f gpurun_out/bench_cfgs_$TAG.log
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
grep -o '"value": [0-9.]*\|"workload": "[^"]*"\|"frac": [0-9.]*' gpurun_out/bench_cfgs_$TAG.log
