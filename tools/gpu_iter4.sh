#!/bin/bash
# iteration pass: GPU tests, the secondary-kernel timings, config 5 kernel rate
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python tools/time_new_kernels.py > gpurun_out/new_kernels_$TAG.jsonl 2>&1; cat gpurun_out/new_kernels_$TAG.jsonl | cut -c1-250
for w in cfg5 cfg2; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done; cut -c1-400 gpurun_out/bench_cfgs_$TAG.log
