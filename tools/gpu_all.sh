#!/bin/bash
# quick full regression + headline bench + secondary configs
TAG=${1:-all}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/bench_$TAG.log 2>&1; grep -o '"value": [0-9.]*\|"single_frame_launch_us": [0-9.]*' gpurun_out/bench_$TAG.log | head -3
rm -f gpurun_out/bench_cfgs_$TAG.log
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
grep -o '"value": [0-9.]*\|"frac": [0-9.]*' gpurun_out/bench_cfgs_$TAG.log | paste - -
