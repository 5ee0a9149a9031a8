#!/bin/bash
# one optimisation iteration on the GPU box: fused parity tests, bench line, ncu full profile of the fused kernel
TAG=${1:-iter}
timeout 600 python -m pytest tests -m gpu -q -k "fused or smoke" > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-260
ncu --set full --clock-control none --import-source on -k regex:k_fused2 -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
