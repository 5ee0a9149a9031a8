#!/bin/bash
# single-frame launch of k_fused3: latency from the bench line, and one ncu full capture of a 1-frame launch
TAG=${1:-single}
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1"
timeout 300 python -m pytest tests -m gpu -q -k "fused_fast_path or headline" > gpurun_out/pytest_$TAG.log 2>&1; tail -1 gpurun_out/pytest_$TAG.log
timeout 200 $B > gpurun_out/bench_$TAG.log 2>&1; grep -o '"value": [0-9.]*\|"single_frame_launch_us": [0-9.]*' gpurun_out/bench_$TAG.log | head -3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 12 -c 1 -o gpurun_out/prof_$TAG python bench.py --batch 1 --steps 20 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
