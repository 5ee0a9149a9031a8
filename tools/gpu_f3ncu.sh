#!/bin/bash
# ncu full capture (with source) of k_fused3 in the headline bench
TAG=${1:-f3}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
