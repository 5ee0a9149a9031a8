#!/bin/bash
# ncu --set full of the marching converter inside the cfg5 batch (crossfade, 4:2:2) and of the cfg2 batch kernels
TAG=${1:-s4f}
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_yuv_march -s 2 -c 1 -o gpurun_out/prof_march5_$TAG python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/ncu_march5_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_yuv_march -s 2 -c 1 -o gpurun_out/prof_march2_$TAG python bench.py --workload cfg2 --steps 2 --warmup 3 > gpurun_out/ncu_march2_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
