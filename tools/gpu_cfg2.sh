#!/bin/bash
# config 2 (16 x 1080p YUV420P -> RGBA32 -> 1280x720): fused k_cvt_resize vs the unfused pair, launch list and one full ncu capture
TAG=${1:-cfg2}
mkdir -p gpurun_out
python bench.py --workload cfg2 --steps 50 > gpurun_out/bench_cfg2_$TAG.log 2>&1; tail -1 gpurun_out/bench_cfg2_$TAG.log | cut -c1-400
PE_NO_CVT_RESIZE=1 python bench.py --workload cfg2 --steps 50 > gpurun_out/bench_cfg2_unfused_$TAG.log 2>&1; tail -1 gpurun_out/bench_cfg2_unfused_$TAG.log | cut -c1-400
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_cvt_resize -s 3 -c 1 -o gpurun_out/prof_cfg2_$TAG python bench.py --workload cfg2 --steps 2 --warmup 3 > gpurun_out/ncu_cfg2_$TAG.log 2>&1
tail -2 gpurun_out/ncu_cfg2_$TAG.log
