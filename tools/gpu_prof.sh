#!/bin/bash
# ncu full capture of one kernel of a bench workload: tools/gpu_prof.sh <tag> <kernel regex> <workload> [skip]
TAG=$1; K=$2; W=$3; SKIP=${4:-6}
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -o gpurun_out/prof_$TAG python bench.py --workload $W --steps 2 --warmup 3 > gpurun_out/ncu_$TAG.log 2>&1
ls -la gpurun_out/prof_$TAG.ncu-rep
