#!/bin/bash
# BASELINE config 5 on N GPUs (gpurun --gpus N) + its single-GPU kernel-level number
N=${1:-2}; TAG=${2:-c5}
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/run_cfg5.py > gpurun_out/cfg5_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/cfg5_n${N}_$TAG.log | cut -c1-300
timeout 120 python bench.py --workload cfg5 --steps 20 > gpurun_out/bench_cfg5_$TAG.log 2>&1; tail -1 gpurun_out/bench_cfg5_$TAG.log | cut -c1-400
