#!/bin/bash
# multi-GPU pass (gpurun --gpus N): the driver's bench launch at N ranks, our arm and the reference arm, and BASELINE config 5
N=${1:-2}; TAG=${2:-multi}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$TAG.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/bench_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/bench_n${N}_$TAG.log | cut -c1-400
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_n${N}_$TAG.log | cut -c1-200
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/run_cfg5.py > gpurun_out/cfg5_n${N}_$TAG.log 2>&1; tail -1 gpurun_out/cfg5_n${N}_$TAG.log | cut -c1-300
timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1_$TAG.log 2>&1; tail -1 gpurun_out/bench_n1_$TAG.log | cut -c1-200
