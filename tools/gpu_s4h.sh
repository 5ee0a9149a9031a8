#!/bin/bash
# e2e: pipeline depth and batch length
TAG=${1:-s4h}
mkdir -p gpurun_out
for v in "PE_PIPE_SLOTS=3" "PE_PIPE_SLOTS=4" "PE_PIPE_SLOTS=6" "PE_PIPE_SLOTS=2" "PE_HOST_COPY2D=1"; do
  echo "$v: $(env $v timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"e2e": {"value": [0-9.]*')"
done | tee gpurun_out/e2e_$TAG.log
echo "frames 192: $(timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --e2e-frames 192 --e2e-steps 2 2>&1 | tail -1 | grep -o '"e2e": {"value": [0-9.]*')" | tee -a gpurun_out/e2e_$TAG.log
timeout 300 python -m pytest tests -m gpu -q -k "host" 2>&1 | tail -2
