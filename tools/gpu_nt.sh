#!/bin/bash
# k_fused3 threads-per-CTA experiment: rebuild on the GPU box with PE_F3_NT and bench
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1"
for nt in "$@"; do
  PE_F3_NT=$nt python -m lives_b200.build --force > gpurun_out/build_nt$nt.log 2>&1
  timeout 300 python -m pytest tests -m gpu -q -k "fused_fast_path or headline" > gpurun_out/pytest_nt$nt.log 2>&1; tail -1 gpurun_out/pytest_nt$nt.log
  for i in 1 2; do r=$(timeout 120 $B 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -1); echo "nt=$nt $r"; done
done 2>&1 | tee gpurun_out/nt_sweep.log
