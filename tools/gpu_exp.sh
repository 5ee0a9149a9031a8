#!/bin/bash
# k_fused3 build-flag experiments: each argument is a set of nvcc -D flags (quote it); rebuilds on the GPU box and benches
mkdir -p gpurun_out
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1"
for flags in "$@"; do
  PE_NVCC_EXTRA="$flags" python -m lives_b200.build --force > gpurun_out/build_exp.log 2>&1 || { echo "build failed: $flags"; tail -3 gpurun_out/build_exp.log; continue; }
  timeout 300 python -m pytest tests -m gpu -q -k "fused_fast_path or headline" 2>&1 | tail -1
  for i in 1 2; do r=$(timeout 120 $B 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -1); echo "[$flags] $r"; done
done 2>&1 | tee gpurun_out/exp_sweep.log
