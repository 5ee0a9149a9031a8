#!/bin/bash
# PCIe ceiling + e2e A/B (linear vs 2-D plane copies)
TAG=${1:-s4b}
mkdir -p gpurun_out
timeout 120 python tools/pcie_probe.py > gpurun_out/pcie_$TAG.log 2>&1; cat gpurun_out/pcie_$TAG.log
for v in "" "PE_HOST_COPY2D=1"; do
  env $v timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"e2e": {[^}]*}' | cut -c1-200
done | tee gpurun_out/e2e_$TAG.log
nvidia-smi -q | grep -i -A4 "GPU Link Info" | head -12
