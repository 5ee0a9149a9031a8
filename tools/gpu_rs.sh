#!/bin/bash
mkdir -p gpurun_out
for cfg in "64 32" "32 32" "32 16" "64 16" "128 16" "16 16" "64 8"; do set -- $cfg
  r=$(PE_RESIZE_TW=$1 PE_RESIZE_TH=$2 timeout 120 python bench.py --workload cfg2 --steps 20 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -1)
  echo "tw=$1 th=$2 $r"
done 2>&1 | tee gpurun_out/resize_tiles.log
