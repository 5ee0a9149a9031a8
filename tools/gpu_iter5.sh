#!/bin/bash
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python tools/time_new_kernels.py > gpurun_out/new_kernels_$TAG.jsonl 2>&1; cut -c1-200 gpurun_out/new_kernels_$TAG.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_chroma_upsample_packed|k_quad_chroma|k_copy2d" -c 30 --csv --log-file gpurun_out/ncu_upsample_$TAG.csv python tools/time_new_kernels.py > gpurun_out/ncu_upsample_$TAG.log 2>&1
python tools/ncu_compact.py gpurun_out/ncu_upsample_$TAG.csv
