#!/bin/bash
# GPU pass without the profiler: all -m gpu tests, smoke, bench line, reference arm, secondary configs
TAG=${1:-pass}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
cp MEASURED_PEAKS.json gpurun_out/MEASURED_PEAKS_$TAG.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -5 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-600
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-300
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
tail -5 gpurun_out/bench_cfgs_$TAG.log | cut -c1-300
