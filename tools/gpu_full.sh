#!/bin/bash
# full GPU pass: all -m gpu tests, smoke, bench line (with CPU baseline), reference arm, secondary configs, launch list,
# ncu full profile of the fused kernel, compute-sanitizer memcheck of the fused-kernel tests
TAG=${1:-full}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
cp MEASURED_PEAKS.json gpurun_out/MEASURED_PEAKS_$TAG.json 2>/dev/null
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "fused_fast_path or fused_chain or planar or resize or config or permutation or crossfade or yuv444p or chroma_resampling or packed422 or clamping or compositor or yuv_family" > gpurun_out/memcheck_$TAG.log 2>&1; tail -4 gpurun_out/memcheck_$TAG.log
ls -la gpurun_out | tail -20
