#!/bin/bash
# round-end pass: all GPU tests, smoke, both bench arms, secondary configs, launch list, k_fused3 capture, memcheck, CPU baselines of
# every BASELINE config on the box's host cores, weed_layer drop-in cost
TAG=${1:-end}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv > gpurun_out/gpu_$TAG.txt 2>&1
lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket" >> gpurun_out/gpu_$TAG.txt
cp MEASURED_PEAKS.json gpurun_out/MEASURED_PEAKS_$TAG.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_$TAG.log 2>&1; tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; tail -1 gpurun_out/smoke_$TAG.log
timeout 400 python bench.py > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-300
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.log 2>&1; tail -1 gpurun_out/bench_ref_$TAG.log | cut -c1-200
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "fused_fast_path or fused_chain or cvt_resize or planar or config or crossfade or interp" > gpurun_out/memcheck_$TAG.log 2>&1; tail -3 gpurun_out/memcheck_$TAG.log
timeout 600 python tools/cpu_baselines.py --out gpurun_out/cpu_baselines_$TAG.json --seconds 2 > gpurun_out/cpu_baselines_$TAG.log 2>&1; tail -8 gpurun_out/cpu_baselines_$TAG.log
timeout 300 python tools/measure_layer_dropin.py > gpurun_out/layer_dropin_$TAG.json 2>&1; tail -2 gpurun_out/layer_dropin_$TAG.json | cut -c1-600
ls gpurun_out | tail -20
# one ncu --set full capture of the dominant kernel of the headline, of config 5 and of config 2
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused3 -s 3 -c 1 -o gpurun_out/prof_f3_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_full_f3_$TAG.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_yuv_march -s 2 -c 1 -o gpurun_out/prof_cfg5_$TAG python bench.py --workload cfg5 --steps 2 --warmup 3 > gpurun_out/ncu_full_cfg5_$TAG.log 2>&1
timeout 200 python tools/time_new_kernels.py > gpurun_out/new_kernels_$TAG.jsonl 2>&1
timeout 200 python tools/f3_422_probe.py > gpurun_out/f3_422_$TAG.jsonl 2>&1
ls -la gpurun_out | tail -30
