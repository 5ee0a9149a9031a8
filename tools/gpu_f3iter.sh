#!/bin/bash
# k_fused3 iteration: the fused parity tests, then the headline bench (batch 32 and single frame come out of the same line)
TAG=${1:-f3i}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_resize_interp.py -m gpu -q -x -k "fused or headline or smoke" > gpurun_out/pytest_f3_$TAG.log 2>&1; tail -3 gpurun_out/pytest_f3_$TAG.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 300 python bench.py --steps 100 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r = d['roofline']
print('fps %.0f  kernel_ms %.4f  frac %.4f  single_frame_us %.1f' % (d['value'], r['kernel_ms'], r['frac'], r['single_frame_launch_us']))"
