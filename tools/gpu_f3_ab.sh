#!/bin/bash
# A/B of k_fused3's bg ring: per-lane cp.async (LDGSTS) vs one elected-lane cp.async.bulk per row (UBLKCP + mbarrier)
TAG=${1:-ab}
mkdir -p gpurun_out
for T in 0 1; do
  echo "== PE_F3_TMA=$T" >> gpurun_out/f3_$TAG.log
  PE_F3_TMA=$T timeout 600 python -m pytest tests -m gpu -x -q -k "fused or smoke or headline" 2>&1 | tail -2 >> gpurun_out/f3_$TAG.log
  for B in 32 1; do
    PE_F3_TMA=$T timeout 300 python bench.py --batch $B --steps 100 --no-cpu-baseline --no-sub-records --e2e-frames 4 --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('batch %d: %.0f fps, kernel %.4f ms, frac %.4f, single-frame %.1f us' % (d['config']['frames_per_step_per_gpu'], d['value'], r['kernel_ms'], r['frac'], r['single_frame_launch_us']))" >> gpurun_out/f3_$TAG.log
  done
done
cat gpurun_out/f3_$TAG.log
