#!/bin/bash
# k_fused3 variants: every tmp_variants/libpe_b200_<name>.so x PE_F3_SPREAD 0 / 1, batch 32 (and the parity tests once per library)
TAG=${1:-var}
mkdir -p gpurun_out
cp lives_b200/libpe_b200.so /tmp/libpe_b200_base.so
for LIB in /tmp/libpe_b200_base.so tmp_variants/libpe_b200_*.so; do
  cp $LIB lives_b200/libpe_b200.so
  timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused" 2>&1 | tail -1 >> gpurun_out/variants_$TAG.log
  for S in 0 1; do
    for B in ${BATCH_LIST:-32 1}; do
    PE_F3_SPREAD=$S timeout 300 python bench.py --batch $B --steps 100 --no-cpu-baseline --no-sub-records --e2e-frames 4 --e2e-steps 1 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$LIB SPREAD=$S batch %d: %.0f fps, kernel %.4f ms, frac %.4f' % (d['config']['frames_per_step_per_gpu'], d['value'], r['kernel_ms'], r['frac']))" >> gpurun_out/variants_$TAG.log 2>&1
    done
  done
done
cp /tmp/libpe_b200_base.so lives_b200/libpe_b200.so
cat gpurun_out/variants_$TAG.log
