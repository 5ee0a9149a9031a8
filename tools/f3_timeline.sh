#!/bin/bash
# where a k_fused3 launch spends its time: rebuild with -DPE_F3_TIMELINE (globaltimer stamps), run the bench at --batch ${1:-1}
PE_NVCC_EXTRA=-DPE_F3_TIMELINE python -m lives_b200.build > /dev/null 2>&1
for b in ${@:-1}; do
PE_F3_TIMELINE_DUMP=1 timeout 300 python bench.py --batch $b --steps 4 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 2>&1 | grep -A12 "f3 timeline (ns" | grep -v '^{' | tail -12
done
