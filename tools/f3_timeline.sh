#!/bin/bash
# where a single-frame k_fused3 launch spends its time: rebuild with -DPE_F3_TIMELINE (globaltimer stamps), run the batch-1 bench
PE_NVCC_EXTRA=-DPE_F3_TIMELINE python -m lives_b200.build > /dev/null 2>&1
PE_F3_TIMELINE_DUMP=1 timeout 300 python bench.py --batch 1 --steps 6 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 2>&1 | grep -A12 "f3 timeline (ns" | tail -13
PE_F3_TIMELINE_DUMP=1 timeout 300 python bench.py --batch 32 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --e2e-frames 2 --e2e-steps 1 2>&1 | grep "f3 timeline" | tail -3
