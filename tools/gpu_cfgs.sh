#!/bin/bash
# kernel-level numbers of the secondary BASELINE configs + their parity tests
TAG=${1:-cfgs}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "config or cfg or planar_yuv or resize or crossfade" > gpurun_out/pytest_$TAG.log 2>&1; tail -2 gpurun_out/pytest_$TAG.log
rm -f gpurun_out/bench_cfgs_$TAG.log
for w in cfg1 cfg2 cfg3 cfg4 cfg5; do timeout 120 python bench.py --workload $w --steps 20 >> gpurun_out/bench_cfgs_$TAG.log 2>&1; done
grep -o '"value": [0-9.]*\|"workload": "[^"]*"\|"frac": [0-9.]*' gpurun_out/bench_cfgs_$TAG.log
# kernel times of the two planar converters on a 4K 4:2:2 clip and a 1080p 4:2:0 frame (launch list, cold cache)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_cfg5_$TAG.csv python bench.py --workload cfg5 --steps 2 --warmup 3 > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_cfg2_$TAG.csv python bench.py --workload cfg2 --steps 2 --warmup 3 > /dev/null 2>&1
grep -o 'k_[a-z_0-9]*<[^>]*>\|k_[a-z_0-9]*\|"[0-9.]*"$' gpurun_out/launches_cfg5_$TAG.csv | paste - - | sort | uniq -c | sort -rn | head -5
grep -o 'k_[a-z_0-9]*<[^>]*>\|k_[a-z_0-9]*\|"[0-9.]*"$' gpurun_out/launches_cfg2_$TAG.csv | paste - - | tail -8
