#!/bin/bash
# one k_fused3 iteration on the GPU box: fused parity tests, bench line, optional ncu full profile, optional env sweeps
TAG=${1:-it}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "fused or smoke or host or planar420 or unhandled" > gpurun_out/pytest_$TAG.log 2>&1; tail -15 gpurun_out/pytest_$TAG.log
B="python bench.py --steps 30 --warmup 5 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1"
timeout 300 $B > gpurun_out/bench_$TAG.log 2>&1; tail -1 gpurun_out/bench_$TAG.log | cut -c1-200
if [ "$2" = "prof" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 3 -c 1 -o gpurun_out/prof_$TAG python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 > gpurun_out/ncu_$TAG.log 2>&1
fi
shift; shift
# remaining args: "VAR=v1,v2,..." sweeps (one variable at a time)
for spec in "$@"; do
  var=${spec%%=*}; vals=${spec#*=}
  for v in ${vals//,/ }; do
    r=$(env $var=$v timeout 120 $B 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -1)
    echo "$var=$v $r"
  done
done > gpurun_out/sweep_$TAG.log 2>&1
cat gpurun_out/sweep_$TAG.log
for b in 32 64; do r=$(timeout 120 $B --batch $b 2>&1 | tail -1 | grep -o '"value": [0-9.]*' | head -1); echo "batch=$b $r"; done >> gpurun_out/sweep_$TAG.log 2>&1
tail -2 gpurun_out/sweep_$TAG.log
