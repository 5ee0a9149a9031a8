#!/bin/bash
# config 5 inside bench.py's line under each operand transport: tools/gpu_cfg5_transports.sh <N> [transport ...]
N=${1:-2}; shift
[ $# -eq 0 ] && set -- chain nccl
P=29500
for T in "$@"; do
  P=$((P+1))
  PE_CFG5_CHAIN_LAG=${LAG:-2} PE_CFG5_TRANSPORT=${T%%:*} PE_CFG5_SM_RESERVE=$(echo $T | awk -F: '{print ($2==""? ($1=="chain"?0:8) : $2)}') timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --e2e-frames 2 --e2e-steps 1 2>gpurun_out/cfg5_${T%%:*}_$N.err | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); c = d['configs']['cfg5']
    print('$T N=$N: cfg5 %.0f clip frames/s, %.1f us / output frame, operand %.0f GB/s, parity %s, headline %.0f fps | %s' % (c['value'], c['ms_per_output_frame'] * 1e3, c['broadcast_gbs'] or 0, c['parity_all_ranks'], d['value'], c['workload'][120:300]))
except Exception as e:
    print('$T N=$N: failed', e)
"
  tail -3 gpurun_out/cfg5_${T%%:*}_$N.err | cut -c1-300
done
