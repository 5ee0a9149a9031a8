#!/bin/bash
# cfg5 sub-record at N ranks: NCCL broadcast vs the multicast publish kernel, with SMs left free for the transfer's CTAs
N=${1:-2}; TAG=${2:-mc}; shift; shift
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --e2e-frames 4 --e2e-steps 1 > gpurun_out/cfg5_mc_last.log 2>&1
  tail -1 gpurun_out/cfg5_mc_last.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['cfg5']
print('$*: cfg5 %.0f clip frames/s, %.1f us / output frame, operand %.0f GB/s, parity %s, headline %.0f fps' % (c['value'], 1e3*c['ms_per_output_frame'], c['broadcast_gbs'] or 0, c['parity_all_ranks'], d['value']))" >> gpurun_out/cfg5_mc_$TAG.log 2>&1 || tail -20 gpurun_out/cfg5_mc_last.log >> gpurun_out/cfg5_mc_$TAG.log
}
for R in ${RESERVES:-0 8 16 32}; do
  run PE_CFG5_TRANSPORT=nccl PE_CFG5_SM_RESERVE=$R
done
for R in ${MC_RESERVES:-0 16}; do
  run PE_CFG5_TRANSPORT=multicast PE_CFG5_SM_RESERVE=$R
done
cat gpurun_out/cfg5_mc_$TAG.log
