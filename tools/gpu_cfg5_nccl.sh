#!/bin/bash
# cfg5 sub-record of the driver's line under NCCL tuning variables (N ranks): is ncclBroadcast channel-bound?
N=${1:-2}; TAG=${2:-nccl}
mkdir -p gpurun_out
run() {
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --e2e-frames 4 --e2e-steps 1 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['configs']['cfg5']
print('$*: cfg5 %.0f clip frames/s, %.1f us / output frame, broadcast %.0f GB/s, headline %.0f fps' % (c['value'], 1e3*c['ms_per_output_frame'], c['broadcast_gbs'] or 0, d['value']))" >> gpurun_out/cfg5_nccl_$TAG.log 2>&1
}
run NCCL_DEBUG=WARN
run NCCL_MIN_NCHANNELS=16
run NCCL_MIN_NCHANNELS=32
run NCCL_MIN_NCHANNELS=32 NCCL_BUFFSIZE=16777216
run NCCL_ALGO=Ring NCCL_PROTO=Simple NCCL_MIN_NCHANNELS=32
run NCCL_MIN_CTAS=32
cat gpurun_out/cfg5_nccl_$TAG.log
