#!/usr/bin/env python
"""How fast is the operand exchange alone?  pe_mc_publish (multicast stores) for several CTA counts vs ncclBroadcast, 199 MB groups.
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/mc_probe.py"""
import os
import sys
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
import lives_b200 as lb  # noqa: E402
from lives_b200 import shard  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = lb.Engine(device=local)
K, H, W = 8, 2160, 3840
src = torch.randint(0, 256, (K, H, W * 3), dtype=torch.uint8, device=dev)
nbytes = src.numel()
ring = shard.OperandMulticast(eng, (K, H, W * 3), nslots=2)
side = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(side)   # a real stream: the legacy stream's handle is 0, which pe_mc_publish reads as 'the engine's stream'
st = torch.cuda.current_stream()


def timed(fn, n=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    for _ in range(n):
        fn()
    e1.record(st)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for ctas in (16, 32, 64, 148, 296):
    def pub():
        if rank == 0:
            eng.mc_publish(ring.hdl.multicast_ptr, src.data_ptr(), nbytes, st.cuda_stream, ctas)
        ring.hdl.barrier(channel=0)
    ms = timed(pub)
    if rank == 0:
        print("pe_mc_publish %3d CTAs: %.3f ms per 199 MB group = %.0f GB/s (N = %d)" % (ctas, ms, nbytes / ms / 1e6, world), flush=True)
# the same kernel with a unicast peer address as destination (plain P2P stores to one receiver)
if world > 1:
    peer = ring.hdl.buffer_ptrs[(rank + 1) % world] if rank == 0 else 0
    def uni():
        if rank == 0:
            eng.mc_publish(peer, src.data_ptr(), nbytes, st.cuda_stream, 148)
        ring.hdl.barrier(channel=0)
    ms = timed(uni)
    if rank == 0:
        print("same kernel, unicast P2P stores to one peer: %.3f ms = %.0f GB/s" % (ms, nbytes / ms / 1e6), flush=True)
ok = True
flag = torch.tensor([int(ok)], device=dev); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
buf = src.clone()
ms = timed(lambda: dist.broadcast(buf, src=0))
if rank == 0:
    print("ncclBroadcast: %.3f ms = %.0f GB/s; multicast data identical on all ranks: %s" % (ms, nbytes / ms / 1e6, bool(flag.item())), flush=True)
# P2P: every other rank pulls from rank 0 with its copy engine
peer0 = ring.hdl.get_buffer(0, (K, H, W * 3), torch.uint8)
if rank == 0:
    ring.buf[0].copy_(src)
torch.cuda.synchronize(); dist.barrier()
dst = torch.empty_like(src)
ms = timed(lambda: dst.copy_(peer0) if rank != 0 else None)
t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print("copy-engine pull from rank 0 by every other rank at once: %.3f ms = %.0f GB/s per receiver" % (t.item(), nbytes / t.item() / 1e6), flush=True)
dist.destroy_process_group()
