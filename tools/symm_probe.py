#!/usr/bin/env python
"""Does torch's symmetric memory rendezvous work on this box, and does it hand out an NVSwitch multicast address?
torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/symm_probe.py"""
import os
import torch
import torch.distributed as dist
import torch.distributed._symmetric_memory as symm

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
try:
    t = symm.empty(64 << 20, dtype=torch.uint8, device=dev)
    h = symm.rendezvous(t, dist.group.WORLD)
    print("rank %d: buffer_ptrs %s multicast_ptr 0x%x has_multicast %s" % (rank, [hex(p) for p in h.buffer_ptrs], h.multicast_ptr,
                                                                         symm._SymmetricMemory.has_multicast_support(DeviceType := torch._C._autograd.DeviceType.CUDA, local) if hasattr(symm._SymmetricMemory, "has_multicast_support") else "?"), flush=True)
    # peer write through the P2P mapping: rank 0 fills rank 1's buffer
    t.zero_()
    h.barrier()
    if rank == 0 and world > 1:
        peer = h.get_buffer(1, (1024,), torch.uint8)
        peer.fill_(7)
    h.barrier()
    if rank == 1:
        print("rank 1 sees", int(t[:1024].sum().item()), "(expect 7168)", flush=True)
except Exception as e:  # noqa: BLE001
    print("rank %d: symmetric memory unavailable: %r" % (rank, e), flush=True)
dist.destroy_process_group()
