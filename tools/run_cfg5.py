#!/usr/bin/env python
"""BASELINE config 5 under torchrun (one rank per GPU): each rank owns one clip (4K YUV422P), the shared transition operand
(4K RGB24) lives on rank 0 and is broadcast over NCCL once per output frame; every rank converts its clip to RGB24 and
crossfades it with the operand ('chroma blend' bf=128).  First a small-frame parity check against the oracle on every rank,
then timing.  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_cfg5.py"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import lives_b200 as lb  # noqa: E402
from lives_b200 import shard  # noqa: E402
import pe_testlib as T  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    eng = lb.Engine(device=local)
    o = T.oracle()
    # ---- parity at 256 x 64: clip seed 10 + rank, operand seed 99 on rank 0 only
    w, h, bf = 256, 64, 128
    y, u, v = T.make_yuv_planar(np.random.default_rng(10 + rank), w, h, True, True)
    op_host = T.make_packed(np.random.default_rng(99), w, h, 3)
    exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], 0, 0, 1, 0, 1, T.Q_HIGH, 1, None)
    o.pe_or_simple_blend(0, 1, T.ptr(exp), exp.strides[0], T.ptr(op_host), op_host.strides[0], T.ptr(exp), exp.strides[0], w, h, bf, op_host.size)
    operand = torch.from_numpy(op_host).to(dev) if rank == 0 else torch.zeros(op_host.shape, dtype=torch.uint8, device=dev)
    clip = lb.Layer.from_host(eng, lb.WEED_PALETTE_YUV422P, w, h, [y, u, v], yuv_subspace=1)
    shard.multitrack_crossfade(eng, clip, operand, w, h, bf)
    got = clip.to_host()[0]
    ok = bool((got[:, :w * 3] == exp[:, :w * 3]).all())
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    # ---- timing at 4K: per output frame = convert + broadcast + blend
    W, H = 3840, 2160
    g = torch.Generator(device=dev)
    g.manual_seed(10 + rank)
    yy = torch.randint(16, 236, (H, W), dtype=torch.uint8, device=dev, generator=g)
    uu = torch.randint(16, 241, (H, W // 2), dtype=torch.uint8, device=dev, generator=g)
    vv = torch.randint(16, 241, (H, W // 2), dtype=torch.uint8, device=dev, generator=g)
    # two operand buffers: the broadcast of output frame t + 1 overlaps the conversion and the blend of frame t
    operands = [torch.randint(0, 256, (H, W * 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    steps = 40
    it = [0]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def frame():
        clip = lb.Layer.wrap_device(eng, lb.WEED_PALETTE_YUV422P, W, H, [yy.data_ptr(), uu.data_ptr(), vv.data_ptr()], [W, W // 2, W // 2], yuv_subspace=1)
        shard.multitrack_crossfade(eng, clip, operands[it[0] & 1], W, H, 128)
        it[0] += 1
        clip.free()  # stream ordered: the block returns to the pool behind the kernels that use it

    for _ in range(3):
        frame()
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(steps):
        frame()
    eng.sync()
    ev1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    # ---- the same in groups of K output frames: one broadcast of K operand frames, one kernel launch for the K crossfades
    K = 8
    groups = [torch.randint(0, 256, (K, H, W * 3), dtype=torch.uint8, device=dev, generator=g) for _ in range(2)]
    # parity of the grouped path at 256 x 64 against the per-frame result above
    gop = torch.from_numpy(op_host).to(dev) if rank == 0 else torch.zeros(op_host.shape, dtype=torch.uint8, device=dev)
    gop = gop.unsqueeze(0).repeat(3, 1, 1).contiguous()
    gclips = [lb.Layer.from_host(eng, lb.WEED_PALETTE_YUV422P, w, h, [y, u, v], yuv_subspace=1) for _ in range(3)]
    shard.multitrack_crossfade_group(eng, gclips, gop, w, h, bf)
    gok = all(bool((c.to_host()[0][:, :w * 3] == exp[:, :w * 3]).all()) for c in gclips)
    gflag = torch.tensor([int(gok)], device=dev)
    dist.all_reduce(gflag, op=dist.ReduceOp.MIN)
    gsteps = 10

    def group():
        clips = [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_YUV422P, W, H, [yy.data_ptr(), uu.data_ptr(), vv.data_ptr()], [W, W // 2, W // 2],
                                      yuv_subspace=1) for _ in range(K)]
        shard.multitrack_crossfade_group(eng, clips, groups[it[0] & 1], W, H, 128)
        it[0] += 1
        for c in clips:
            c.free()

    for _ in range(2):
        group()
    dist.barrier()
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(gsteps):
        group()
    eng.sync()
    ev1.record()
    torch.cuda.synchronize()
    gms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(gms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"config": "cfg5 grouped: %d clips x 4K YUV422P -> RGB24 + crossfade, %d operand frames per broadcast, one launch per group" % (world, K),
                          "n_gpus": world, "parity_all_ranks": bool(gflag.item()), "clip_frames_per_s": world * gsteps * K / (gms.item() / 1e3),
                          "ms_per_output_frame": gms.item() / (gsteps * K), "broadcast_bytes": K * H * W * 3}), flush=True)
    if rank == 0:
        print(json.dumps({"config": "cfg5: %d clips x 4K YUV422P -> RGB24 + crossfade with broadcast operand" % world, "n_gpus": world,
                          "parity_all_ranks": bool(flag.item()), "clip_frames_per_s": world * steps / (ms.item() / 1e3),
                          "ms_per_output_frame": ms.item() / steps, "broadcast_bytes": H * W * 3}), flush=True)
    dist.destroy_process_group()
    if not flag.item() or not gflag.item():
        sys.exit(1)


if __name__ == "__main__":
    main()
