"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: frame partitioning and the operand broadcast."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lives_b200 import shard


def test_frames_for_rank_partitions_exactly():
    for n, w in ((256, 8), (256, 1), (7, 2), (3, 4), (0, 2), (33, 8)):
        blocks = [shard.frames_for_rank(n, r, w) for r in range(w)]
        assert blocks[0][0] == 0 and blocks[-1][1] == n
        for a, b in zip(blocks, blocks[1:]):
            assert a[1] == b[0]
        sizes = [hi - lo for lo, hi in blocks]
        assert max(sizes) - min(sizes) <= 1
    assert shard.frames_for_rank(256, 3, 8) == (96, 128)  # config 4: 32 frames per GPU
    assert shard.clip_for_rank(8, 5, 8) == [5] and shard.clip_for_rank(8, 1, 2) == [1, 3, 5, 7]
    with pytest.raises(ValueError):
        shard.frames_for_rank(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.frames_for_rank(10, rank, world)
        # shared transition operand: a 64 x 36 RGB24 frame (rowstride 192) owned by rank 0
        h, rs = 36, 192
        ref = torch.from_numpy(np.random.default_rng(99).integers(0, 256, (h, rs), dtype=np.uint8))
        buf = ref.clone() if rank == 0 else torch.zeros((h, rs), dtype=torch.uint8)
        shard.broadcast_operand(buf, src=0)
        # a group of K = 3 operand frames in one broadcast (shard.multitrack_crossfade_group)
        gref = torch.from_numpy(np.random.default_rng(100).integers(0, 256, (3, h, rs), dtype=np.uint8))
        gbuf = gref.clone() if rank == 0 else torch.zeros((3, h, rs), dtype=torch.uint8)
        shard.broadcast_operand(gbuf, src=0)
        assert bool((gbuf == gref).all())
        hist = torch.bincount(buf.flatten().long(), minlength=256)
        shard.allreduce_histogram(hist)
        q.put((rank, (lo, hi), bool((buf == ref).all()), int(hist.sum())))
    finally:
        dist.destroy_process_group()


def test_operand_broadcast_and_partition_world_size_2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == (0, 5) and res[1][1] == (5, 10)
    assert res[0][2] and res[1][2]
    assert res[0][3] == res[1][3] == 2 * 36 * 192


def test_operand_chain_schedule_is_safe():
    """shard.chain_schedule (OperandChain's lockstep rule), simulated for 2 .. 8 ranks, lag 1 / 2 and lag + 2 .. lag + 3 slots: every rank
    consumes group step - lag * rank, a pull always finds that group in its predecessor's slot -- copied there `lag` steps earlier, so
    the barrier that ended THAT step is the only one the pull depends on --, and a slot is only rewritten a whole step after this rank
    consumed its previous group and the successor pulled it"""
    from lives_b200 import shard
    for world in (2, 3, 4, 8):
        for lag in (1, 2):
            for nslots in (lag + 2, lag + 3):
                slots = [[None] * nslots for _ in range(world)]       # (group, step it was written) held by rank r, slot s
                pulled_at = [dict() for _ in range(world)]            # rank r: group -> step rank r + 1 pulled it
                consumed = [set() for _ in range(world)]
                for step in range(60):
                    moves = []
                    for r in range(world):
                        g, s = shard.chain_schedule(step, r, nslots, lag)
                        if g is None:
                            assert step < lag * r
                            continue
                        assert g == step - lag * r and s == g % nslots
                        if r > 0:
                            held = slots[r - 1][s]
                            assert held is not None and held[0] == g and held[1] <= step - lag, (world, lag, nslots, step, r)
                        old = slots[r][s]
                        if old is not None:
                            assert old[0] == g - nslots and old[0] in consumed[r]
                            # the successor's pull of the old group ended at least one whole step (one barrier) before `step - lag`'s barrier
                            assert r == world - 1 or pulled_at[r][old[0]] <= step - lag, (world, lag, nslots, step, r)
                        moves.append((r, s, g))
                    for r, s, g in moves:   # all copies of a step run at once
                        slots[r][s] = (g, step)
                        consumed[r].add(g)
                        if r > 0:
                            pulled_at[r - 1][g] = step
                for r in range(world):
                    assert consumed[r] == set(range(60 - lag * r))
