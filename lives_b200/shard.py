"""Multi-GPU sharding of the pixel path: one process per GPU, frames / clips are independent units.

  * render batch (BASELINE config 4, bench.py): contiguous blocks of frames per rank, NO data-path collective;
  * multitrack (config 5): one clip per rank; the shared transition operand (one RGB24 / RGBA32 frame) lives on the
    owning rank and is sent with a single broadcast per output frame (NCCL over NVLink on GPUs, gloo in the CPU tests).

Plumbing only (torch.distributed); every pixel is computed by the engine's CUDA kernels on the rank's own GPU.
"""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames, rank, world_size):
    """contiguous block [lo, hi) of a batch of n_frames for this rank; blocks differ by at most one frame"""
    if world_size <= 0 or not 0 <= rank < world_size:
        raise ValueError("bad rank %d / world size %d" % (rank, world_size))
    base, rem = divmod(n_frames, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def clip_for_rank(n_clips, rank, world_size):
    """multitrack: clips owned by this rank (round robin when there are more clips than ranks)"""
    return list(range(rank, n_clips, world_size))


def broadcast_operand(buf, src=0, group=None):
    """Broadcast the shared transition operand (a uint8 tensor aliasing the frame's pixel memory) from `src`.
    Returns the same tensor; on the owning rank it is the source, elsewhere it is overwritten."""
    if buf.dtype != torch.uint8 or not buf.is_contiguous():
        raise ValueError("operand must be a contiguous uint8 tensor")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(buf, src=src, group=group)
    return buf


def allreduce_histogram(hist, group=None):
    """optional diagnostics: sum of per-frame histograms over ranks (<= 4 KB)"""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)
    return hist


_ENGINE_STREAMS = {}
_CONSUMED = {}   # (engine id, data_ptr) -> event of the last kernel that read that operand buffer


def _await_buffer(engine, tensor, stream):
    """a transfer into `tensor` may only start once the last kernel that read it has finished -- that buffer's kernel, not the stream's
    latest one: with several operand buffers the transfer of group t + 1 then really overlaps the kernel of group t (waiting on the
    newest kernel, as this module did before, serialised them: 0.48 ms per group = 0.31 ms broadcast + 0.17 ms kernel on 2 GPUs)"""
    ev = _CONSUMED.get((id(engine), tensor.data_ptr()))
    if ev is not None:
        stream.wait_event(ev)


def _mark_consumed(engine, tensor, es):
    ev = torch.cuda.Event()
    ev.record(es)
    _CONSUMED[(id(engine), tensor.data_ptr())] = ev
    return ev


def _engine_stream(engine):
    """the engine's CUDA stream as a torch stream (for event ordering against NCCL, which runs on torch's stream)"""
    key = id(engine)
    if key not in _ENGINE_STREAMS:
        _ENGINE_STREAMS[key] = torch.cuda.ExternalStream(engine.stream)
    return _ENGINE_STREAMS[key]


def multitrack_crossfade(engine, clip_layer, operand_tensor, width, height, blend_factor, src_rank=0, out_palette=1):
    """config 5 on this rank: clip (YUV422P / UYVY / ...) -> RGB24 (convert_layer_palette), operand broadcast from the owner,
    then 'chroma blend' (the crossfade / auto-transition of src/multitrack.h:84) of the clip with the operand, in place.
    `operand_tensor`: uint8 CUDA tensor of height x rowstride bytes on every rank.
    Stream ordered, no host synchronisation: the fused convert + crossfade kernel (engine stream) waits for the broadcast
    (torch's current stream), and the next broadcast into the SAME tensor waits for that kernel; with two or more operand buffers
    the broadcast of output frame t + 1 overlaps the kernel of frame t."""
    from . import engine as E
    es, ts = _engine_stream(engine), torch.cuda.current_stream()
    _await_buffer(engine, operand_tensor, ts)
    broadcast_operand(operand_tensor, src=src_rank)
    arrived = torch.cuda.Event()
    arrived.record(ts)
    es.wait_event(arrived)
    rs = operand_tensor.shape[1] if operand_tensor.dim() == 2 else operand_tensor.numel() // height
    operand = E.Layer.wrap_device(engine, out_palette, width, height, [operand_tensor.data_ptr()], [rs])
    if out_palette in (E.WEED_PALETTE_RGB24, E.WEED_PALETTE_BGR24) and clip_layer.palette in (512, 513, 522):
        E.convert_crossfade(clip_layer, operand, out_palette, 0, blend_factor)  # one kernel: convert + crossfade
    else:
        if not E.convert_layer_palette(clip_layer, out_palette, 0):
            raise RuntimeError("clip conversion failed: " + E.capi.last_error())
        E.simple_blend("chroma blend", clip_layer, operand, clip_layer, blend_factor)
    _mark_consumed(engine, operand_tensor, es)
    return clip_layer


def multitrack_crossfade_group(engine, clip_layers, operand_group, width, height, blend_factor, src_rank=0, out_palette=1):
    """config 5 for a GROUP of K consecutive output frames of this rank's clip (render-to-disk, not realtime): `operand_group` is a
    uint8 CUDA tensor of K x height x rowstride bytes holding K consecutive frames of the shared transition operand; it travels in
    ONE broadcast, and the K conversions + crossfades leave as ONE kernel launch (pe_fx_convert_crossfade_batchv).  The host cost
    per output frame (a Python call, a collective launch, two event waits: ~100 us, four times the kernel) is paid once per
    group.  Same stream ordering as multitrack_crossfade."""
    from . import engine as E
    k = len(clip_layers)
    if operand_group.dim() != 3 or operand_group.shape[0] != k or operand_group.shape[1] != height:
        raise ValueError("operand_group must be K x height x rowstride")
    es, ts = _engine_stream(engine), torch.cuda.current_stream()
    _await_buffer(engine, operand_group, ts)
    broadcast_operand(operand_group, src=src_rank)
    arrived = torch.cuda.Event()
    arrived.record(ts)
    es.wait_event(arrived)
    rs = operand_group.shape[2]
    ops = [E.Layer.wrap_device(engine, out_palette, width, height, [operand_group[i].data_ptr()], [rs]) for i in range(k)]
    if E.convert_crossfade_batchv(clip_layers, ops, out_palette, 0, blend_factor) != k:
        raise RuntimeError("grouped crossfade failed: " + E.capi.last_error())
    _mark_consumed(engine, operand_group, es)
    return clip_layers


class OperandMulticast:
    """The shared operand of config 5 without a collective library on the data path: `nslots` group buffers in SYMMETRIC memory (the
    same allocation mapped on every rank, torch.distributed._symmetric_memory) behind an NVSwitch multicast address.  The owner
    publishes a group with ONE kernel (pe_mc_publish: ld.global -> multimem.st, lives_b200/csrc/pe_kernels_mc.cu) that reads its frames
    once and lands them in every rank's buffer -- the owner's NVLink egress carries the operand once whatever the number of
    receivers, nothing is staged, no receiver runs a copy kernel.  Groups are handed over with the symmetric-memory stream barrier.

    Protocol, step t (slot t % nslots), everything stream ordered, no host synchronisation:
      owner      publish(src)   on its own stream, up to nslots - 1 groups ahead: waits until the slot's previous group (t - nslots) has
                                been consumed everywhere, i.e. until the owner has passed barrier t - nslots + 1
      every rank acquire()      waits for this rank's previous crossfade kernel and (owner) for the publish of group t, then the barrier:
                                behind it group t is in every buffer and every kernel <= t - 1 of every rank has finished
                 consumed(ev)   after launching the kernel that reads the slot
    Raises RuntimeError when the group has no multicast support (the caller keeps the NCCL broadcast)."""

    def __init__(self, engine, slot_shape, nslots=3, src_rank=0, group=None):
        import torch.distributed._symmetric_memory as symm
        self.engine, self.nslots, self.src = engine, nslots, src_rank
        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.buf = symm.empty((nslots,) + tuple(slot_shape), dtype=torch.uint8, device=dev)
        self.hdl = symm.rendezvous(self.buf, self.group)
        if not self.hdl.multicast_ptr:
            raise RuntimeError("no multicast address for this group")
        self.slot_bytes = self.buf[0].numel()
        if self.slot_bytes % 16:
            raise ValueError("a slot must be a multiple of 16 bytes")
        self.pub_stream = torch.cuda.Stream(device=dev)
        self.pub_ev = [None] * nslots      # owner: the publish of the group in the slot
        self.free_ev = {}                  # owner: step -> event behind that step's barrier
        self.kernel_ev = None              # this rank's latest crossfade kernel
        self.t_pub = self.t_acq = 0

    def publish(self, src_tensor, ready_event=None):
        """owner only: group number t_pub from `src_tensor` (contiguous uint8, slot_bytes) into every rank's slot"""
        if self.rank != self.src:
            self.t_pub += 1
            return
        if src_tensor.numel() != self.slot_bytes or not src_tensor.is_contiguous():
            raise ValueError("operand group must be %d contiguous bytes" % self.slot_bytes)
        t, slot = self.t_pub, self.t_pub % self.nslots
        need = t - self.nslots + 1          # the barrier that proves the slot free
        if need >= 0:
            if need not in self.free_ev:
                raise RuntimeError("publish runs more than nslots - 1 groups ahead of acquire")
            self.pub_stream.wait_event(self.free_ev[need])
        if ready_event is not None:
            self.pub_stream.wait_event(ready_event)
        self.engine.mc_publish(self.hdl.multicast_ptr + slot * self.slot_bytes, src_tensor.data_ptr(), self.slot_bytes,
                               self.pub_stream.cuda_stream)
        ev = torch.cuda.Event()
        ev.record(self.pub_stream)
        self.pub_ev[slot] = ev
        self.t_pub += 1

    def acquire(self):
        """every rank: the local buffer of group t_acq, valid behind the returned event (record it -- the caller's kernel stream waits)"""
        t, slot = self.t_acq, self.t_acq % self.nslots
        ts = torch.cuda.current_stream()
        if self.rank == self.src:
            if self.pub_ev[slot] is None or self.t_pub <= t:
                raise RuntimeError("acquire before publish")
            ts.wait_event(self.pub_ev[slot])
        if self.kernel_ev is not None:
            ts.wait_event(self.kernel_ev)
        self.hdl.barrier(channel=0)
        ev = torch.cuda.Event()
        ev.record(ts)
        if self.rank == self.src:
            self.free_ev[t] = ev
            self.free_ev.pop(t - self.nslots - 1, None)
        self.t_acq += 1
        return self.buf[slot], ev

    def consumed(self, es):
        self.kernel_ev = torch.cuda.Event()
        self.kernel_ev.record(es)


def multitrack_crossfade_group_mc(engine, clip_layers, ring, width, height, blend_factor, out_palette=1):
    """multitrack_crossfade_group with the operand group taken from an OperandMulticast ring (the owner has published it): barrier,
    then ONE kernel launch for the K conversions + crossfades of this rank's clip."""
    from . import engine as E
    k = len(clip_layers)
    es = _engine_stream(engine)
    group, ready = ring.acquire()
    if group.dim() != 3 or group.shape[0] != k or group.shape[1] != height:
        raise ValueError("ring slots must be K x height x rowstride")
    es.wait_event(ready)
    rs = group.shape[2]
    ops = [E.Layer.wrap_device(engine, out_palette, width, height, [group[i].data_ptr()], [rs]) for i in range(k)]
    if E.convert_crossfade_batchv(clip_layers, ops, out_palette, 0, blend_factor) != k:
        raise RuntimeError("grouped crossfade failed: " + E.capi.last_error())
    ring.consumed(es)
    return clip_layers


def chain_schedule(step, rank, nslots, lag=1):
    """OperandChain's lockstep schedule (pure; tests/test_shard_cpu.py simulates it): at global step `step`, rank 0 stages group `step`
    into its slot, rank r >= 1 pulls group `step - lag * r` from rank r - 1 (which obtained it `lag` steps earlier); the group a rank
    obtains in a step is the group it consumes in that step.  Returns (group, slot) or (None, None) while the pipeline fills.
    A slot is rewritten nslots steps after it was filled and read by the successor `lag` steps after: nslots >= lag + 2 keeps a whole
    step (and its barrier) between the successor's read and the rewrite."""
    g = step - lag * rank
    if g < 0:
        return None, None
    return g, g % nslots


class OperandChain:
    """The shared operand of config 5 as a SYSTOLIC CHAIN on the copy engines: rank r pulls each operand group from rank r - 1's
    symmetric buffer with a peer-to-peer cudaMemcpyAsync (no SM runs a copy, the marching kernel keeps every SM), one hop per step,
    all hops at once -- every NVLink carries one group per step in one direction, so every GPU takes the operand in at the
    point-to-point rate (803 GB/s measured between two B200s, profiles/r02t_cfg5_transports.log) instead of the 530 - 640 GB/s of
    ncclBroadcast or the 119 GB/s per receiver of seven pulls from one owner.  Rank r runs lag * r groups behind rank 0: a render of G
    groups takes G + lag * (world - 1) steps; prime() runs the fill steps.

    Step T on every rank (stream ordered, no host synchronisation; chain_schedule() is the same rule as a pure function):
      copy stream    wait for this rank's kernel that read the slot's previous group (nslots steps ago) and for barrier T - lag
                     rank 0: stage group T from `src` into its slot; rank r: pull group T - lag * r from rank r - 1's slot into its own
                     record `arrived`
      barrier stream wait `arrived`, symmetric-memory barrier T (behind it every rank's step-T copy has landed)
      engine stream  wait `arrived`, ONE kernel for the K conversions + crossfades, record `consumed`.
    With lag = 1 the barrier of step T sits between the copies of steps T and T + 1 (every step pays it and the slowest rank's skew);
    with lag = 2 (the default) a copy waits for the barrier of TWO steps ago, which has long passed: the copies of consecutive steps
    run back to back."""

    def __init__(self, engine, slot_shape, nslots=4, lag=2, group=None):
        import torch.distributed._symmetric_memory as symm
        if lag < 1 or nslots < lag + 2:
            raise ValueError("the chain needs lag >= 1 and at least lag + 2 slots")
        self.engine, self.nslots, self.lag = engine, nslots, lag
        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        dev = torch.device("cuda", torch.cuda.current_device())
        self.shape = (nslots,) + tuple(slot_shape)
        self.buf = symm.empty(self.shape, dtype=torch.uint8, device=dev)
        self.hdl = symm.rendezvous(self.buf, self.group)
        self.prev = self.hdl.get_buffer(self.rank - 1, self.shape, torch.uint8) if self.rank > 0 else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.barrier_stream = torch.cuda.Stream(device=dev)
        self.kernel_ev = [None] * nslots   # this rank's kernel that read the slot last
        self.barrier_ev = {}               # step -> event behind that step's barrier
        self.t = 0
        self.pending_slot = None

    def step(self, src=None):
        """one lockstep step; returns (buffer of the group this rank consumes now, event behind which it is valid) or (None, None)
        while the pipeline fills.  `src`: on rank 0 the next group (contiguous uint8 of one slot's size)."""
        g, slot = chain_schedule(self.t, self.rank, self.nslots, self.lag)
        ev = None
        with torch.cuda.stream(self.copy_stream):
            need = self.barrier_ev.pop(self.t - self.lag, None)
            if need is not None:
                self.copy_stream.wait_event(need)
            if g is not None:
                if self.kernel_ev[slot] is not None:
                    self.copy_stream.wait_event(self.kernel_ev[slot])
                if self.rank == 0:
                    if src is None or src.numel() != self.buf[slot].numel():
                        raise ValueError("rank 0 stages one slot-sized group per step")
                    self.buf[slot].copy_(src.view(self.buf[slot].shape), non_blocking=True)
                else:
                    self.buf[slot].copy_(self.prev[slot], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        with torch.cuda.stream(self.barrier_stream):
            self.barrier_stream.wait_event(ev)
            self.hdl.barrier(channel=0)
            bev = torch.cuda.Event()
            bev.record(self.barrier_stream)
            self.barrier_ev[self.t] = bev
        self.t += 1
        self.pending_slot = slot
        return (self.buf[slot], ev) if g is not None else (None, None)

    def prime(self, src=None):
        """the lag * (world - 1) steps that fill the pipeline: afterwards every step() hands every rank a group"""
        for _ in range(self.lag * (self.world - 1)):
            self.step(src)

    def consumed(self, es):
        ev = torch.cuda.Event()
        ev.record(es)
        self.kernel_ev[self.pending_slot] = ev


def multitrack_crossfade_group_chain(engine, clip_layers, chain, src, width, height, blend_factor, out_palette=1):
    """multitrack_crossfade_group with the operand group taken from an OperandChain (primed): one chain step, then ONE kernel launch for
    the K conversions + crossfades of this rank's clip.  `src`: rank 0's next operand group (ignored elsewhere)."""
    from . import engine as E
    k = len(clip_layers)
    es = _engine_stream(engine)
    group, ready = chain.step(src)
    if group is None:
        raise RuntimeError("the chain is not primed")
    if group.dim() != 3 or group.shape[0] != k or group.shape[1] != height:
        raise ValueError("chain slots must be K x height x rowstride")
    es.wait_event(ready)
    rs = group.shape[2]
    ops = [E.Layer.wrap_device(engine, out_palette, width, height, [group[i].data_ptr()], [rs]) for i in range(k)]
    if E.convert_crossfade_batchv(clip_layers, ops, out_palette, 0, blend_factor) != k:
        raise RuntimeError("grouped crossfade failed: " + E.capi.last_error())
    chain.consumed(es)
    return clip_layers
