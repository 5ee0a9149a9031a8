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
    (torch's current stream), and the next broadcast into the same tensor waits for that kernel; with two operand buffers
    the broadcast of output frame t + 1 overlaps the kernel of frame t."""
    from . import engine as E
    es, ts = _engine_stream(engine), torch.cuda.current_stream()
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
    consumed = torch.cuda.Event()
    consumed.record(es)
    ts.wait_event(consumed)
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
    broadcast_operand(operand_group, src=src_rank)
    arrived = torch.cuda.Event()
    arrived.record(ts)
    es.wait_event(arrived)
    rs = operand_group.shape[2]
    ops = [E.Layer.wrap_device(engine, out_palette, width, height, [operand_group[i].data_ptr()], [rs]) for i in range(k)]
    if E.convert_crossfade_batchv(clip_layers, ops, out_palette, 0, blend_factor) != k:
        raise RuntimeError("grouped crossfade failed: " + E.capi.last_error())
    consumed = torch.cuda.Event()
    consumed.record(es)
    ts.wait_event(consumed)
    return clip_layers
