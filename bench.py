#!/usr/bin/env python
"""bench.py -- frames/s of the LiVES per-frame pixel path on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            our arm  (torchrun for N > 1, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K ...  the reference's CPU loops on this box's host cores

Workload (N = 1 and every N: weak scaling, frames are independent -- SURVEY.md 8e): the north-star headline row,
    3840x2160 YUV420P fg -> RGBA32 (reference converter arithmetic), letterboxed to a 3840x1608 inner rectangle in
    a 3840x2160 frame (vertical bilinear squeeze), scalar alpha-over (0.5) a 3840x2160 RGBA32 bg, gamma LUT8
    (linear -> sRGB) -- ONE fused kernel per batch of frames.
A "step" is one pass over a batch of --batch independent frames per GPU (all distinct, resident in HBM: the batch
is far larger than the 126 MB L2, so nothing is served from cache between steps).
  value     frames/s over all ranks, inputs resident in HBM, CUDA-event timed on the engine's stream (max over ranks)
  e2e       the same chain through the host-buffer C-ABI call (pe_host_fused_convert_letterbox_over_gamma): pinned host
            frames in, pinned host frame out, H2D + kernel + D2H all inside the timed region
  roofline  algorithmic bytes (78 796 800 B / frame: fg planes + bg read once, out written once) / kernel time,
            against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference CPU chain on a bounded sample of the same workload (rank 0, N = 1 only)
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

METRIC = "frames/sec 4K RGBA convert+resize+composite @1/2/4/8 B200; % HBM roofline"
FW, FH = 3840, 2160          # fg (YUV420P) and bg / out (RGBA32) size
IW, IH = 3840, 1608          # letterbox inner rectangle
ALPHA = 0.5
G_LINEAR, G_SRGB = -1, 1
FG_BYTES = FW * FH + 2 * (FW // 2) * (FH // 2)      # 12 441 600
RGBA_BYTES = FW * FH * 4                            # 33 177 600
ALGO_BYTES_PER_FRAME = FG_BYTES + 2 * RGBA_BYTES    # 78 796 800
WORKLOAD = "headline: 3840x2160 YUV420P fg -> RGBA32 + letterbox(3840x1608 in 3840x2160, bilinear) + alpha-over(0.5) RGBA32 bg + gamma LUT8, fused"


def config_dict(batch):
    """the same dict on both arms (the driver compares them): the CPU arm's own step size is in its cpu_baseline.sample"""
    return {"workload": WORKLOAD, "frames_per_step_per_gpu": batch, "parallelism": "frames sharded, no collective",
            "l2": "inputs larger than L2: %.0f MB touched per step vs 126 MB L2" % (ALGO_BYTES_PER_FRAME * batch / 1e6)}


def measured_peak():
    try:
        with open(os.path.join(REPO, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------ clocks

class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU reference chain

class CpuChain:
    """The reference's CPU implementation of the workload, one frame per call (test infrastructure: oracle/, tests/swscale_ref.py).

    The stock path for this chain (letterbox_layer -> resize_layer(tpal = RGBA32) -> resize_layer_full, src/colourspace.c:15343,
    :14759): a YUV420P layer with an RGBA32 target is converted AND scaled by ONE sws_scale call (get_resizable :14601-14620,
    flags SWS_BILINEAR :14997, sws_setColorspaceDetails :15079) -- convert_yuv420p_to_rgb_frame is not on this path.
    convert + resize  libswscale's sws_scale, YUV420P 3840x2160 -> RGBA 3840x1608 in one call (the library the reference links; the
                      one loadable here is the opencv wheel's, version in `sws_version`); context cached per thread as the reference
                      caches its own (sws_getCachedContext :15059)
    letterbox         the centred row copies of letterbox_layer (:15522-15549) into a black frame (oracle port of a memcpy loop)
    over              compiled reference compositor.c paint_pixel (oracle/_ref/ref_paint_pixel.so)
    gamma             compiled reference gamma_convert_layer_thread with the LUT of create_gamma_lut8
    stages = "sws":     the above (kind "reference")
    stages = "loops":   convert = compiled reference convert_yuv420p_to_rgb_frame, resize = oracle port (round 1's arm; what the
                        reference would run with libswscale compiled out; kind "reference" / "port")
    Without oracle/_ref every compiled stage falls back to the oracle port (kind "port")."""

    def __init__(self, stages="auto"):
        sys.path.insert(0, os.path.join(REPO, "tests"))
        import pe_testlib as T
        import swscale_ref as S
        self.T, self.S = T, S
        self.o = T.oracle()
        self.have_ref = T.have_ref()
        sws, ver = S.load()
        self.sws_version = ver if sws is not None else None
        if stages == "auto":
            stages = "sws" if sws is not None else "loops"
        if stages == "sws" and sws is None:
            raise SystemExit("bench.py: no loadable libswscale (%s)" % ver)
        self.stages = stages
        self.kind = "reference" if self.have_ref else "port"
        if self.have_ref:
            self.r = T.ref()
            self.r.ref_set_prefs(1, T.Q_HIGH, 1.4)  # frames are spread over the cores, one band per frame
            self.p = T.ref_paint()
        self.lut = np.zeros(256, np.uint8)
        assert self.o.pe_or_gamma_lut8(1.0, G_LINEAR, G_SRGB, 1.4, T.ptr(self.lut)) == 0
        self.tls = threading.local()

    def describe(self):
        if self.stages == "sws":
            return ("convert + resize = ONE sws_scale call YUV420P 3840x2160 -> RGBA 3840x1608 (libswscale %s, SWS_BILINEAR, as "
                    "resize_layer_full issues it, colourspace.c:14601-14620); letterbox = row copies; alpha-over / gamma = %s"
                    % (self.sws_version, "compiled reference loops" if self.have_ref else "oracle port"))
        return ("convert / alpha-over / gamma = %s, resize + letterbox = oracle port (the reference's path with libswscale compiled out)"
                % ("compiled reference loops" if self.have_ref else "oracle port"))

    def _bufs(self):
        t = self.tls
        if not hasattr(t, "inner"):
            t.inner = np.zeros((IH, IW * 4), np.uint8)
            t.boxed = np.zeros((FH, FW * 4), np.uint8)
            if self.stages == "sws":
                t.scaler = self.S.Scaler("yuv420p", FW, FH, "rgba", IW, IH, self.S.SWS_BILINEAR, yuv=(False, False, False))
            else:
                t.rgba_full = np.zeros((FH + 16, FW * 4), np.uint8)  # slack rows: the reference has stray writes (:3584)
                t.rgba = t.rgba_full[8:8 + FH]
        return t

    def stage_times(self, y, u, v, bg, out):
        """one frame, per-stage wall clock (single thread)"""
        T, t = self.T, self._bufs()
        times = {}
        t0 = time.perf_counter()
        self._convert_resize(t, y, u, v)
        times["convert+resize"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.o.pe_or_letterbox_packed(T.ptr(t.inner), IW * 4, IW, IH, T.ptr(t.boxed), FW * 4, FW, FH, 3)
        times["letterbox"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        np.copyto(out, bg)
        self._over(t, out)
        times["alpha-over"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        self._gamma(out)
        times["gamma"] = time.perf_counter() - t0
        return times

    def _convert_resize(self, t, y, u, v):
        T = self.T
        if self.stages == "sws":
            t.scaler.run([y, u, v], t.inner)
            return
        pl, st = T.planes_arg(y, u, v), T.strides_arg(y, u, v)
        if self.have_ref:
            self.r.ref_yuv420p_to_rgb(pl, FW, FH, st, FW * 4, T.ptr(t.rgba), 0, 1, 0, 0, 0, 1, 0, 0)
        else:
            self.o.pe_or_yuv420p_to_rgb(pl, st, FW, FH, T.ptr(t.rgba), FW * 4, 0, 1, 0, 0, 1, T.Q_HIGH, 1, None)
        self.o.pe_or_resize_packed(T.ptr(t.rgba), FW * 4, FW, FH, T.ptr(t.inner), IW * 4, IW, IH, 4)

    def _over(self, t, out):
        T = self.T
        if self.have_ref:
            self.p.ref_paint_rows(T.ptr(out), T.ptr(t.boxed), FW * FH, 4, ALPHA)
        else:
            self.o.pe_or_alpha_over(T.ptr(out), FW * 4, T.ptr(t.boxed), FW * 4, 3, FW, FH, ALPHA)
        out[:, 3::4] = 255

    def _gamma(self, out):
        T = self.T
        if self.have_ref:
            self.r.ref_gamma_apply(T.ptr(out), FW * 4, 4, 0, FW, FH, 0, T.ptr(self.lut))
        else:
            self.o.pe_or_gamma_apply(T.ptr(out), FW * 4, 3, 0, 0, FW, FH, T.ptr(self.lut))

    def frame(self, y, u, v, bg, out):
        T, t = self.T, self._bufs()
        self._convert_resize(t, y, u, v)
        self.o.pe_or_letterbox_packed(T.ptr(t.inner), IW * 4, IW, IH, T.ptr(t.boxed), FW * 4, FW, FH, 3)
        np.copyto(out, bg)
        self._over(t, out)
        self._gamma(out)


def host_frames(n, seed0=20):
    """synthetic frames exactly as SURVEY.md 8d: Y in [16,235], U,V in [16,240], bg uniform u8"""
    frames = []

    def plane(rng, rows, cols, lo, hi):
        # + guard bytes: the reference converter reads one byte past the last chroma row (colourspace.c:3508)
        buf = np.zeros(rows * cols + 16, np.uint8)
        pl = buf[:rows * cols].reshape(rows, cols)
        pl[...] = rng.integers(lo, hi, (rows, cols), dtype=np.uint8)
        buf[rows * cols] = pl[-1, -1]
        return pl

    for i in range(n):
        rng = np.random.default_rng(seed0 + 2 * i)
        y = plane(rng, FH, FW, 16, 236)
        u = plane(rng, FH // 2, FW // 2, 16, 241)
        v = plane(rng, FH // 2, FW // 2, 16, 241)
        bg = np.random.default_rng(seed0 + 2 * i + 1).integers(0, 256, (FH, FW * 4), dtype=np.uint8)
        frames.append((y, u, v, bg))
    return frames


class CpuBench:
    """n_frames independent frames of the CPU chain spread over `cores` threads (ctypes drops the GIL)"""

    def __init__(self, n_frames, cores, stages="auto"):
        from concurrent.futures import ThreadPoolExecutor
        self.chain = CpuChain(stages)
        self.n, self.cores = n_frames, cores
        self.distinct = host_frames(2)
        self.outs = [np.zeros((FH, FW * 4), np.uint8) for _ in range(min(cores, n_frames))]
        self.ex = ThreadPoolExecutor(max_workers=cores)
        list(self.ex.map(self._work, range(min(cores, n_frames))))  # warm-up: page in buffers, build tables

    def _work(self, i):
        y, u, v, bg = self.distinct[i % len(self.distinct)]
        self.chain.frame(y, u, v, bg, self.outs[i % len(self.outs)])

    def step(self):
        t0 = time.perf_counter()
        list(self.ex.map(self._work, range(self.n)))
        return time.perf_counter() - t0

    def close(self):
        self.ex.shutdown()


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ------------------------------------------------------------------------------------------------ arms

def cpu_single_thread(chain):
    """nfx_threads = 1: one frame at a time on one host thread, per-stage wall clock (median of 3 frames after one warm-up)"""
    fr = host_frames(2)
    out = np.zeros((FH, FW * 4), np.uint8)
    runs = []
    for i in range(4):
        y, u, v, bg = fr[i % 2]
        runs.append(chain.stage_times(y, u, v, bg, out))
    stages = {k: float(np.median([r[k] for r in runs[1:]])) * 1e3 for k in runs[0]}
    total = sum(stages.values())
    return {"fps": 1000.0 / total, "ms_per_frame": total, "stage_ms": stages}


def run_reference(args, rank, world):
    if rank != 0:
        return
    cores = host_cores()
    n_frames = cores  # one frame per host thread and step
    cb = CpuBench(n_frames, cores, args.cpu_stages)
    chain = cb.chain
    fps_steps = []
    for s in range(args.warmup + args.steps):
        dt = cb.step()
        if s >= args.warmup:
            fps_steps.append((n_frames / dt, dt))
    single = cpu_single_thread(chain)
    cb.close()
    total_frames = n_frames * len(fps_steps)
    total_t = sum(dt for _, dt in fps_steps)
    value = total_frames / total_t
    sample = "%d frames/step of the headline workload over %d host threads (one frame per thread); %s" % (n_frames, cores, chain.describe())
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total_t / max(len(fps_steps), 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_dict(args.batch),
            "cpu_baseline": {"value": value, "unit": "frames/s", "cores": cores, "kind": chain.kind, "sample": sample,
                             "stages": chain.stages, "libswscale": chain.sws_version,
                             "nfx_threads_1": single, "nfx_threads_nproc": {"fps": value, "threads": cores}},
            "e2e": {"value": value, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_ours(args, rank, world, local_rank):
    import torch
    import lives_b200 as lb

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the pixel engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=dev)

    eng = lb.Engine(device=local_rank)
    B = args.batch
    g = torch.Generator(device=dev)
    g.manual_seed(20 + rank)
    keep, fgs, bgs, outs = [], [], [], []
    for i in range(B):
        y = torch.randint(16, 236, (FH, FW), dtype=torch.uint8, device=dev, generator=g)
        u = torch.randint(16, 241, (FH // 2, FW // 2), dtype=torch.uint8, device=dev, generator=g)
        v = torch.randint(16, 241, (FH // 2, FW // 2), dtype=torch.uint8, device=dev, generator=g)
        bg = torch.randint(0, 256, (FH, FW * 4), dtype=torch.uint8, device=dev, generator=g)
        out = torch.empty((FH, FW * 4), dtype=torch.uint8, device=dev)
        keep += [y, u, v, bg, out]
        fgs.append(lb.Layer.wrap_device(eng, lb.WEED_PALETTE_YUV420P, FW, FH, [y.data_ptr(), u.data_ptr(), v.data_ptr()],
                                        [FW, FW // 2, FW // 2], yuv_clamping=0, yuv_subspace=1))
        bgs.append(lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [bg.data_ptr()], [FW * 4], gamma_type=G_LINEAR))
        outs.append(lb.Layer.wrap_device(eng, lb.WEED_PALETTE_RGBA32, FW, FH, [out.data_ptr()], [FW * 4]))
    torch.cuda.synchronize()

    def step():
        lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, IW, IH, ALPHA, G_LINEAR, G_SRGB)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    launches0 = eng.launch_count
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop_ms()
    barrier()
    launches = eng.launch_count - launches0
    clk = clocks.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    frames_total = B * args.steps * world
    value = frames_total / (ms / 1000.0)
    kernel_ms = ms / max(launches, 1) if launches else ms / args.steps
    peak, peak_src = measured_peak()
    achieved = ALGO_BYTES_PER_FRAME * B / (kernel_ms / 1000.0) / 1e9

    # ---- single-frame latency: one frame per launch, device resident (SURVEY.md 8d asks for it beside the batch figure)
    for _ in range(3):
        lb.fused_convert_letterbox_over_gamma_batch(fgs[:1], bgs[:1], outs[:1], IW, IH, ALPHA, G_LINEAR, G_SRGB)
    eng.sync()
    eng.timer_start()
    for i in range(20):
        j = i % B
        lb.fused_convert_letterbox_over_gamma_batch(fgs[j:j + 1], bgs[j:j + 1], outs[j:j + 1], IW, IH, ALPHA, G_LINEAR, G_SRGB)
    single_frame_us = eng.timer_stop_ms() * 1000.0 / 20

    # ---- e2e: host buffers through the C-ABI drop-in, H2D + kernel + D2H timed (per rank, frames independent)
    hframes = host_frames(4, seed0=20 + 100 * rank)
    pinned = []
    def pinned_block(nbytes):
        p = lb._capi.lib().pe_host_alloc(nbytes)
        if not p:
            raise SystemExit("pinned host allocation failed")
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    for (y, u, v, bg) in hframes:
        # the three planes of the planar frame sit back to back in one allocation, the way LiVES allocates planar pixel data
        # (create_empty_pixel_data, WEED_LEAF_HOST_PIXEL_DATA_CONTIGUOUS): the drop-in then moves them in one copy
        blk, off, bufs = pinned_block(y.nbytes + u.nbytes + v.nbytes), 0, []
        for a in (y, u, v):
            arr = blk[off:off + a.nbytes].reshape(a.shape)
            arr[...] = a
            bufs.append(arr)
            off += a.nbytes
        for a in (bg, np.zeros((FH, FW * 4), np.uint8)):
            arr = pinned_block(a.nbytes).reshape(a.shape)
            arr[...] = a
            bufs.append(arr)
        pinned.append(bufs)
    hl = [(lb.HostLayer(lb.WEED_PALETTE_YUV420P, FW, FH, b[:3], yuv_subspace=1),
           lb.HostLayer(lb.WEED_PALETTE_RGBA32, FW, FH, [b[3]], gamma_type=G_LINEAR),
           lb.HostLayer(lb.WEED_PALETTE_RGBA32, FW, FH, [b[4]])) for b in pinned]
    e2e_frames = args.e2e_frames
    fg_l = [hl[i % len(hl)][0] for i in range(e2e_frames)]
    bg_l = [hl[i % len(hl)][1] for i in range(e2e_frames)]
    out_l = [hl[i % len(hl)][2] for i in range(e2e_frames)]

    def e2e_pass():
        # one call = a batch of host frames; H2D, kernel and D2H of consecutive frames overlap inside the call
        lb.host_fused_convert_letterbox_over_gamma_batch(eng, fg_l, bg_l, out_l, IW, IH, ALPHA, G_LINEAR, G_SRGB)

    e2e_pass()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_pass()
    eng.sync()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_value = e2e_frames * args.e2e_steps * world / e2e_s
    checksum = int(pinned[0][4][::97, ::101].astype(np.uint64).sum())  # the D2H result is really read

    # ---- the same call on PAGEABLE host buffers (what LiVES' lives_calloc_safety hands out; the `e2e` above is the case of a host
    #      that page-locked its pixel blocks once, pe_host_register): plain numpy allocations, never registered, a few frames
    e2e_pageable = None
    if rank == 0 or dist is not None:
        try:
            pg = [(lb.HostLayer(lb.WEED_PALETTE_YUV420P, FW, FH, [y, u, v], yuv_subspace=1),
                   lb.HostLayer(lb.WEED_PALETTE_RGBA32, FW, FH, [bg], gamma_type=G_LINEAR),
                   lb.HostLayer(lb.WEED_PALETTE_RGBA32, FW, FH, [np.full((FH, FW * 4), 1, np.uint8)])) for (y, u, v, bg) in hframes]   # (np.full: pages touched)
            npg = min(8, max(4, e2e_frames // 4))
            pf, pb, po = [pg[i % 4][0] for i in range(npg)], [pg[i % 4][1] for i in range(npg)], [pg[i % 4][2] for i in range(npg)]
            lb.host_fused_convert_letterbox_over_gamma_batch(eng, pf[:4], pb[:4], po[:4], IW, IH, ALPHA, G_LINEAR, G_SRGB)
            eng.sync()
            t0 = time.perf_counter()
            lb.host_fused_convert_letterbox_over_gamma_batch(eng, pf, pb, po, IW, IH, ALPHA, G_LINEAR, G_SRGB)
            eng.sync()
            e2e_pageable = {"value": npg / (time.perf_counter() - t0), "unit": "frames/s per GPU", "frames": npg,
                            "note": "pe_host_fused_convert_letterbox_over_gamma_batch on pageable (never registered) host buffers, this rank alone"}
        except Exception as ex:  # noqa: BLE001
            e2e_pageable = {"value": None, "note": "failed: %s" % str(ex).splitlines()[0][:100]}

    # ---- e2e with device-resident ingest: fg and bg frames come from clips that already sit in HBM (the decoder
    #      plugin's get_frame shape over a device clip: zero H2D), the chain runs fused, and ONLY the final RGB24 frame (the render
    #      tail's layer_to_pixbuf, src/events.c:4263) crosses PCIe, on its own stream, four frames in flight
    fg_clip = lb.ClipCache(eng, lb.WEED_PALETTE_YUV420P, FW, FH, 4, yuv_subspace=1)
    bg_clip = lb.ClipCache(eng, lb.WEED_PALETTE_RGBA32, FW, FH, 4, gamma_type=G_LINEAR)
    for i, (y, u, v, bg) in enumerate(hframes):
        fg_clip.load(i, [y, u, v])
        bg_clip.load(i, [bg])
    rgb_host = [pinned_block(FH * FW * 3).reshape(FH, FW * 3) for _ in range(4)]
    ing_out = [lb.Layer.create(eng, lb.WEED_PALETTE_RGBA32, FW, FH) for _ in range(4)]

    def ingest_pass(nframes):
        for i in range(nframes):
            k = i & 3
            lb.render_out_wait(eng, k)  # the slot's previous download has left ing_out[k]
            if ing_out[k].palette != lb.WEED_PALETTE_RGBA32:
                ing_out[k].free()
                ing_out[k] = lb.Layer.create(eng, lb.WEED_PALETTE_RGBA32, FW, FH)
            fg, bgl = fg_clip.borrow(i), bg_clip.borrow(i)
            lb.fused_convert_letterbox_over_gamma(fg, bgl, ing_out[k], IW, IH, ALPHA, G_LINEAR, G_SRGB)
            lb.render_out_begin(ing_out[k], lb.WEED_PALETTE_RGB24, rgb_host[k], k)
            fg.free()
            bgl.free()
        for k in range(4):
            lb.render_out_wait(eng, k)

    ingest_pass(8)
    barrier()
    ing_frames = args.e2e_frames * args.e2e_steps
    t0 = time.perf_counter()
    ingest_pass(ing_frames)
    ing_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([ing_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ing_s = float(t.item())
    e2e_ingest = {"value": ing_frames * world / ing_s, "unit": "frames/s", "h2d_bytes_per_frame": 0, "d2h_bytes_per_frame": FW * FH * 3,
                  "api": "pe_clip_cache_borrow (device-resident clips) -> pe_fused_convert_letterbox_over_gamma -> pe_render_out_begin / _wait "
                         "(RGBA32 -> RGB24 on the device, only that frame downloaded)",
                  "checksum": int(rgb_host[0][::97, ::101].astype(np.uint64).sum())}
    fg_clip.close()
    bg_clip.close()

    # ---- the link alone, all ranks copying at once (what the e2e figure is bounded by on this box)
    probe = pcie_probe(dev)
    if dist is not None:
        t = torch.tensor([probe["frames_per_s"]], dtype=torch.float64, device=dev)
        tmin, tsum = t.clone(), t.clone()
        dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        probe = {"frames_per_s_all_ranks": float(tsum.item()), "frames_per_s_slowest_rank": float(tmin.item()), "rank0": probe}
    else:
        probe = {"frames_per_s_all_ranks": probe["frames_per_s"], "frames_per_s_slowest_rank": probe["frames_per_s"], "rank0": probe}
    subs = None if args.no_sub_records else sub_records(args, eng, dist, rank, world, dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        cb = CpuBench(cores, cores, args.cpu_stages)
        chain, n, dt = cb.chain, 0, 0.0
        while dt < 10.0 and n < 4096 * cores:  # bounded sample: >= 10 s of wall clock over all host threads
            dt += cb.step()
            n += cores
        single = cpu_single_thread(chain)
        cb.close()
        fps = n / dt
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": chain.kind, "stages": chain.stages, "libswscale": chain.sws_version,
               "sample": "%d frames of the same workload over %d host threads in %.1f s; %s" % (n, cores, dt, chain.describe()),
               "nfx_threads_1": single, "nfx_threads_nproc": {"fps": fps, "threads": cores}}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic",
                "config": config_dict(B),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": None, "peak_source": peak_src, "kernel": "k_fused3", "kernel_ms": kernel_ms,
                             "algorithmic_bytes_per_launch": ALGO_BYTES_PER_FRAME * B, "single_frame_launch_us": single_frame_us},
                "cpu_baseline": cpu,
                "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": (FG_BYTES + RGBA_BYTES) * e2e_frames,
                        "d2h_bytes_per_step": RGBA_BYTES * e2e_frames, "frames_per_step": e2e_frames, "steps": args.e2e_steps,
                        "api": "pe_host_fused_convert_letterbox_over_gamma_batch (pinned host frames in / out; H2D, kernel, D2H "
                               "of consecutive frames overlapped on three streams)", "checksum": checksum},
                "e2e_device_ingest": e2e_ingest, "e2e_pageable": e2e_pageable,
                "gpu_launches": int(launches), "clocks": clk, "configs": subs, "pcie_ceiling": probe}
        prof = os.path.join(REPO, "profiles", "traffic.json")
        if os.path.exists(prof):
            try:
                # dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, per frame
                per_frame = json.load(open(prof)).get("k_fused3_dram_bytes_per_frame")
                line["roofline"]["traffic"] = per_frame * B if per_frame else None
            except Exception:
                pass
        print(json.dumps(line), flush=True)
    for b in pinned:
        for a in b:
            lb._capi.lib().pe_host_free(a.ctypes.data)
    if dist is not None:
        dist.destroy_process_group()


def sub_records(args, eng, dist, rank, world, dev):
    """BASELINE configs 4 and 5 at N GPUs, inside the driver's line (the one collective of the path had no driver-side
    measurement).  Device-resident, CUDA-event timed on the engine stream, max over ranks; weak scaling.
      cfg4  32 x 1080p RGB24 'chroma blend' bf = 100 per GPU and step (256 frames over 8 GPUs), one launch, no collective
      cfg5  one clip per GPU: K = 8 consecutive 4K YUV422P frames -> RGB24 + crossfade (chroma blend bf = 128) with K frames of the
            shared operand, which rank 0 owns and every step BROADCASTS over NCCL (199 MB per step); three operand group buffers so
            that two broadcasts can run ahead of the kernel.  parity_all_ranks: every rank also crossfades against a locally generated
            copy of the operand (same seed everywhere) and compares the results bit for bit."""
    import torch
    import lives_b200 as lb
    from lives_b200 import shard
    out = {}
    g = torch.Generator(device=dev)
    g.manual_seed(1000 + rank)

    def timed(step, steps, warm=3):
        for _ in range(warm):
            step()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()
        l0 = eng.launch_count
        eng.timer_start()
        for _ in range(steps):
            step()
        ms = eng.timer_stop_ms()
        torch.cuda.synchronize()
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, eng.launch_count - l0

    peak, _ = measured_peak()
    # ---- cfg4
    W, H, n = 1920, 1080, 32
    keep = [torch.randint(0, 256, (3, n, H, W * 3), dtype=torch.uint8, device=dev, generator=g)]
    a = [lb.Layer.wrap_device(eng, 1, W, H, [keep[0][0, i].data_ptr()], [W * 3]) for i in range(n)]
    b = [lb.Layer.wrap_device(eng, 1, W, H, [keep[0][1, i].data_ptr()], [W * 3]) for i in range(n)]
    o = [lb.Layer.wrap_device(eng, 1, W, H, [keep[0][2, i].data_ptr()], [W * 3]) for i in range(n)]
    steps4 = 20
    ms, launches = timed(lambda: lb.simple_blend_batch("chroma blend", a, b, o, 100), steps4)
    fps = world * n * steps4 / (ms / 1e3)
    algo4 = 3 * W * H * 3
    out["cfg4"] = {"workload": "256 x 1080p RGB24 'chroma blend' bf=100 over 8 GPUs = 32 frames per GPU and step, one launch, no collective",
                   "value": fps, "unit": "frames/s", "ms_per_step": ms / steps4, "gpu_launches": int(launches),
                   "roofline_frac_per_gpu": algo4 * n * steps4 / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_frame": algo4}
    del a, b, o, keep
    # ---- cfg5
    W, H, K = 3840, 2160, 8
    yy = torch.randint(16, 236, (H, W), dtype=torch.uint8, device=dev, generator=g)
    uu = torch.randint(16, 241, (H, W // 2), dtype=torch.uint8, device=dev, generator=g)
    vv = torch.randint(16, 241, (H, W // 2), dtype=torch.uint8, device=dev, generator=g)
    og = torch.Generator(device=dev)
    og.manual_seed(99)  # the SAME operand frames on every rank: rank 0's copy travels, the others are the parity reference
    local_ops = torch.randint(0, 256, (K, H, W * 3), dtype=torch.uint8, device=dev, generator=og)
    # transport of the operand: an NVSwitch multicast ring (shard.OperandMulticast: the owner's pe_mc_publish kernel stores each group
    # into every rank's buffer at once) when the group has multicast support, NCCL broadcast into three group buffers otherwise
    ring, chain, transport = None, None, "none (N = 1)"
    sel = os.environ.get("PE_CFG5_TRANSPORT", "chain")
    if world > 1:
        ok, why = 1, ""
        try:
            if sel == "multicast":
                ring = shard.OperandMulticast(eng, (K, H, W * 3), nslots=3)
            elif sel == "chain":
                lag = int(os.environ.get("PE_CFG5_CHAIN_LAG", "2"))
                chain = shard.OperandChain(eng, (K, H, W * 3), nslots=max(4, lag + 2), lag=lag)
            else:
                raise RuntimeError("PE_CFG5_TRANSPORT=%s" % sel)
        except Exception as ex:  # noqa: BLE001
            ok, why = 0, str(ex).splitlines()[0][:80]
        okt = torch.tensor([ok], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        if int(okt.item()) and ring is not None:
            transport = "NVSwitch multicast: one pe_mc_publish kernel on rank 0 stores each group into all ranks' symmetric buffers (3 slots, 2 groups ahead)"
        elif int(okt.item()) and chain is not None:
            transport = ("systolic chain on the copy engines: rank r pulls each group from rank r - 1's symmetric buffer (peer-to-peer cudaMemcpyAsync, one hop "
                         "per step, all hops at once, %d slots; rank r runs %d r groups behind rank 0); no collective library and no SM on the data path" % (chain.nslots, chain.lag))
        else:
            ring = chain = None
            transport = "broadcast from rank 0 over NCCL once per step (3 group buffers, per-buffer ordering)" + ("" if ok else "; " + sel + ": " + why)
    # the collective's CTAs need somewhere to run: the marching kernel is persistent (one CTA per SM), so it leaves some SMs free
    # (the chain moves the operand with copy engines: nothing to reserve)
    sm_reserve = int(os.environ.get("PE_CFG5_SM_RESERVE", "0" if chain is not None else "8")) if world > 1 else 0
    if sm_reserve > 0:
        eng.set_sm_limit(eng.sm_count - sm_reserve)
    groups = [local_ops.clone() if rank == 0 else torch.zeros_like(local_ops) for _ in range(3)] if ring is None and chain is None else []
    it = [0]
    if ring is not None:
        ring.publish(local_ops)
        ring.publish(local_ops)
    if chain is not None:
        chain.prime(local_ops if rank == 0 else None)   # world - 1 transfer-only steps: from here on every step hands every rank a group

    def clips():
        return [lb.Layer.wrap_device(eng, lb.WEED_PALETTE_YUV422P, W, H, [yy.data_ptr(), uu.data_ptr(), vv.data_ptr()], [W, W // 2, W // 2],
                                     yuv_subspace=1) for _ in range(K)]

    def step5(keep_result=False):
        cl = clips()
        if ring is not None:
            shard.multitrack_crossfade_group_mc(eng, cl, ring, W, H, 128)
            ring.publish(local_ops)   # group t + 2, behind barrier t: it travels while the kernels of groups t and t + 1 run
        elif chain is not None:
            shard.multitrack_crossfade_group_chain(eng, cl, chain, local_ops if rank == 0 else None, W, H, 128)
        else:
            shard.multitrack_crossfade_group(eng, cl, groups[it[0] % 3], W, H, 128)
        it[0] += 1
        if keep_result:
            return cl
        for c in cl:
            c.free()

    # parity: broadcast path vs the local copy of the operand
    got = step5(keep_result=True)
    ref = clips()
    ops = [lb.Layer.wrap_device(eng, 1, W, H, [local_ops[i].data_ptr()], [W * 3]) for i in range(K)]
    assert lb.convert_crossfade_batchv(ref, ops, 1, 0, 128) == K
    eng.sync()
    torch.cuda.synchronize()
    same = True
    for x, y_ in zip(got, ref):
        dx, dy = x.desc, y_.desc
        hx, hy = x.to_host()[0], y_.to_host()[0]
        same = same and bool((hx == hy).all()) and dx.palette == 1 and dy.palette == 1
    for c in got + ref:
        c.free()
    flag = torch.tensor([int(same)], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    steps5 = 10
    ms, launches = timed(step5, steps5, warm=2)
    cps = world * K * steps5 / (ms / 1e3)
    algo5 = W * H * 2 + 2 * W * H * 3
    eng.set_sm_limit(0)
    out["cfg5"] = {"workload": "multitrack: one clip per GPU, %d x 4K YUV422P -> RGB24 + crossfade bf=128 per step against %d frames of the shared "
                               "operand, %s" % (K, K, (transport + "; %d SMs left to the transfer" % sm_reserve) if world > 1 else "no broadcast at N = 1"),
                   "value": cps, "unit": "clip frames/s", "ms_per_output_frame": ms / (steps5 * K), "gpu_launches": int(launches),
                   "parity_all_ranks": bool(flag.item()), "broadcast_bytes_per_step": (K * H * W * 3) if world > 1 else 0,
                   "broadcast_gbs": (K * H * W * 3 * steps5 / (ms / 1e3) / 1e9) if world > 1 else None,
                   "roofline_frac_per_gpu": algo5 * K * steps5 / (ms / 1e3) / 1e9 / peak, "algorithmic_bytes_per_frame": algo5}
    return out


def pcie_probe(dev):
    """this rank's pinned-copy ceiling for the e2e path (45.6 MB up + 33.2 MB down per frame, both directions busy), so that an e2e
    figure that does not scale can be told apart from the box: frames/s the link alone would allow"""
    import torch
    UP, DOWN, N = FG_BYTES + RGBA_BYTES, RGBA_BYTES, 12
    hu, hd = torch.empty(UP, dtype=torch.uint8).pin_memory(), torch.empty(DOWN, dtype=torch.uint8).pin_memory()
    du, dd = torch.empty(UP, dtype=torch.uint8, device=dev), torch.empty(DOWN, dtype=torch.uint8, device=dev)
    s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def run():
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        s_up.wait_stream(torch.cuda.current_stream())
        s_dn.wait_stream(torch.cuda.current_stream())
        for _ in range(N):
            with torch.cuda.stream(s_up):
                du.copy_(hu, non_blocking=True)
            with torch.cuda.stream(s_dn):
                hd.copy_(dd, non_blocking=True)
        torch.cuda.current_stream().wait_stream(s_up)
        torch.cuda.current_stream().wait_stream(s_dn)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 1e3
    run()
    t = min(run() for _ in range(2))
    return {"frames_per_s": N / t, "h2d_gbs": UP * N / t / 1e9, "d2h_gbs": DOWN * N / t / 1e9}


def run_secondary(args):
    """Kernel-level numbers of the other BASELINE configs (N = 1, inputs resident; not the driver's headline line)."""
    import torch
    import lives_b200 as lb
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    eng = lb.Engine(device=0)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    keep = []

    def rnd(h, nbytes, lo=0, hi=256):
        t = torch.randint(lo, hi, (h, nbytes), dtype=torch.uint8, device=dev, generator=g)
        keep.append(t)
        return t

    def wrap(pal, w, h, tensors, **kw):
        return lb.Layer.wrap_device(eng, pal, w, h, [t.data_ptr() for t in tensors], [t.shape[1] for t in tensors], **kw)

    wl = args.workload
    prepare = None
    if wl == "cfg4":    # batch of 1080p RGB24 chroma blends, one launch
        W, H, n = 1920, 1080, 64
        a = [wrap(1, W, H, [rnd(H, W * 3)]) for _ in range(n)]
        b = [wrap(1, W, H, [rnd(H, W * 3)]) for _ in range(n)]
        o = [wrap(1, W, H, [rnd(H, W * 3)]) for _ in range(n)]
        step = lambda: lb.simple_blend_batch("chroma blend", a, b, o, 100)
        frames, algo, name = n, 3 * W * H * 3, "cfg4: %d x 1080p RGB24 chroma blend bf=100, one launch" % n
    elif wl == "cfg3":  # 4K RGBA alpha-over + gamma (unfused ops: fill + 2 x alpha_over + lut)
        W, H, n = 3840, 2160, 8
        bg = [wrap(3, W, H, [rnd(H, W * 4)]) for _ in range(n)]
        fg = [wrap(3, W, H, [rnd(H, W * 4)]) for _ in range(n)]
        out = [wrap(3, W, H, [rnd(H, W * 4)], gamma_type=G_LINEAR) for _ in range(n)]

        per_frame = os.environ.get("PE_CFG3_PER_FRAME") is not None

        def step():
            for i in range(n):
                out[i].gamma_type = G_LINEAR
            if per_frame:
                for i in range(n):
                    lb.compositor_gamma(out[i], [fg[i], bg[i]], [0.5, 1.0], G_SRGB)
            else:
                assert lb.compositor_gamma_batch(out, [[fg[i], bg[i]] for i in range(n)], [0.5, 1.0], G_SRGB) == n
        frames, algo, name = n, 3 * W * H * 4, "cfg3: %d x 4K RGBA32 alpha-over(0.5) + gamma LUT8 (%s)" % (
            n, "pe_fx_compositor_gamma, one kernel / frame" if per_frame else "pe_fx_compositor_gamma_batch, one kernel / %d frames" % n)
    elif wl == "cfg2":  # 1080p YUV420P -> RGBA32 -> 1280x720 (unfused convert + resize)
        W, H, n = 1920, 1080, 16
        src = [(rnd(H, W, 16, 236), rnd(H // 2, W // 2, 16, 241), rnd(H // 2, W // 2, 16, 241)) for _ in range(n)]

        pool = []

        def prepare(nsteps):
            pool.extend([wrap(512, W, H, [y, u, v], yuv_subspace=1) for y, u, v in src] for _ in range(nsteps))

        def step():
            lays = pool.pop()
            assert lb.resize_layer_batch(lays, 1280, 720, 1, 3, 0) == n
            for lay in lays:
                lay.free()
        frames, algo, name = n, W * H * 3 // 2 + 1280 * 720 * 4, "cfg2: %d x 1080p YUV420P -> RGBA32 -> 1280x720, pe_resize_layer_batch (convert + tiled resize per frame)" % n
    elif wl == "cfg5":  # per clip: 4K YUV422P -> RGB24 + crossfade (chroma blend bf=128) with the shared operand; no broadcast at N = 1
        W, H, n = 3840, 2160, 8
        src = [(rnd(H, W, 16, 236), rnd(H, W // 2, 16, 241), rnd(H, W // 2, 16, 241)) for _ in range(n)]
        operand = wrap(1, W, H, [rnd(H, W * 3)])

        pool = []  # layers wrapped ahead of the timed region: the host only issues the fused call per clip

        def prepare(nsteps):
            pool.extend(wrap(522, W, H, [y, u, v], yuv_subspace=1) for _ in range(nsteps) for (y, u, v) in src)

        per_clip = os.environ.get("PE_CFG5_PER_CLIP") is not None  # the one-kernel-per-clip form (what one rank of the N-GPU run issues)

        def step():
            lays = [pool.pop() for _ in range(n)]
            if per_clip:
                for lay in lays:
                    lb.convert_crossfade(lay, operand, 1, 0, 128)
            else:
                assert lb.convert_crossfade_batch(lays, operand, 1, 0, 128) == n
            for lay in lays:
                lay.free()  # stream ordered: the converted frame goes back to the pool behind the kernel
        frames, algo, name = n, W * H * 2 + 2 * W * H * 3, "cfg5 (1 GPU, no broadcast): %d x 4K YUV422P -> RGB24 + chroma blend bf=128 with one shared operand (%s)" % (
            n, "pe_fx_convert_crossfade, 1 kernel / clip" if per_clip else "pe_fx_convert_crossfade_batch, 1 kernel / %d clips" % n)
    elif wl == "cfg1":  # 640x480 RGB24 -> BGR24 in place
        W, H, n = 640, 480, 256
        lay = [wrap(1, W, H, [rnd(H, W * 3)]) for _ in range(n)]

        def step():
            assert lb.convert_layer_palette_batch(lay, 2 if lay[0].palette == 1 else 1, 0) == n
        frames, algo, name = n, 2 * W * H * 3, "cfg1: %d x 640x480 RGB24 <-> BGR24 in place (pe_convert_layer_palette_batch, one launch per frame)" % n
    else:
        raise SystemExit("unknown workload " + wl)
    if prepare is not None:
        prepare(max(args.warmup, 3) + args.steps)
    for _ in range(max(args.warmup, 3)):
        step()
    eng.sync()
    l0 = eng.launch_count
    eng.timer_start()
    for _ in range(args.steps):
        step()
    ms = eng.timer_stop_ms()
    launches = eng.launch_count - l0
    peak, peak_src = measured_peak()
    fps = frames * args.steps / (ms / 1e3)
    achieved = algo * frames * args.steps / (ms / 1e3) / 1e9
    print(json.dumps({"metric": "frames/s (secondary workload)", "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": args.steps,
                      "ms_per_step": ms / args.steps, "config": {"workload": name},
                      "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                                   "algorithmic_bytes_per_frame": algo, "peak_source": peak_src},
                      "gpu_launches": int(launches)}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="independent frames per step per GPU (one kernel launch)")
    ap.add_argument("--e2e-frames", type=int, default=48, help="host frames per e2e step (one batch call)")
    ap.add_argument("--e2e-steps", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub-records", action="store_true", help="skip the cfg4 / cfg5 sub-records of the line")
    ap.add_argument("--cpu-stages", default="auto", choices=["auto", "sws", "loops"],
                    help="CPU arm: sws = convert + resize as the ONE sws_scale call the reference issues (needs a loadable libswscale; auto picks "
                         "it when there is one), loops = the reference's own converter loop + the oracle's resize")
    ap.add_argument("--workload", default="headline", help="headline (the driver's line) | cfg1 | cfg2 | cfg3 | cfg4 | cfg5: kernel-level "
                    "numbers of the other BASELINE configs, N = 1 only")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if args.workload != "headline":
        run_secondary(args)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under it, as the driver would
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr",
               "127.0.0.1", "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
