#!/usr/bin/env python
"""CPU-reference timings for every BASELINE config: the compiled reference loops (oracle/_ref, built by
oracle/build_ref.py from the reference sources where they lie) and, where the reference calls it, a real libswscale
(tests/swscale_ref.py), on this box's host cores.  Two figures per config as SURVEY.md 8d asks: one frame at a time on ONE thread
(nfx_threads = 1), and independent frames spread over all host threads (nproc; the most favourable way to use the cores for a batch --
the reference's own per-frame row-band fan-out is slower).  TEST INFRASTRUCTURE: nothing here is product code.

    python tools/cpu_baselines.py [--out profiles/r02_cpu_baselines.json] [--seconds 3]
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, os.path.join(REPO, "tests"))
import pe_testlib as T  # noqa: E402
import swscale_ref as S  # noqa: E402


def cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def timed(make_worker, seconds, nthreads):
    """make_worker(i) -> callable processing ONE frame with thread-private buffers; returns (fps 1 thread, fps nthreads)"""
    w0 = make_worker(0)
    w0()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds or n < 3:
        w0()
        n += 1
    fps1 = n / (time.perf_counter() - t0)
    workers = [make_worker(i) for i in range(nthreads)]
    per = max(2, int(fps1 * seconds) + 1)

    def loop(w):
        for _ in range(per):
            w()
    with ThreadPoolExecutor(max_workers=nthreads) as ex:
        list(ex.map(lambda w: w(), workers))  # warm
        t0 = time.perf_counter()
        list(ex.map(loop, workers))
        dt = time.perf_counter() - t0
    return fps1, per * nthreads / dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(REPO, "profiles", "r02_cpu_baselines.json"))
    ap.add_argument("--seconds", type=float, default=3.0)
    a = ap.parse_args()
    assert T.have_ref(), "oracle/_ref not built (python oracle/build_ref.py needs /root/reference)"
    r, o, p = T.ref(), T.oracle(), T.ref_paint()
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    sws, sws_ver = S.load()
    nthr = cores()
    mh = C.CDLL(os.path.join(T.REF_DIR, "libweed_minihost.so"))
    mh.mh_open.argtypes = [C.c_char_p]
    mh.mh_run2.argtypes = [T.I, T.I, T.I, T.I, T.I, T.VP, T.I, T.VP, T.I, T.VP, T.I, T.I, T.I]
    blend = mh.mh_open(os.path.join(T.REF_DIR, "simple_blend.so").encode())
    res = {"host_threads": nthr, "libswscale": sws_ver if sws else None, "cpu_model": "", "configs": {}}
    try:
        res["cpu_model"] = [ln.split(":", 1)[1].strip() for ln in open("/proc/cpuinfo") if ln.startswith("model name")][0]
    except Exception:
        pass

    def record(name, what, fps):
        res["configs"][name] = {"what": what, "fps_1_thread": fps[0], "fps_nproc": fps[1]}
        print("%-4s %-100s %10.1f / %10.1f fps" % (name, what[:100], fps[0], fps[1]), flush=True)

    # cfg 1: 640x480 RGB24 -> BGR24 in place, the threaded entry's semantics (hsize = width * 3, colourspace.c:9279) on one thread
    def mk1(i):
        rng = np.random.default_rng(1 + i)
        img = T.make_packed(rng, 640, 480, 3)
        return lambda: r.ref_rgb_permute(0, T.ptr(img), 640 * 3, 480, img.strides[0], img.strides[0], T.ptr(img), None, 0, 0)
    record("1", "640x480 RGB24 -> BGR24 in place, _convert_swap3_frame (colourspace.c:9259) whole rows", timed(mk1, a.seconds, nthr))

    # cfg 2: 1080p YUV420P -> RGBA32 -> 1280x720
    def planes(rng, w, h, is422=False):
        y, u, v = T.make_yuv_planar(rng, w, h, is422, True)
        return y, u, v
    if sws:
        def mk2(i):
            y, u, v = planes(np.random.default_rng(2 + i), 1920, 1080)
            sc = S.Scaler("yuv420p", 1920, 1080, "rgba", 1280, 720, S.SWS_BILINEAR, yuv=(False, False, False))
            dst = np.zeros((720, 1280 * 4), np.uint8)
            return lambda: sc.run([y, u, v], dst)
        record("2", "1080p YUV420P -> RGBA32 1280x720: ONE sws_scale call, as resize_layer_full issues it (colourspace.c:14601-14620), "
               "libswscale " + sws_ver, timed(mk2, a.seconds, nthr))

    def mk2b(i):
        y, u, v = planes(np.random.default_rng(2 + i), 1920, 1080)
        full = np.zeros((1080 + 16, 1920 * 4), np.uint8)
        rgba = full[8:8 + 1080]
        pl, st = T.planes_arg(y, u, v), T.strides_arg(y, u, v)
        if sws:
            sc = S.Scaler("rgba", 1920, 1080, "rgba", 1280, 720, S.SWS_BILINEAR)
            dst = np.zeros((720, 1280 * 4), np.uint8)

            def f():
                r.ref_yuv420p_to_rgb(pl, 1920, 1080, st, 1920 * 4, T.ptr(rgba), 0, 1, 0, 0, 0, 1, 0, 0)
                sc.run([rgba], dst)
            return f
        return lambda: r.ref_yuv420p_to_rgb(pl, 1920, 1080, st, 1920 * 4, T.ptr(rgba), 0, 1, 0, 0, 0, 1, 0, 0)
    record("2b", "1080p convert_layer_palette (convert_yuv420p_to_rgb_frame :3260, HIGH) then resize_layer RGBA -> 1280x720 (sws_scale): "
           "the two-call form", timed(mk2b, a.seconds, nthr))

    # cfg 3: 4K RGBA32 alpha-over + gamma
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, -1, 1, 1.4, T.ptr(lut))

    def mk3(i):
        rng = np.random.default_rng(3 + i)
        bg, fg = T.make_packed(rng, 3840, 2160, 4), T.make_packed(rng, 3840, 2160, 4)
        out = np.zeros_like(bg)

        def f():
            np.copyto(out, bg)
            p.ref_paint_rows(T.ptr(out), T.ptr(fg), 3840 * 2160, 4, 0.5)
            r.ref_gamma_apply(T.ptr(out), 3840 * 4, 4, 0, 3840, 2160, 0, T.ptr(lut))
        return f
    record("3", "4K RGBA32 alpha-over 0.5 (compositor.c paint_pixel :120) + gamma LUT8 (gamma_convert_layer_thread :14034)",
           timed(mk3, a.seconds, nthr))

    # cfg 4: 1080p RGB24 chroma blend bf = 100 through the real plugin and the reference's libweed
    def mk4(i):
        rng = np.random.default_rng(5 + i)
        s1, s2 = T.make_packed(rng, 1920, 1080, 3), T.make_packed(rng, 1920, 1080, 3)
        d = np.zeros_like(s1)
        return lambda: mh.mh_run2(blend, 0, 1, 1920, 1080, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(d), d.strides[0], 100, 1)
    # the minihost is not re-entrant (one plugin table): 1 thread measured, nproc = 1 thread x nproc (independent processes would scale)
    f4 = mk4(0)
    f4()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < a.seconds:
        f4()
        n += 1
    fps4 = n / (time.perf_counter() - t0)
    record("4", "1080p RGB24 'chroma blend' bf = 100: simple_blend.so (simple_blend.c:58) through the reference's libweed, init + process + "
           "deinit per frame; nproc figure = 1-thread x nproc (upper bound)", (fps4, fps4 * nthr))

    # cfg 5: 4K YUV422P -> RGB24 + crossfade with a shared operand
    def mk5(i):
        y, u, v = planes(np.random.default_rng(10 + i), 3840, 2160, True)
        full = np.zeros((2160 + 16, 3840 * 3), np.uint8)
        rgb = full[8:8 + 2160]
        pl, st = T.planes_arg(y, u, v), T.strides_arg(y, u, v)
        op = T.make_packed(np.random.default_rng(99), 3840, 2160, 3)

        def f():
            r.ref_yuv420p_to_rgb(pl, 3840, 2160, st, 3840 * 3, T.ptr(rgb), 0, 0, 1, 0, 0, 1, 0, 0)
            o.pe_or_simple_blend(0, 1, T.ptr(rgb), 3840 * 3, T.ptr(op), op.strides[0], T.ptr(rgb), 3840 * 3, 3840, 2160, 128, op.size)
        return f
    record("5", "4K YUV422P -> RGB24 (convert_yuv420p_to_rgb_frame is_422, HIGH) + chroma blend bf = 128 with the operand (oracle port of "
           "simple_blend.c:58; the plugin's own loop is timed in config 4)", timed(mk5, a.seconds, nthr))

    with open(a.out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", a.out)


if __name__ == "__main__":
    main()
