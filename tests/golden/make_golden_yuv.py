#!/usr/bin/env python
"""Generate tests/golden/ref_vectors_yuv.npz -- the YUV <-> YUV family -- from the COMPILED REFERENCE (oracle/_ref/libref_oracle.so,
built by oracle/build_ref.py from /root/reference).  Run in the build container only; the GPU box consumes the committed .npz.

Only reference behaviour that is defined goes in (see the X rows of DESIGN.md's quirk table): planar 4:4:4 -> RGB24 / RGBA32 /
BGRA32, combineplanes without source alpha on padded planes, splitplanes 3 -> 3 planes, halve / double chroma, packed 4:2:2 ->
planar 4:2:2 (dense buffers; the never-advanced source pointer included) / 4:4:4 / YUV888, swab, the four clamping tables, the packed
chroma up-samplers (planar 4:2:x -> YUV888 / YUVA8888).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pe_testlib as T  # noqa: E402

W, H = 48, 10


def planes444(rng, n):
    st = T.rowstride(W, 1)
    out = []
    for _ in range(n):
        a = np.zeros((H, st), np.uint8)
        a[:, :W] = rng.integers(0, 256, (H, W), dtype=np.uint8)
        out.append(a)
    return out


def main():
    assert T.have_ref(), "build oracle/_ref first (python oracle/build_ref.py)"
    r = T.ref()
    r.ref_set_prefs(1, T.Q_HIGH, 1.4)
    rng = np.random.default_rng(2026)
    out = {}
    for w in range(4):
        t = np.zeros(256, np.uint8)
        assert r.ref_get_yy_table(w, T.ptr(t)) == 0
        out["yy%d" % w] = t
    pl = planes444(rng, 4)
    for k in range(4):
        out["p444_%d" % k] = pl[k]
    for cl in (0, 1):
        for nm, order, ia, oa in (("rgb24", 0, 0, 0), ("rgba32", 0, 1, 1), ("bgra32", 1, 0, 1)):
            ps = 4 if oa else 3
            d = np.zeros((H, T.rowstride(W, ps)), np.uint8)
            r.ref_yuv444p_to_rgb(T.planes_arg(*pl), W, H, pl[0].strides[0], d.strides[0], T.ptr(d), order, ia, oa, cl)
            out["yuv444p_to_%s_cl%d" % (nm, cl)] = d
    for oa in (0, 1):
        ps = 4 if oa else 3
        d = np.zeros((H, T.rowstride(W, ps)), np.uint8)
        r.ref_combineplanes(T.planes_arg(*pl), W, H, pl[0].strides[0], d.strides[0], T.ptr(d), 0, oa)
        out["combine_a%d" % oa] = d
    src = T.make_packed(rng, W, H, 3)
    out["yuv888_src"] = src
    sp = [np.zeros((H, T.rowstride(W, 1)), np.uint8) for _ in range(4)]
    r.ref_splitplanes(T.ptr(src), W, H, src.strides[0], T.strides_arg(*sp), T.planes_arg(*sp), 0, 0)
    for k in range(3):
        out["split_%d" % k] = sp[k]
    # chroma planes 24 x 10 (4:2:2 source) / 24 x 5 (4:2:0 source)
    cw, st = W // 2, T.rowstride(W, 1) // 2
    c422 = [np.zeros((H, st), np.uint8) for _ in range(3)]
    c420 = [np.zeros((H // 2, st), np.uint8) for _ in range(3)]
    for p in c422[1:] + c420[1:]:
        p[:, :cw] = rng.integers(0, 256, (p.shape[0], cw), dtype=np.uint8)
    out["c422_u"], out["c422_v"], out["c420_u"], out["c420_v"] = c422[1], c422[2], c420[1], c420[2]
    for cl in (0, 1):
        d = [np.zeros((H // 2, st), np.uint8) for _ in range(3)]
        r.ref_halve_chroma(T.planes_arg(*c422), cw, H, T.strides_arg(*c422), T.strides_arg(*d), T.planes_arg(*d), cl)
        out["halve_cl%d_u" % cl], out["halve_cl%d_v" % cl] = d[1], d[2]
        d = [np.zeros((H, st), np.uint8) for _ in range(3)]
        r.ref_double_chroma(T.planes_arg(*c420), cw, H // 2, T.strides_arg(*c420), T.strides_arg(*d), T.planes_arg(*d), cl)
        out["double_cl%d_u" % cl], out["double_cl%d_v" % cl] = d[1], d[2]
    wm = W // 2
    m = T.make_packed(rng, wm, H, 4)
    out["mpx_src"] = m
    dense = np.ascontiguousarray(m[:, :wm * 4])
    for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
        d = [np.zeros((H, W), np.uint8), np.zeros((H, wm), np.uint8), np.zeros((H, wm), np.uint8)]
        r.ref_packed422_to_yuv422p(fmt, T.ptr(dense), wm, H, T.planes_arg(*d))
        for k, pn in enumerate("yuv"):
            out["%s_to_yuv422p_%s" % (nm, pn)] = d[k]
        d = [np.zeros((H, T.rowstride(W, 1)), np.uint8) for _ in range(4)]
        r.ref_packed422_to_yuv444p(fmt, T.ptr(m), wm, H, m.strides[0], T.strides_arg(*d), T.planes_arg(*d), 1)
        for k, pn in enumerate("yuva"):
            out["%s_to_yuva4444p_%s" % (nm, pn)] = d[k]
        for aa in (0, 1):
            d8 = np.zeros((H, T.rowstride(W, 4 if aa else 3)), np.uint8)
            r.ref_packed422_to_yuv888(fmt, T.ptr(m), wm, H, m.strides[0], d8.strides[0], T.ptr(d8), aa)
            out["%s_to_yuv888_a%d" % (nm, aa)] = d8
    dense = [np.ascontiguousarray(p[:, :W]) for p in pl[:3]]
    for cl in (0, 1):
        for fmt, nm in ((0, "uyvy"), (1, "yuyv")):
            d = np.zeros((H, 2 * W), np.uint8)
            r.ref_yuv444p_to_packed422(fmt, T.planes_arg(*dense), W, H, W, 2 * W, T.ptr(d), cl)
            out["yuv444p_to_%s_cl%d" % (nm, cl)] = d
        ys = T.rowstride(W, 1)
        d3 = [np.zeros((H, ys), np.uint8), np.zeros((H // 2, ys // 2), np.uint8), np.zeros((H // 2, ys // 2), np.uint8)]
        r.ref_yuv444p_to_yuv420p(T.planes_arg(*pl[:3]), W, H, T.strides_arg(*pl[:3]), T.strides_arg(*d3), T.planes_arg(*d3), cl)
        out["yuv444p_to_yuv420p_cl%d_u" % cl], out["yuv444p_to_yuv420p_cl%d_v" % cl] = d3[1], d3[2]
    # 4:2:0 -> 4:4:4 chroma (convert_quad_chroma): odd height 9 (every row defined), padded planes
    qh, qch = 9, 5
    qs = [np.zeros((qch, T.align_ceil(cw + 1, 16)), np.uint8) for _ in range(3)]
    for p in qs[1:]:
        p[:, :cw] = rng.integers(0, 256, (qch, cw), dtype=np.uint8)
    out["quad_src_u"], out["quad_src_v"] = qs[1], qs[2]
    for samp in (0, 1):
        for cl in (0, 1):
            d = [np.zeros((qh + 2, T.align_ceil(W + 1, 32)), np.uint8) for _ in range(4)]
            r.ref_quad_chroma(T.planes_arg(*qs), W, qh, T.strides_arg(*qs), d[0].strides[0], T.planes_arg(*d), 0, samp, cl)
            out["quad_s%d_cl%d_u" % (samp, cl)], out["quad_s%d_cl%d_v" % (samp, cl)] = d[1][:qh, :W].copy(), d[2][:qh, :W].copy()
    # YUV888 -> UYVY / YUYV / YUV422P / YUV420P on dense buffers
    s888 = np.ascontiguousarray(src[:, :W * 3])
    out["yuv888_dense"] = s888
    for cl in (0, 1):
        for mode in range(4):
            if mode <= 1:
                d = [np.zeros((H, 2 * W), np.uint8)]
            else:
                chh = H if mode == 2 else H // 2
                d = [np.zeros((H, W), np.uint8), np.zeros((chh, W // 2), np.uint8), np.zeros((chh, W // 2), np.uint8)]
            pa = d + [d[0]] * (3 - len(d))
            r.ref_yuv888_subsample(mode, T.ptr(s888), W, H, s888.strides[0], T.strides_arg(*pa), T.planes_arg(*pa), 0, cl)
            for k, a in enumerate(d):
                out["yuv888_sub_m%d_cl%d_%d" % (mode, cl, k)] = a
    sw = m.copy()
    r.ref_swab(T.ptr(sw), wm, H, sw.strides[0])
    out["swab"] = sw
    # planar 4:2:2 / 4:2:0 -> YUV888 / YUVA8888 with the chroma up-sampled on the fly (convert_double_chroma_packed /
    # convert_quad_chroma_packed) on padded planes.  4:2:0: height 9, rows 0 .. 6 (the reference's defined part).  The alpha bytes the
    # reference never writes (second pixel of a pair / odd rows: X) are set to 255 before freezing.
    rg2 = np.random.default_rng(2027)
    uy = np.zeros((H, T.rowstride(W, 1)), np.uint8)  # the engine's own strides: whole rows travel to the device, padding included
    uy[:, :W] = rg2.integers(0, 256, (H, W), dtype=np.uint8)
    out["up_y"] = uy
    for is420, nm, hh, chh in ((0, "422", H, H), (1, "420", 9, 5)):
        cs = [np.zeros((chh, T.rowstride(W, 1) >> 1), np.uint8) for _ in range(2)]
        for p_ in cs:
            p_[:, :cw] = rg2.integers(0, 256, (chh, cw), dtype=np.uint8)
        out["up%s_u" % nm], out["up%s_v" % nm] = cs
        rows = hh if not is420 else hh - 2
        for samp in (0, 1):
            for cl in (0, 1):
                for aa in (0, 1):
                    ps = 4 if aa else 3
                    ors = T.align_ceil(W * ps + 8, 32)
                    d = np.zeros((hh + 2, ors), np.uint8)
                    r.ref_chroma_upsample_packed(is420, T.planes_arg(uy, *cs), W, hh, T.strides_arg(uy, *cs), ors, T.ptr(d), aa, samp, cl)
                    if aa:
                        d[:hh, 3:W * 4:4] = 255
                    out["up%s_s%d_cl%d_a%d" % (nm, samp, cl, aa)] = d[:rows, :W * ps].copy()
    np.savez_compressed(os.path.join(HERE, "ref_vectors_yuv.npz"), **out)
    print("wrote", os.path.join(HERE, "ref_vectors_yuv.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()
