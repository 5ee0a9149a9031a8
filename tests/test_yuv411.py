"""YUV411 (IYU1) as a conversion source -- convert_yuv411_to_{rgb,bgr,argb,yuv888,yuvp,uyvy,yuyv}_frame, src/colourspace.c:8305-8910.

CPU: the oracle against the compiled reference on the buffers the reference defines (dense rows; for RGBA / BGRA the destination is
prefilled with 255, because the reference never writes the alpha bytes of the first pixel pair of a loop iteration, :8338-8370).
GPU: the CUDA converter (pe_convert_layer_palette on a YUV411 layer) against the oracle, padded rowstrides included."""
import itertools
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

# (target, order, add_alpha, palette): target 0 RGB, 1 packed 4:4:4, 2 planar 4:4:4, 3 UYVY, 4 YUYV
CASES = [(0, 0, 0, "RGB24"), (0, 0, 1, "RGBA32"), (0, 1, 0, "BGR24"), (0, 1, 1, "BGRA32"), (0, 2, 1, "ARGB32"),
         (1, 0, 0, "YUV888"), (1, 0, 1, "YUVA8888"), (2, 0, 0, "YUV444P"), (2, 0, 1, "YUVA4444P"), (3, 0, 0, "UYVY"), (4, 0, 0, "YUYV")]


def _src(rng, wm, h, clamped, stride=None):
    """wm macropixels {u2, y0, y1, v2, y2, y3} per row"""
    stride = stride or wm * 6
    a = np.zeros((h, stride), np.uint8)
    lo, hy, hc = (16, 236, 241) if clamped else (0, 256, 256)
    m = rng.integers(lo, hy, (h, wm, 6), dtype=np.uint8)
    m[:, :, 0] = rng.integers(lo, hc, (h, wm), dtype=np.uint8)
    m[:, :, 3] = rng.integers(lo, hc, (h, wm), dtype=np.uint8)
    a[:, :wm * 6] = m.reshape(h, wm * 6)
    return a


def _out_planes(target, add_alpha, wm, h, dense):
    w = 4 * wm
    if target == 2:
        st = w if dense else T.rowstride(w, 1)
        return [np.zeros((h, st), np.uint8) for _ in range(4 if add_alpha else 3)]
    ps = (4 if add_alpha else 3) if target < 3 else 2
    st = w * ps if dense else T.rowstride(w, ps)
    return [np.zeros((h, st), np.uint8)]


def _oracle(target, order, add_alpha, src, wm, h, cl, planes, quirks=1):
    o = T.oracle()
    pl = list(planes) + [planes[0]] * (4 - len(planes))
    o.pe_or_yuv411_to(target, T.ptr(src), src.strides[0], wm, h, T.planes_arg(*pl), T.strides_arg(*pl), order, add_alpha, cl, T.Q_HIGH, quirks)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", CASES, ids=[c[3] for c in CASES])
def test_oracle_equals_compiled_reference(case):
    target, order, add_alpha, _ = case
    r = T.ref()
    rng = np.random.default_rng(411 + target * 7 + order)
    for (wm, h), cl in itertools.product(((1, 2), (2, 3), (9, 4), (40, 5)), (T.CLAMPED, T.UNCLAMPED)):
        src = _src(rng, wm, h, cl == T.CLAMPED)
        exp = _out_planes(target, add_alpha, wm, h, dense=True)
        got = [np.zeros_like(p) for p in exp]
        if target == 0 and add_alpha:
            for p in exp:
                p[:] = 255   # the alpha bytes the reference leaves unwritten
        pl = exp + [exp[0]] * (4 - len(exp))
        r.ref_yuv411_to(target, T.ptr(src), wm, h, exp[0].strides[0], T.planes_arg(*pl), order, add_alpha, cl)
        _oracle(target, order, add_alpha, src, wm, h, cl, got)
        for k in range(len(exp)):
            assert (got[k] == exp[k]).all(), (case, wm, h, cl, k)


def test_known_answers():
    """a flat frame stays flat, and the ladder is monotone between two macropixels"""
    o = T.oracle()
    wm, h = 3, 1
    src = np.zeros((h, wm * 6), np.uint8)
    src[0] = [100, 50, 60, 200, 70, 80, 140, 90, 100, 160, 110, 120, 140, 130, 140, 160, 150, 160]
    out = _out_planes(1, 0, wm, h, dense=True)
    _oracle(1, 0, 0, src, wm, h, T.UNCLAMPED, out)
    px = out[0].reshape(12, 3)
    assert list(px[:, 0]) == [50, 60, 70, 80, 90, 100, 110, 120, 130, 140, 150, 160]
    assert list(px[:2, 1]) == [100, 100] and list(px[-2:, 1]) == [140, 140]
    # the ladder as the reference climbs it (its comments promise 1/8, 3/8, 5/8, 7/8): p = 100, c = 140 -> h 120, qp 110, qc 130
    assert list(px[2:6, 1]) == [105, 125, 115, 135] and list(px[2:6, 2]) == [195, 175, 185, 165]
    assert list(px[6:10, 1]) == [140] * 4


def test_bgr_quirk_and_intended_order():
    """convert_yuv411_to_bgr_frame writes the first pixel of a row and its last two in R, G, B order (:8445, :8514)"""
    rng = np.random.default_rng(5)
    wm, h = 5, 2
    src = _src(rng, wm, h, True)
    rgb, bq, bi = (_out_planes(0, 0, wm, h, dense=True) for _ in range(3))
    _oracle(0, 0, 0, src, wm, h, T.CLAMPED, rgb)
    _oracle(0, 1, 0, src, wm, h, T.CLAMPED, bq, quirks=1)
    _oracle(0, 1, 0, src, wm, h, T.CLAMPED, bi, quirks=0)
    r, q, i = (a[0].reshape(h, 4 * wm, 3) for a in (rgb, bq, bi))
    assert (i == r[:, :, ::-1]).all()
    assert (q[:, 1:-2] == i[:, 1:-2]).all() and (q[:, 0] == r[:, 0]).all() and (q[:, -2:] == r[:, -2:]).all()


# (order, has_alpha, palette) of the RGB sources of convert_{rgb,bgr,argb}_to_yuv411_frame (:6499-6614)
RGB_SOURCES = [(0, 0, "RGB24"), (0, 1, "RGBA32"), (1, 0, "BGR24"), (1, 1, "BGRA32"), (2, 1, "ARGB32")]


def _oracle_from_rgb(order, has_alpha, src, w, h, cl, dest):
    T.oracle().pe_or_rgb_to_yuv411(T.ptr(src), src.strides[0], w, h, T.ptr(dest), dest.strides[0], order, has_alpha, cl)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", RGB_SOURCES, ids=[c[2] for c in RGB_SOURCES])
def test_rgb_to_yuv411_oracle_equals_compiled_reference(case):
    order, has_alpha, _ = case
    r = T.ref()
    rng = np.random.default_rng(4110 + order)
    for (w, h), cl in itertools.product(((4, 1), (8, 3), (23, 4), (64, 5)), (T.CLAMPED, T.UNCLAMPED)):
        src = T.make_packed(rng, w, h, 4 if has_alpha else 3)
        wm = w >> 2
        exp, got = np.zeros((h, wm * 6), np.uint8), np.zeros((h, wm * 6), np.uint8)
        r.ref_rgb_to_yuv411(T.ptr(src), w, h, src.strides[0], T.ptr(exp), order, has_alpha, cl)
        _oracle_from_rgb(order, has_alpha, src, w, h, cl, got)
        assert (got == exp).all(), (case, w, h, cl)


@pytest.mark.gpu
@pytest.mark.parametrize("case", RGB_SOURCES, ids=[c[2] for c in RGB_SOURCES])
def test_rgb_to_yuv411_cuda_equals_oracle_and_round_trip_geometry(case):
    lb = pytest.importorskip("lives_b200")
    order, has_alpha, pal = case
    eng = lb.Engine()
    rng = np.random.default_rng(5950 + order)
    for (w, h), cl in itertools.product(((4, 1), (23, 4), (64, 5), (1920, 270)), (T.CLAMPED, T.UNCLAMPED)):
        src = T.make_packed(rng, w, h, 4 if has_alpha else 3)
        lay = lb.Layer.from_host(eng, T.PAL[pal], w, h, [src])
        assert lb.convert_layer_palette(lay, T.PAL["YUV411"], cl), lb.capi.last_error()
        wm = w >> 2
        assert lay.palette == T.PAL["YUV411"] and lay.width == 4 * wm and lay.height == h and lay.desc.yuv_clamping == cl
        got = lay.to_host()[0]
        exp = np.zeros_like(got)
        _oracle_from_rgb(order, has_alpha, src, w, h, cl, exp)
        assert (got[:, :wm * 6] == exp[:, :wm * 6]).all(), (case, w, h, cl)
        assert lb.convert_layer_palette(lay, T.PAL[pal], cl) and lay.palette == T.PAL[pal] and lay.width == 4 * wm   # and back
    eng.close()


# ---- the remaining converters: YUV411 -> planar 4:2:2 / 4:2:0, and UYVY / YUYV / YUV420P / YUV422P / YUV888 / YUVA8888 / YUV444P -> YUV411
TO_411 = [(0, "UYVY"), (1, "YUYV"), (2, "YUV420P"), (3, "YUV422P"), (4, "YUV888"), (5, "YUVA8888"), (6, "YUV444P")]


def _yuv_source(rng, mode, w, h, dense):
    """planes of a w x h frame of the source palette of `mode` (see pe_or_to_yuv411)"""
    def plane(cols, rows):
        st = cols if dense else T.rowstride(cols, 1)
        a = np.zeros((rows, st), np.uint8)
        a[:, :cols] = rng.integers(16, 236, (rows, cols), dtype=np.uint8)
        return a
    if mode <= 1:
        return [plane(2 * w, h)]
    if mode == 2:
        return [plane(w, h), plane(w // 2, (h + 1) // 2), plane(w // 2, (h + 1) // 2)]
    if mode == 3:
        return [plane(w, h), plane(w // 2, h), plane(w // 2, h)]
    if mode <= 5:
        return [plane(w * (3 if mode == 4 else 4), h)]
    return [plane(w, h), plane(w, h), plane(w, h)]


def _oracle_to_411(mode, planes, w, h, cl, dest):
    pl = list(planes) + [planes[0]] * (3 - len(planes))
    T.oracle().pe_or_to_yuv411(mode, T.planes_arg(*pl), T.strides_arg(*pl), w, h, T.ptr(dest), dest.strides[0], cl)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("case", TO_411, ids=[c[1] for c in TO_411])
def test_to_yuv411_oracle_equals_compiled_reference(case):
    mode, _ = case
    r = T.ref()
    rng = np.random.default_rng(4200 + mode)
    for (w, h), cl in itertools.product(((4, 1), (8, 2), (24, 6), (64, 9)), (T.CLAMPED, T.UNCLAMPED)):
        if mode == 2 and h & 1 and h > 1:
            h += 1   # 4:2:0 frames are even (colourspace.c:11603)
        if mode == 6:
            # convert_yuvp_to_yuv411_frame never advances its output pointer (:7755-7797): pinned macropixel by macropixel
            for _ in range(64):
                pl = _yuv_source(rng, 6, 4, 1, dense=True)
                exp, got = np.zeros((1, 6), np.uint8), np.zeros((1, 6), np.uint8)
                r.ref_to_yuv411(6, T.planes_arg(*pl), 4, 1, 4, T.ptr(exp), cl)
                _oracle_to_411(6, pl, 4, 1, cl, got)
                assert (got == exp).all(), (cl, pl)
            continue
        pl = _yuv_source(rng, mode, w, h, dense=True)
        wm = w >> 2
        exp, got = np.zeros((h, wm * 6), np.uint8), np.zeros((h, wm * 6), np.uint8)
        r.ref_to_yuv411(mode, T.planes_arg(*(pl + [pl[0]] * (3 - len(pl)))), (w // 2) if mode <= 1 else w, h, pl[0].strides[0], T.ptr(exp), cl)
        _oracle_to_411(mode, pl, w, h, cl, got)
        rows = h
        if mode in (4, 5):   # the reference stops after width * height BYTES of its 3- / 4-byte pixels (:8278)
            ips = 3 if mode == 4 else 4
            rows = -(-h // ips)
            assert (exp[rows:] == 0).all()
        assert (got[:rows] == exp[:rows]).all(), (case, w, h, cl)


@pytest.mark.skipif(not T.have_ref(), reason="oracle/_ref not built")
def test_yuv411_to_planar_422_and_420_oracle_equals_compiled_reference():
    r = T.ref()
    rng = np.random.default_rng(4222)
    for (wm, h), cl in itertools.product(((1, 1), (2, 3), (9, 4), (40, 5)), (T.CLAMPED, T.UNCLAMPED)):
        src = _src(rng, wm, h, cl == T.CLAMPED)
        w = 4 * wm
        exp = [np.zeros((h, w), np.uint8), np.zeros((h, w // 2), np.uint8), np.zeros((h, w // 2), np.uint8)]
        got = [np.zeros_like(p) for p in exp]
        r.ref_yuv411_to(5, T.ptr(src), wm, h, 0, T.planes_arg(*(exp + [exp[0]])), 0, 0, cl)
        _oracle(5, 0, 0, src, wm, h, cl, got)
        for k in range(3):
            assert (got[k] == exp[k]).all(), ("422P", wm, h, cl, k)
    # 4:2:0: the reference only ever writes chroma row 0 (:9060-9140); a one-row frame is the case it defines
    for wm, cl in itertools.product((1, 2, 9, 40), (T.CLAMPED, T.UNCLAMPED)):
        src = _src(rng, wm, 1, cl == T.CLAMPED)
        w = 4 * wm
        exp = [np.zeros((1, w), np.uint8), np.zeros((1, w // 2), np.uint8), np.zeros((1, w // 2), np.uint8)]
        got = [np.zeros_like(p) for p in exp]
        r.ref_yuv411_to(6, T.ptr(src), wm, 1, 0, T.planes_arg(*(exp + [exp[0]])), 0, 0, cl)
        _oracle(6, 0, 0, src, wm, 1, cl, got)
        for k in range(3):
            assert (got[k] == exp[k]).all(), ("420P", wm, cl, k)


def test_yuv411_to_420_is_the_vertical_average_of_its_422():
    rng = np.random.default_rng(4201)
    wm, h = 7, 5
    src = _src(rng, wm, h, True)
    w = 4 * wm
    p422 = [np.zeros((h, w), np.uint8), np.zeros((h, w // 2), np.uint8), np.zeros((h, w // 2), np.uint8)]
    p420 = [np.zeros((h, w), np.uint8), np.zeros(((h + 1) // 2, w // 2), np.uint8), np.zeros(((h + 1) // 2, w // 2), np.uint8)]
    _oracle(5, 0, 0, src, wm, h, T.CLAMPED, p422)
    _oracle(6, 0, 0, src, wm, h, T.CLAMPED, p420)
    avg = np.zeros(65536, np.uint8)
    T.oracle().pe_or_avg_table(0, T.ptr(avg))
    assert (p420[0] == p422[0]).all()
    for k in (1, 2):
        for r_ in range((h + 1) // 2):
            a = p422[k][2 * r_]
            e = a if 2 * r_ + 1 >= h else avg[(a.astype(int) << 8) + p422[k][2 * r_ + 1]]
            assert (p420[k][r_] == e).all()


@pytest.mark.gpu
@pytest.mark.parametrize("case", TO_411 + [(2, "YVU420P"), (6, "YUVA4444P")], ids=[c[1] for c in TO_411] + ["YVU420P", "YUVA4444P"])
def test_to_yuv411_cuda_equals_oracle(case):
    lb = pytest.importorskip("lives_b200")
    mode, pal = case
    eng = lb.Engine()
    rng = np.random.default_rng(5960 + mode)
    for (w, h), cl in itertools.product(((4, 2), (24, 6), (64, 10), (1920, 270)), (T.CLAMPED, T.UNCLAMPED)):
        pl = _yuv_source(rng, mode, w, h, dense=False)
        if pal == "YUVA4444P":
            pl = pl + [np.full_like(pl[0], 200)]
        up = [pl[0], pl[2], pl[1]] if pal == "YVU420P" else pl     # a YVU layer holds Cr in plane 1
        lay = lb.Layer.from_host(eng, T.PAL[pal], w, h, up, yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, T.PAL["YUV411"], cl), lb.capi.last_error()
        assert lay.palette == T.PAL["YUV411"] and lay.width == w and lay.height == h
        got = lay.to_host()[0]
        exp = np.zeros_like(got)
        _oracle_to_411(mode, pl[:3], w, h, cl, exp)
        assert (got[:, :(w >> 2) * 6] == exp[:, :(w >> 2) * 6]).all(), (case, w, h, cl)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("pal", ["YUV422P", "YUV420P", "YVU420P"])
def test_yuv411_to_planar_42x_cuda_equals_oracle(pal):
    lb = pytest.importorskip("lives_b200")
    eng = lb.Engine()
    rng = np.random.default_rng(5970)
    target = 5 if pal == "YUV422P" else 6
    for (wm, h), cl in itertools.product(((1, 2), (9, 4), (40, 18), (480, 270)), (T.CLAMPED, T.UNCLAMPED)):
        w = 4 * wm
        src = _src(rng, wm, h, cl == T.CLAMPED, stride=T.align_ceil(wm * 6, 32))
        lay = lb.Layer.from_host(eng, T.PAL["YUV411"], w, h, [src], yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, T.PAL[pal], cl), lb.capi.last_error()
        got = lay.to_host()
        assert lay.palette == T.PAL[pal] and lay.width == w and lay.height == h
        exp = [np.zeros_like(p) for p in got]
        if pal == "YVU420P":
            exp = [exp[0], exp[2], exp[1]]       # the oracle writes Cb to dest[1]; the finished YVU layer holds it in plane 2
        _oracle(target, 0, 0, src, wm, h, cl, exp)
        if pal == "YVU420P":
            exp = [exp[0], exp[2], exp[1]]
        for k in range(3):
            cols = w if k == 0 else w // 2
            assert (got[k][:, :cols] == exp[k][:, :cols]).all(), (pal, wm, h, cl, k)
    eng.close()


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_vectors_yuv411.npz")


def test_oracle_equals_golden_vectors():
    """the outputs of the compiled reference, frozen by tests/golden/make_golden_yuv411.py (for boxes without /root/reference)"""
    g = np.load(GOLDEN)
    wm, h = 12, 6
    for cl in (T.CLAMPED, T.UNCLAMPED):
        src = g["src_cl%d" % cl]
        for target, order, add_alpha, pal in CASES:
            got = _out_planes(target, add_alpha, wm, h, dense=True)
            _oracle(target, order, add_alpha, src, wm, h, cl, got)
            for k, p in enumerate(got):
                assert (p == g["%s_cl%d_p%d" % (pal, cl, k)]).all(), (pal, cl, k)
        for order, has_alpha, pal in RGB_SOURCES:
            rgb = g["rgbsrc_%s" % pal]
            got = np.zeros((rgb.shape[0], 6 * 6), np.uint8)
            _oracle_from_rgb(order, has_alpha, rgb, 24, rgb.shape[0], cl, got)
            assert (got == g["from_%s_cl%d" % (pal, cl)]).all(), (pal, cl)


@pytest.mark.gpu
def test_cuda_equals_golden_vectors():
    lb = pytest.importorskip("lives_b200")
    g = np.load(GOLDEN)
    eng = lb.Engine()
    wm, h = 12, 6
    for cl in (T.CLAMPED, T.UNCLAMPED):
        src = g["src_cl%d" % cl]
        for target, order, add_alpha, pal in CASES:
            lay = lb.Layer.from_host(eng, T.PAL["YUV411"], 4 * wm, h, [src], yuv_clamping=cl)
            assert lb.convert_layer_palette(lay, T.PAL[pal], cl)
            got = lay.to_host()
            for k, p in enumerate(got):
                e = g["%s_cl%d_p%d" % (pal, cl, k)]
                assert (p[:, :e.shape[1]] == e).all(), (pal, cl, k)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c[3] for c in CASES])
def test_cuda_equals_oracle(case):
    lb = pytest.importorskip("lives_b200")
    target, order, add_alpha, pal = case
    eng = lb.Engine()
    rng = np.random.default_rng(595 + target)
    for (wm, h), cl in itertools.product(((1, 1), (2, 3), (9, 4), (40, 17), (480, 270)), (T.CLAMPED, T.UNCLAMPED)):
        w = 4 * wm
        src = _src(rng, wm, h, cl == T.CLAMPED, stride=T.align_ceil(wm * 6, 32))
        lay = lb.Layer.from_host(eng, T.PAL["YUV411"], w, h, [src], yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, T.PAL[pal], cl), (case, lb.capi.last_error())
        got = lay.to_host()
        assert lay.palette == T.PAL[pal] and lay.width == w and lay.height == h
        exp = [np.zeros_like(p) for p in got]
        _oracle(target, order, add_alpha, src, wm, h, cl, exp)
        rb = [lb.plane_row_bytes(T.PAL[pal], w, k) for k in range(len(got))]
        for k in range(len(got)):
            assert (got[k][:, :rb[k]] == exp[k][:, :rb[k]]).all(), (case, wm, h, cl, k)
    eng.close()


@pytest.mark.gpu
def test_cuda_refuses_bad_geometry_and_leaves_the_layer_untouched():
    lb = pytest.importorskip("lives_b200")
    eng = lb.Engine()
    rgb_src = T.make_packed(np.random.default_rng(2), 3, 2, 3)
    rgb = lb.Layer.from_host(eng, T.PAL["RGB24"], 3, 2, [rgb_src])
    assert not lb.convert_layer_palette(rgb, T.PAL["YUV411"], T.CLAMPED)      # narrower than one macropixel
    assert rgb.palette == T.PAL["RGB24"] and (rgb.to_host()[0][:, :9] == rgb_src[:, :9]).all()
    eng.close()
