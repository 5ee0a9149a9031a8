"""GPU parity: the CUDA path, called through the C ABI (lives_b200 -> libpe_b200.so), against the CPU oracle
(oracle/libpe_oracle.so, itself pinned to the compiled reference by tests/test_oracle_vs_reference.py).

Bit-exact everywhere (integer / byte work).  Run on the B200 box: pytest -m gpu.
"""
import ctypes as C
import itertools
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import pe_testlib as T  # noqa: E402

pytestmark = pytest.mark.gpu

lb = pytest.importorskip("lives_b200")

RGB_PALS = (1, 2, 3, 4, 5)
ORDER_OF = {1: (0, 0), 2: (1, 0), 3: (0, 1), 4: (1, 1), 5: (2, 1)}  # palette -> (oracle order, add_alpha)


@pytest.fixture(scope="module")
def eng():
    e = lb.Engine()
    yield e
    e.close()


def packed_layer(eng, pal, w, h, arr, **kw):
    return lb.Layer.from_host(eng, pal, w, h, [arr], **kw)


def payload(arr, w, ps):
    return arr[:, :w * ps]


# ------------------------------------------------------------------------------------------------ RGB <-> RGB

@pytest.mark.parametrize("size", [(640, 480), (37, 11), (1, 1), (1921, 7)])
def test_config1_rgb24_to_bgr24_and_all_rgb_pairs(eng, size):
    """BASELINE config 1 (640x480 RGB24 -> BGR24 convert_layer_palette) + every pair of the 5 RGB palettes"""
    o = T.oracle()
    rng = np.random.default_rng(1)
    w, h = size
    for ipal, opal in itertools.product(RGB_PALS, RGB_PALS):
        if ipal == opal:
            continue
        ips, ops = T.psize_of(ipal), T.psize_of(opal)
        src = T.make_packed(rng, w, h, ips)
        exp = np.zeros((h, T.rowstride(w, ops)), np.uint8)
        assert o.pe_or_rgb_to_rgb(ipal, opal, T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], None) == 0
        lay = packed_layer(eng, ipal, w, h, src)
        assert lb.convert_layer_palette(lay, opal, 0)
        assert lay.palette == opal and lay.width == w and lay.height == h
        got = lay.to_host()[0]
        assert got.shape == exp.shape
        assert (payload(got, w, ops) == payload(exp, w, ops)).all(), (ipal, opal)
        lay.free()


def test_rgb_permutation_batch(eng):
    """pe_convert_layer_palette_batch on RGB layers: runs of same-shaped layers leave as one launch, results as layer by layer"""
    o = T.oracle()
    rng = np.random.default_rng(3)
    for ipal, opal in ((1, 2), (1, 3), (4, 1), (3, 5)):
        ips, ops = T.psize_of(ipal), T.psize_of(opal)
        shapes = [(640, 480), (640, 480), (640, 480), (37, 11), (640, 480), (640, 480)]
        srcs = [T.make_packed(rng, w, h, ips) for w, h in shapes]
        lays = [packed_layer(eng, ipal, w, h, s) for (w, h), s in zip(shapes, srcs)]
        before = eng.launch_count
        assert lb.convert_layer_palette_batch(lays, opal, 0) == len(lays)
        assert eng.launch_count - before == 3  # runs of 3, 1 and 2 layers
        for (w, h), s, lay in zip(shapes, srcs, lays):
            exp = np.zeros((h, T.rowstride(w, ops)), np.uint8)
            assert o.pe_or_rgb_to_rgb(ipal, opal, T.ptr(s), s.strides[0], w, h, T.ptr(exp), exp.strides[0], None) == 0
            assert lay.palette == opal
            assert (payload(lay.to_host()[0], w, ops) == payload(exp, w, ops)).all(), (ipal, opal, w, h)


def test_rgb_to_rgb_with_gamma_lut(eng):
    """gamma_lut8 inside the permutation (colourspace.c:12372), tgt_gamma given"""
    o = T.oracle()
    rng = np.random.default_rng(2)
    w, h = 123, 17
    for (gf, gt), (ipal, opal) in itertools.product(((T.G_SRGB, T.G_LINEAR), (T.G_LINEAR, T.G_SRGB), (T.G_SRGB, T.G_BT709),
                                                     (T.G_LINEAR, T.G_MONITOR)), ((1, 2), (1, 3), (3, 1), (4, 5), (5, 3))):
        ips, ops = T.psize_of(ipal), T.psize_of(opal)
        src = T.make_packed(rng, w, h, ips)
        lut = np.zeros(256, np.uint8)
        assert o.pe_or_gamma_lut8(1.0, gf, gt, 1.4, T.ptr(lut)) == 0
        exp = np.zeros((h, T.rowstride(w, ops)), np.uint8)
        o.pe_or_rgb_to_rgb(ipal, opal, T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], T.ptr(lut))
        lay = packed_layer(eng, ipal, w, h, src, gamma_type=gf)
        assert lb.convert_layer_palette_full(lay, opal, 0, 0, 0, gt)
        assert lay.gamma_type == gt
        got = lay.to_host()[0]
        assert (payload(got, w, ops) == payload(exp, w, ops)).all(), (gf, gt, ipal, opal)
        assert (eng.gamma_lut8(1.0, gf, gt) == lut).all()


def test_rgb_roundtrip_4k_property(eng):
    """size-independent property at 4K: RGB24 -> BGRA32 -> ARGB32 -> RGB24 is the identity"""
    rng = np.random.default_rng(3)
    w, h = 3840, 2160
    src = T.make_packed(rng, w, h, 3)
    lay = packed_layer(eng, 1, w, h, src)
    for pal in (4, 5, 2, 3, 1):
        assert lb.convert_layer_palette(lay, pal, 0)
    got = lay.to_host()[0]
    assert (payload(got, w, 3) == payload(src, w, 3)).all()


# ------------------------------------------------------------------------------------------------ planar YUV -> RGB

def _oracle_planar(o, y, u, v, w, h, opal, is422, cl, sub, q, quirks=1, lut16=None):
    order, add_alpha = ORDER_OF[opal]
    ps = T.psize_of(opal)
    exp = np.zeros((h, T.rowstride(w, ps)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], order, add_alpha,
                           is422, cl, sub, q, quirks, T.ptr(lut16))
    return exp


@pytest.mark.parametrize("size", [(64, 48), (130, 34), (2, 2), (6, 4), (642, 362), (644, 362), (256, 34), (4, 2), (8, 4)])
@pytest.mark.parametrize("is422", [0, 1])
def test_planar_yuv_to_rgb(eng, size, is422):
    """convert_yuv420p_to_{rgb,bgr,argb}_frame (colourspace.c:3260,3927,4527) through convert_layer_palette_full"""
    o = T.oracle()
    rng = np.random.default_rng(4)
    w, h = size
    inpal = T.PAL["YUV422P"] if is422 else T.PAL["YUV420P"]
    for opal, cl, sub in itertools.product(RGB_PALS, (0, 1), (1, 2)):
        y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), cl == 0)
        exp = _oracle_planar(o, y, u, v, w, h, opal, is422, cl, sub, T.Q_HIGH)
        lay = lb.Layer.from_host(eng, inpal, w, h, [y, u, v], yuv_clamping=cl, yuv_subspace=sub)
        assert lb.convert_layer_palette_full(lay, opal, cl, 0, sub, 0)
        got = lay.to_host()[0]
        ps = T.psize_of(opal)
        assert (payload(got, w, ps) == payload(exp, w, ps)).all(), (opal, cl, sub)
        assert lay.yuv_clamping == 0 and lay.yuv_subspace == 0
        lay.free()


def test_planar_yuv_marching_kernel(eng, monkeypatch):
    """k_yuv_march (the strip-marching converter that large frames take) forced onto small frames: every RGB palette, both
    samplings, clamped / unclamped, quirks on / off, padded and unpadded chroma planes, with and without the fused crossfade;
    then a 4K 4:2:2 clip"""
    o = T.oracle()
    rng = np.random.default_rng(41)
    monkeypatch.setenv("PE_YUV_MARCH_MIN_ROWS", "0")
    e_nq = lb.Engine(ref_quirks=False)
    for (w, h), is422 in itertools.product(((64, 48), (644, 362), (256, 34), (8, 4), (132, 50), (4, 2)), (0, 1)):
        inpal = 522 if is422 else 512
        for opal, cl, sub, quirks in itertools.product(RGB_PALS, (0, 1), (1, 2), (1, 0)):
            if (cl, sub, quirks) not in ((0, 1, 1), (1, 2, 1), (0, 2, 0), (1, 1, 0)) and (w, h) != (64, 48):
                continue
            e = eng if quirks else e_nq
            y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), cl == 0)
            exp = _oracle_planar(o, y, u, v, w, h, opal, is422, cl, sub, T.Q_HIGH, quirks=quirks)
            lay = lb.Layer.from_host(e, inpal, w, h, [y, u, v], yuv_clamping=cl, yuv_subspace=sub)
            assert lb.convert_layer_palette_full(lay, opal, cl, 0, sub, 0)
            ps = T.psize_of(opal)
            bad = np.argwhere(payload(lay.to_host()[0], w, ps) != payload(exp, w, ps))
            assert len(bad) == 0, (w, h, is422, opal, cl, sub, quirks, len(bad), bad[:4])
            lay.free()
        # fused crossfade through the marching kernel
        y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), True)
        operand = T.make_packed(rng, w, h, 3)
        exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
        o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], 0, 0, is422, 0, 1,
                               T.Q_HIGH, 1, None)
        o.pe_or_simple_blend(0, 1, T.ptr(exp), exp.strides[0], T.ptr(operand), operand.strides[0], T.ptr(exp), exp.strides[0], w, h, 77,
                             operand.size)
        clip = lb.Layer.from_host(eng, inpal, w, h, [y, u, v], yuv_subspace=1)
        lb.convert_crossfade(clip, packed_layer(eng, 1, w, h, operand), 1, 0, 77)
        assert (payload(clip.to_host()[0], w, 3) == payload(exp, w, 3)).all(), (w, h, is422, "crossfade")
    e_nq.close()
    monkeypatch.setenv("PE_YUV_MARCH_MIN_ROWS", "12")
    w, h = 3840, 2160
    y, u, v = T.make_yuv_planar(rng, w, h, True, True)
    exp = _oracle_planar(o, y, u, v, w, h, 1, 1, 0, 1, T.Q_HIGH)
    lay = lb.Layer.from_host(eng, 522, w, h, [y, u, v], yuv_subspace=1)
    assert lb.convert_layer_palette(lay, 1, 0)
    assert (payload(lay.to_host()[0], w, 3) == payload(exp, w, 3)).all()


def test_planar_yuv_quality_quirks_and_yvu(eng):
    """PB_QUALITY_LOW chroma shortcut (RGB order only, :3470), ref_quirks off, YVU420P plane swap (:12354)"""
    o = T.oracle()
    rng = np.random.default_rng(5)
    w, h = 98, 50
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    e_low = lb.Engine(pb_quality=lb.PB_QUALITY_LOW)
    for opal in RGB_PALS:
        exp = _oracle_planar(o, y, u, v, w, h, opal, 0, 0, 1, T.Q_LOW)
        lay = lb.Layer.from_host(e_low, 512, w, h, [y, u, v], yuv_subspace=1)
        assert lb.convert_layer_palette(lay, opal, 0)
        ps = T.psize_of(opal)
        assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), opal
    e_low.close()
    e_nq = lb.Engine(ref_quirks=False)
    for (w2, h2), is422 in itertools.product(((w, h), (132, 50), (320, 48)), (0, 1)):  # 98: the general kernel; 132, 320: the fast one
        yy, uu, vv = T.make_yuv_planar(rng, w2, h2, bool(is422), True)
        exp = _oracle_planar(o, yy, uu, vv, w2, h2, 3, is422, 0, 1, T.Q_HIGH, quirks=0)
        lay = lb.Layer.from_host(e_nq, 522 if is422 else 512, w2, h2, [yy, uu, vv], yuv_subspace=1)
        assert lb.convert_layer_palette(lay, 3, 0)
        assert (payload(lay.to_host()[0], w2, 4) == payload(exp, w2, 4)).all(), (w2, h2, is422)
    e_nq.close()
    exp = _oracle_planar(o, y, u, v, w, h, 1, 0, 0, 1, T.Q_HIGH)
    lay = lb.Layer.from_host(eng, 513, w, h, [y, v, u], yuv_subspace=1)  # YVU: V plane first
    assert lb.convert_layer_palette(lay, 1, 0)
    assert (payload(lay.to_host()[0], w, 3) == payload(exp, w, 3)).all()


def test_planar_yuv_inline_gamma(eng):
    """xyuv2rgb_with_gamma (colourspace.c:2386): 16-bit LUT inside the converter when the layer has a gamma type"""
    o = T.oracle()
    rng = np.random.default_rng(6)
    w, h = 64, 48
    for gf, gt, opal in ((T.G_SRGB, T.G_LINEAR, 3), (T.G_SRGB, T.G_BT709, 1), (T.G_LINEAR, T.G_SRGB, 2), (T.G_BT709, T.G_SRGB, 5)):
        y, u, v = T.make_yuv_planar(rng, w, h, False, True)
        lut = np.zeros(65536, np.uint16)
        assert o.pe_or_gamma_lut16(1.0, gf, gt, 1.4, T.ptr(lut)) == 0
        exp = _oracle_planar(o, y, u, v, w, h, opal, 0, 0, 1, T.Q_HIGH, lut16=lut)
        lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_subspace=1, gamma_type=gf)
        assert lb.convert_layer_palette_full(lay, opal, 0, 0, 1, gt)
        assert lay.gamma_type == gt
        ps = T.psize_of(opal)
        assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), (gf, gt, opal)


def test_config2_1080p_yuv420p_to_rgba_resize_720p(eng):
    """BASELINE config 2: 1920x1080 YUV420P (clamped, BT.601) -> RGBA32 -> bilinear 1280x720.  Conversion is
    bit-exact with the reference arithmetic; the resize follows OUR published filter (parity unpinned: libswscale)."""
    o = T.oracle()
    rng = np.random.default_rng(2)
    w, h, dw, dh = 1920, 1080, 1280, 720
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    rgba = _oracle_planar(o, y, u, v, w, h, 3, 0, 0, 1, T.Q_HIGH)
    exp = np.zeros((dh, T.rowstride(dw, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], w, h, T.ptr(exp), exp.strides[0], dw, dh, 4)
    lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_clamping=0, yuv_subspace=1)
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, lb.WEED_PALETTE_RGBA32, 0)
    assert (lay.palette, lay.width, lay.height) == (3, dw, dh)
    got = lay.to_host()[0]
    assert (payload(got, dw, 4) == payload(exp, dw, 4)).all()
    # sanity against an independent area-average (float) of the converted frame: mean abs diff well below 1 LSB
    ref = rgba[:, :w * 4].reshape(h, w, 4).astype(np.float64)
    assert abs(float(got[:, :dw * 4].mean()) - float(ref.mean())) < 0.5  # 1.5x down: compare the global mean
    # the batch entry points give the same frames (a second, different frame rides along)
    y2, u2, v2 = T.make_yuv_planar(rng, w, h, False, True)
    rgba2 = _oracle_planar(o, y2, u2, v2, w, h, 3, 0, 0, 1, T.Q_HIGH)
    exp2 = np.zeros_like(exp)
    o.pe_or_resize_packed(T.ptr(rgba2), rgba2.strides[0], w, h, T.ptr(exp2), exp2.strides[0], dw, dh, 4)
    lays = [lb.Layer.from_host(eng, 512, w, h, p, yuv_clamping=0, yuv_subspace=1) for p in ([y, u, v], [y2, u2, v2])]
    assert lb.resize_layer_batch(lays, dw, dh, lb.LIVES_INTERP_NORMAL, lb.WEED_PALETTE_RGBA32, 0) == 2
    for lay, e in zip(lays, (exp, exp2)):
        assert (payload(lay.to_host()[0], dw, 4) == payload(e, dw, 4)).all()
    lays = [lb.Layer.from_host(eng, 512, w, h, p, yuv_clamping=0, yuv_subspace=1) for p in ([y, u, v], [y2, u2, v2])]
    assert lb.convert_layer_palette_batch(lays, 3, 0) == 2
    for lay, e in zip(lays, (rgba, rgba2)):
        assert (payload(lay.to_host()[0], w, 4) == payload(e, w, 4)).all()
    # a longer batch with a differently shaped layer in the middle: three launches (run of 3, the odd one, run of 2)
    ys, us, vs = T.make_yuv_planar(rng, 320, 180, False, True)
    rgba_s = _oracle_planar(o, ys, us, vs, 320, 180, 3, 0, 0, 1, T.Q_HIGH)
    srcs = [(w, h, (y, u, v)), (w, h, (y2, u2, v2)), (w, h, (y, u, v)), (320, 180, (ys, us, vs)), (w, h, (y2, u2, v2)), (w, h, (y, u, v))]
    lays = [lb.Layer.from_host(eng, 512, ww, hh, list(p), yuv_clamping=0, yuv_subspace=1) for ww, hh, p in srcs]
    before = eng.launch_count
    assert lb.convert_layer_palette_batch(lays, 3, 0) == 6
    assert eng.launch_count - before == 3
    for lay, e, ww in zip(lays, (rgba, rgba2, rgba, rgba_s, rgba2, rgba), (w, w, w, 320, w, w)):
        assert (payload(lay.to_host()[0], ww, 4) == payload(e, ww, 4)).all()


# ------------------------------------------------------------------------------------------------ packed YUV

@pytest.mark.parametrize("fmt", [0, 1])
def test_packed422_to_rgb(eng, fmt):
    o = T.oracle()
    rng = np.random.default_rng(7)
    wm, h = 49, 21
    inpal = 564 if fmt == 0 else 565
    for opal, cl, sub in itertools.product(RGB_PALS, (0, 1), (1, 2)):
        order, add_alpha = ORDER_OF[opal]
        ps = T.psize_of(opal)
        src = T.make_packed(rng, wm, h, 4)
        exp = np.zeros((h, T.rowstride(wm * 2, ps)), np.uint8)
        # table choice of the dispatcher: uyvy->RGB24 honours the subspace, uyvy->RGBA32 gets the SAMPLING (0 here)
        osub = sub if (fmt == 0 and opal == 1) else 1
        o.pe_or_packed422_to_rgb(fmt, T.ptr(src), src.strides[0], wm, h, T.ptr(exp), exp.strides[0], order, add_alpha, cl,
                                 osub, T.Q_HIGH)
        lay = lb.Layer.from_host(eng, inpal, wm * 2, h, [src], yuv_clamping=cl, yuv_subspace=sub)
        assert lb.convert_layer_palette_full(lay, opal, cl, 0, sub, 0)
        assert lay.width == wm * 2
        assert (payload(lay.to_host()[0], wm * 2, ps) == payload(exp, wm * 2, ps)).all(), (opal, cl, sub)


def test_yuv888_both_directions(eng):
    o = T.oracle()
    rng = np.random.default_rng(8)
    w, h = 50, 13
    for inpal, opal, cl in itertools.product((588, 589), RGB_PALS, (0, 1)):
        order, out_alpha = ORDER_OF[opal]
        ips, ops = T.psize_of(inpal), T.psize_of(opal)
        src = T.make_packed(rng, w, h, ips)
        exp = np.zeros((h, T.rowstride(w, ops)), np.uint8)
        o.pe_or_yuv888_to_rgb(T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], order, int(inpal == 589), out_alpha,
                              cl, 1, T.Q_HIGH)
        lay = lb.Layer.from_host(eng, inpal, w, h, [src], yuv_clamping=cl, yuv_subspace=1)
        assert lb.convert_layer_palette(lay, opal, cl)
        assert (payload(lay.to_host()[0], w, ops) == payload(exp, w, ops)).all(), (inpal, opal, cl)
    for ipal, opal, cl in itertools.product(RGB_PALS, (588, 589), (0, 1)):
        order, in_alpha = ORDER_OF[ipal]
        ips, ops = T.psize_of(ipal), T.psize_of(opal)
        src = T.make_packed(rng, w, h, ips)
        exp = np.zeros((h, T.rowstride(w, ops)), np.uint8)
        o.pe_or_rgb_to_yuv888(T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], order, in_alpha, int(opal == 589),
                              cl, T.Q_HIGH)
        lay = packed_layer(eng, ipal, w, h, src)
        assert lb.convert_layer_palette(lay, opal, cl)
        assert lay.yuv_clamping == cl and lay.yuv_subspace == 1
        assert (payload(lay.to_host()[0], w, ops) == payload(exp, w, ops)).all(), (ipal, opal, cl)


def test_rgb_to_packed422_and_planar444(eng):
    """convert_{rgb,bgr,argb}_to_{uyvy,yuyv}_frame :5129-5700 and ..._to_yuvp_frame :5786-6240 through the dispatcher,
    with and without the inline 16-bit gamma LUT"""
    o = T.oracle()
    rng = np.random.default_rng(24)
    for (w, h), ipal, cl in itertools.product(((48, 10), (101, 7), (642, 33)), RGB_PALS, (0, 1)):
        order, in_alpha = ORDER_OF[ipal]
        ips = T.psize_of(ipal)
        src = T.make_packed(rng, w, h, ips)
        for opal, fmt in ((564, 0), (565, 1)):
            we = (w >> 1) << 1
            exp = np.zeros((h, T.rowstride(we // 2, 4)), np.uint8)
            o.pe_or_rgb_to_packed422(fmt, T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], order, in_alpha, cl, T.Q_HIGH, None)
            lay = packed_layer(eng, ipal, w, h, src)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.palette, lay.width, lay.height) == (opal, we, h) and lay.yuv_clamping == cl and lay.yuv_subspace == 1
            assert (payload(lay.to_host()[0], we // 2, 4) == payload(exp, we // 2, 4)).all(), (w, ipal, cl, opal)
        for opal in (544, 545):
            we = (w >> 1) << 1
            ors = T.rowstride(w, 1)
            pl = [np.zeros((h, ors), np.uint8) for _ in range(4)]
            o.pe_or_rgb_to_yuv444p(T.ptr(src), src.strides[0], w, h, T.planes_arg(*pl), ors, order, in_alpha, int(opal == 545), cl, T.Q_HIGH)
            lay = packed_layer(eng, ipal, w, h, src)
            assert lb.convert_layer_palette(lay, opal, cl)
            got = lay.to_host()
            assert len(got) == (4 if opal == 545 else 3)
            for k, g in enumerate(got):
                assert (g[:, :we] == pl[k][:, :we]).all(), (w, ipal, cl, opal, k)
    # gamma: a linear RGB layer to UYVY gets the LINEAR -> sRGB 16-bit LUT inside the converter (colourspace.c:12320-12322)
    w, h = 64, 12
    src = T.make_packed(rng, w, h, 3)
    lut = np.zeros(65536, np.uint16)
    assert o.pe_or_gamma_lut16(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut)) == 0
    exp = np.zeros((h, T.rowstride(w // 2, 4)), np.uint8)
    o.pe_or_rgb_to_packed422(0, T.ptr(src), src.strides[0], w, h, T.ptr(exp), exp.strides[0], 0, 0, 0, T.Q_HIGH, T.ptr(lut))
    lay = packed_layer(eng, 1, w, h, src, gamma_type=T.G_LINEAR)
    assert lb.convert_layer_palette_full(lay, 564, 0, 0, 1, 0)
    assert lay.gamma_type == T.G_SRGB
    assert (payload(lay.to_host()[0], w // 2, 4) == payload(exp, w // 2, 4)).all()


def test_rgb_to_planar420_and_422(eng):
    """convert_{rgb,bgr}_to_yuv420_frame :6250 / :6385 through the dispatcher: 4:2:0 (tables of osubspace) and 4:2:2 (YCbCr),
    padded planes, odd sizes cut to even; ARGB32 (whose reference loop reads G1 / B1 past the pixel, :6357) by its intent"""
    o = T.oracle()
    rng = np.random.default_rng(26)
    for (w, h), ipal, cl, sub in itertools.product(((48, 10), (101, 7), (642, 34), (6, 2)), (1, 2, 3, 4, 5), (0, 1), (1, 2)):
        order, in_alpha = ORDER_OF[ipal]
        src = T.make_packed(rng, w, h, T.psize_of(ipal))
        we, he = w & ~1, h & ~1
        for opal, is422 in ((512, 0), (522, 1)):
            rs = T.rowstride(we, 1)
            ch = he if is422 else he // 2
            pl = [np.zeros((he, rs), np.uint8), np.zeros((ch, rs // 2), np.uint8), np.zeros((ch, rs // 2), np.uint8)]
            strides = (C.c_int * 3)(rs, rs // 2, rs // 2)
            o.pe_or_rgb_to_yuv420p(T.ptr(src), src.strides[0], w, h, T.planes_arg(*pl), strides, order, in_alpha, is422, cl,
                                   1 if is422 else sub, T.Q_HIGH)
            lay = packed_layer(eng, ipal, w, h, src)
            assert lb.convert_layer_palette_full(lay, opal, cl, 0, sub, 0)
            assert (lay.palette, lay.width, lay.height, lay.yuv_clamping) == (opal, we, he, cl)
            got = lay.to_host()
            assert (got[0][:he, :we] == pl[0][:, :we]).all(), (w, h, ipal, cl, sub, opal, "Y")
            assert (got[1][:ch, :we // 2] == pl[1][:, :we // 2]).all(), (w, h, ipal, cl, sub, opal, "U")
            assert (got[2][:ch, :we // 2] == pl[2][:, :we // 2]).all(), (w, h, ipal, cl, sub, opal, "V")


def test_convert_crossfade_fused(eng):
    """BASELINE config 5 per clip as ONE kernel: planar YUV -> RGB24 / BGR24 + 'chroma blend' with the operand ==
    convert_layer_palette followed by simple_blend (oracle), 4:2:2 and 4:2:0, aligned and ragged sizes"""
    o = T.oracle()
    rng = np.random.default_rng(31)
    for (w, h), is422, opal, bf in itertools.product(((256, 64), (130, 34), (642, 362)), (1, 0), (1, 2), (128, 37)):
        y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), True)
        operand = T.make_packed(rng, w, h, 3)
        order = 0 if opal == 1 else 1
        exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
        o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], order, 0, is422, 0, 1,
                               T.Q_HIGH, 1, None)
        o.pe_or_simple_blend(0, opal, T.ptr(exp), exp.strides[0], T.ptr(operand), operand.strides[0], T.ptr(exp), exp.strides[0], w, h, bf,
                             operand.size)
        clip = lb.Layer.from_host(eng, 522 if is422 else 512, w, h, [y, u, v], yuv_subspace=1)
        op_l = packed_layer(eng, opal, w, h, operand)
        before = eng.launch_count
        lb.convert_crossfade(clip, op_l, opal, 0, bf)
        assert eng.launch_count - before == 1
        assert (clip.palette, clip.width, clip.height) == (opal, w, h)
        assert (payload(clip.to_host()[0], w, 3) == payload(exp, w, 3)).all(), (w, h, is422, opal, bf)
    # RGBA32 output is not a crossfade target (the alpha path of simple_blend.c:132 is a different formula): fails loudly
    y, u, v = T.make_yuv_planar(rng, 64, 32, True, True)
    clip = lb.Layer.from_host(eng, 522, 64, 32, [y, u, v], yuv_subspace=1)
    op_l = packed_layer(eng, 3, 64, 32, T.make_packed(rng, 64, 32, 4))
    with pytest.raises(Exception):
        lb.convert_crossfade(clip, op_l, 3, 0, 128)
    assert clip.palette == 522


def test_convert_crossfade_batch_shared_operand(eng):
    """The clips of a multitrack stack against ONE shared operand (BASELINE config 5): pe_fx_convert_crossfade_batch == the per-clip
    call == convert_layer_palette + simple_blend of the oracle; same-shaped clips leave as one launch, odd ones on their own"""
    o = T.oracle()
    rng = np.random.default_rng(33)
    for (w, h), is422, nclips in (((256, 64), 1, 5), ((640, 360), 0, 3), ((1920, 1080), 1, 4), ((132, 34), 1, 2)):
        operand = T.make_packed(rng, w, h, 3)
        op_l = packed_layer(eng, 1, w, h, operand)
        clips, exps = [], []
        for _ in range(nclips):
            y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), True)
            exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
            o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], 0, 0, is422, 0, 1,
                                   T.Q_HIGH, 1, None)
            o.pe_or_simple_blend(0, 1, T.ptr(exp), exp.strides[0], T.ptr(operand), operand.strides[0], T.ptr(exp), exp.strides[0], w, h,
                                 128, operand.size)
            clips.append(lb.Layer.from_host(eng, 522 if is422 else 512, w, h, [y, u, v], yuv_subspace=1))
            exps.append(exp)
        before = eng.launch_count
        assert lb.convert_crossfade_batch(clips, op_l, 1, 0, 128) == nclips
        assert eng.launch_count - before == 1, (w, h, eng.launch_count - before)
        for c, exp in zip(clips, exps):
            assert (c.palette, c.width, c.height) == (1, w, h)
            assert (payload(c.to_host()[0], w, 3) == payload(exp, w, 3)).all(), (w, h, is422)
    # a clip of another size is refused and left alone; the others still convert
    w, h = 128, 32
    operand = T.make_packed(rng, w, h, 3)
    op_l = packed_layer(eng, 1, w, h, operand)
    y, u, v = T.make_yuv_planar(rng, w, h, True, True)
    good = lb.Layer.from_host(eng, 522, w, h, [y, u, v], yuv_subspace=1)
    y2, u2, v2 = T.make_yuv_planar(rng, 64, 32, True, True)
    bad = lb.Layer.from_host(eng, 522, 64, 32, [y2, u2, v2], yuv_subspace=1)
    assert lb.convert_crossfade_batch([good, bad], op_l, 1, 0, 128) == 1
    assert (good.palette, bad.palette, bad.width) == (1, 522, 64)


# ------------------------------------------------------------------------------------------------ YUV <-> YUV family

def _planes444(rng, w, h, n, lo=0, hi=256):
    st = T.rowstride(w, 1)
    out = []
    for _ in range(n):
        a = np.zeros((h, st), np.uint8)
        a[:, :w] = rng.integers(lo, hi, (h, w), dtype=np.uint8)
        out.append(a)
    return out


@pytest.mark.parametrize("size", [(64, 12), (37, 5), (1920, 1080), (3, 2)])
def test_yuv444p_to_rgb_and_packed_yuv(eng, size):
    """YUV444P / YUVA4444P -> the five RGB palettes (convert_yuv_planar_to_*_frame), -> YUV888 / YUVA8888 (combineplanes) and back
    (splitplanes), YUV444P <-> YUVA4444P.  The layers carry subspace YUV (0), what convert_layer_palette asks for: a layer
    tagged YCbCr would be sent through RGB first (colourspace.c:12241-12262)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(80 + w)
    for ipal, cl in itertools.product((544, 545), (0, 1)):
        ia = int(ipal == 545)
        pl = _planes444(rng, w, h, 4 if ia else 3)
        pl4 = pl + ([np.zeros_like(pl[0])] if not ia else [])
        for opal in (1, 2, 3, 4, 5):
            order, oa = {1: (0, 0), 2: (1, 0), 3: (0, 1), 4: (1, 1), 5: (2, 1)}[opal]
            ps = 4 if oa else 3
            exp = np.zeros((h, T.rowstride(w, ps)), np.uint8)
            o.pe_or_yuv444p_to_rgb(T.planes_arg(*pl4), pl[0].strides[0], w, h, T.ptr(exp), exp.strides[0], order, ia, oa, cl, T.Q_HIGH)
            lay = lb.Layer.from_host(eng, ipal, w, h, pl, yuv_clamping=cl, yuv_subspace=0)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.palette, lay.width, lay.height) == (opal, w, h)
            assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), (w, h, ipal, cl, opal)
        for opal in (588, 589):
            oa = int(opal == 589)
            ps = 4 if oa else 3
            exp = np.zeros((h, T.rowstride(w, ps)), np.uint8)
            o.pe_or_combine_planes(T.planes_arg(*pl4), pl[0].strides[0], w, h, T.ptr(exp), exp.strides[0], ia, oa)
            lay = lb.Layer.from_host(eng, ipal, w, h, pl, yuv_clamping=cl, yuv_subspace=0)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.palette, lay.yuv_clamping) == (opal, cl)
            got = lay.to_host()[0]
            assert (payload(got, w, ps) == payload(exp, w, ps)).all(), (w, h, ipal, cl, opal)
            # and back into planes
            for bpal in (544, 545):
                ba = int(bpal == 545)
                st = T.rowstride(w, 1)
                ep = [np.zeros((h, st), np.uint8) for _ in range(4)]
                o.pe_or_split_planes(T.ptr(exp), exp.strides[0], w, h, T.planes_arg(*ep), T.strides_arg(*ep), oa, ba)
                lay2 = packed_layer(eng, opal, w, h, exp, yuv_clamping=cl, yuv_subspace=0)
                assert lb.convert_layer_palette(lay2, bpal, cl)
                g2 = lay2.to_host()
                assert len(g2) == (4 if ba else 3)
                for k, g in enumerate(g2):
                    assert (g[:, :w] == ep[k][:, :w]).all(), (w, h, opal, bpal, k)
        # YUV444P <-> YUVA4444P
        opal = 544 if ia else 545
        lay = lb.Layer.from_host(eng, ipal, w, h, pl, yuv_clamping=cl, yuv_subspace=0)
        assert lb.convert_layer_palette(lay, opal, cl)
        got = lay.to_host()
        assert len(got) == (3 if ia else 4)
        for k in range(3):
            assert (got[k][:, :w] == pl[k][:, :w]).all()
        if not ia:
            assert (got[3][:, :w] == 255).all()


@pytest.mark.parametrize("size", [(64, 12), (38, 6), (1920, 1080), (2, 2)])
def test_yuv420p_yuv422p_chroma_resampling(eng, size):
    """YUV420P -> YUV422P (convert_double_chroma) and YUV422P -> YUV420P (convert_halve_chroma over the whole plane)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(90 + w)
    for cl in (0, 1):
        y, u, v = T.make_yuv_planar(rng, w, h, False, cl == 0)
        st = T.rowstride(w, 1)
        e = [np.zeros((h, st), np.uint8), np.zeros((h, st >> 1), np.uint8), np.zeros((h, st >> 1), np.uint8)]
        o.pe_or_double_chroma(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w >> 1, h >> 1, T.planes_arg(*e), T.strides_arg(*e), cl)
        lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v], yuv_clamping=cl, yuv_subspace=0)
        assert lb.convert_layer_palette(lay, 522, cl)
        assert (lay.palette, lay.width, lay.height, lay.yuv_clamping) == (522, w, h, cl)
        got = lay.to_host()
        assert (got[0][:, :w] == y[:, :w]).all()
        for k in (1, 2):
            assert (got[k][:, :w >> 1] == e[k][:, :w >> 1]).all(), ("double", w, h, cl, k)
        # and down again
        y2, u2, v2 = T.make_yuv_planar(rng, w, h, True, cl == 0)
        e = [np.zeros((h, st), np.uint8), np.zeros((h >> 1, st >> 1), np.uint8), np.zeros((h >> 1, st >> 1), np.uint8)]
        o.pe_or_halve_chroma(T.planes_arg(y2, u2, v2), T.strides_arg(y2, u2, v2), w >> 1, h, T.planes_arg(*e), T.strides_arg(*e), cl)
        lay = lb.Layer.from_host(eng, 522, w, h, [y2, u2, v2], yuv_clamping=cl, yuv_subspace=0)
        assert lb.convert_layer_palette(lay, 512, cl)
        assert (lay.palette, lay.width, lay.height) == (512, w, h)
        got = lay.to_host()
        assert (got[0][:, :w] == y2[:, :w]).all()
        for k in (1, 2):
            assert (got[k][:, :w >> 1] == e[k][:, :w >> 1]).all(), ("halve", w, h, cl, k)


@pytest.mark.parametrize("size", [(64, 6), (34, 5), (1920, 1080), (2, 1)])
def test_packed422_to_planar_packed444_and_swab(eng, size):
    """UYVY / YUYV -> YUV422P (with and without the reference's never-advanced source pointer), -> YUV444P / YUVA4444P,
    -> YUV888 / YUVA8888, UYVY <-> YUYV in place"""
    o = T.oracle()
    w, h = size
    wm = w >> 1
    rng = np.random.default_rng(100 + w)
    e_nq = lb.Engine(ref_quirks=False)
    for ipal in (564, 565):
        fmt = ipal - 564
        src = T.make_packed(rng, wm, h, 4)
        for quirks, en in ((1, eng), (0, e_nq)):
            st = T.rowstride(w, 1)
            ep = [np.zeros((h, st), np.uint8), np.zeros((h, st >> 1), np.uint8), np.zeros((h, st >> 1), np.uint8)]
            o.pe_or_packed422_to_yuv422p(fmt, T.ptr(src), src.strides[0], wm, h, T.planes_arg(*ep), T.strides_arg(*ep), quirks)
            lay = packed_layer(en, ipal, w, h, src, yuv_subspace=0)
            assert lb.convert_layer_palette(lay, 522, 0)
            assert (lay.palette, lay.width, lay.height) == (522, w, h)
            got = lay.to_host()
            assert (got[0][:, :w] == ep[0][:, :w]).all(), ("422p y", w, h, ipal, quirks)
            for k in (1, 2):
                assert (got[k][:, :wm] == ep[k][:, :wm]).all(), ("422p", w, h, ipal, quirks, k)
        for opal in (544, 545):
            aa = int(opal == 545)
            st = T.rowstride(w, 1)
            ep = [np.zeros((h, st), np.uint8) for _ in range(4)]
            o.pe_or_packed422_to_yuv444p(fmt, T.ptr(src), src.strides[0], wm, h, T.planes_arg(*ep), T.strides_arg(*ep), aa)
            lay = packed_layer(eng, ipal, w, h, src, yuv_subspace=0)
            assert lb.convert_layer_palette(lay, opal, 0)
            got = lay.to_host()
            assert len(got) == 3 + aa
            for k, g in enumerate(got):
                assert (g[:, :w] == ep[k][:, :w]).all(), ("444p", w, h, ipal, opal, k)
        for opal in (588, 589):
            aa = int(opal == 589)
            ps = 4 if aa else 3
            exp = np.zeros((h, T.rowstride(w, ps)), np.uint8)
            o.pe_or_packed422_to_yuv888(fmt, T.ptr(src), src.strides[0], wm, h, T.ptr(exp), exp.strides[0], aa)
            lay = packed_layer(eng, ipal, w, h, src, yuv_subspace=0)
            assert lb.convert_layer_palette(lay, opal, 0)
            assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), ("888", w, h, ipal, opal)
        exp = src.copy()
        o.pe_or_swab(T.ptr(exp), exp.strides[0], wm, h)
        lay = packed_layer(eng, ipal, w, h, src, yuv_subspace=0)
        before = eng.launch_count
        assert lb.convert_layer_palette(lay, 565 if ipal == 564 else 564, 0)
        assert eng.launch_count - before == 1
        assert lay.palette == (565 if ipal == 564 else 564)
        assert (payload(lay.to_host()[0], wm, 4) == payload(exp, wm, 4)).all(), ("swab", w, h, ipal)


@pytest.mark.parametrize("size", [(64, 12), (38, 7), (1920, 1080), (2, 2)])
def test_yuv444p_to_packed422_and_yuv420p(eng, size):
    """YUV444P / YUVA4444P -> UYVY / YUYV (convert_yuv_planar_to_{uyvy,yuyv}_frame) and -> YUV420P (convert_yuvp_to_yuv420_frame)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(120 + w)
    for ipal, cl in itertools.product((544, 545), (0, 1)):
        pl = _planes444(rng, w, h, 4 if ipal == 545 else 3)
        for opal in (564, 565):
            wm = w >> 1
            exp = np.zeros((h, T.rowstride(wm, 4)), np.uint8)
            o.pe_or_yuv444p_to_packed422(opal - 564, T.planes_arg(*pl[:3]), pl[0].strides[0], w, h, T.ptr(exp), exp.strides[0], cl)
            lay = lb.Layer.from_host(eng, ipal, w, h, pl, yuv_clamping=cl)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.palette, lay.width, lay.height, lay.yuv_clamping) == (opal, wm * 2, h, cl)
            assert (payload(lay.to_host()[0], wm, 4) == payload(exp, wm, 4)).all(), (w, h, ipal, cl, opal)
        he = h & ~1
        ys = T.rowstride(w & ~1, 1)
        ep = [np.zeros((he, ys), np.uint8), np.zeros((he >> 1, ys >> 1), np.uint8), np.zeros((he >> 1, ys >> 1), np.uint8)]
        o.pe_or_yuv444p_to_yuv420p(T.planes_arg(*pl[:3]), T.strides_arg(*pl[:3]), w & ~1, he, T.planes_arg(*ep), T.strides_arg(*ep), cl)
        lay = lb.Layer.from_host(eng, ipal, w, h, pl, yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, 512, cl)
        assert (lay.palette, lay.width, lay.height) == (512, w & ~1, he)
        got = lay.to_host()
        assert (got[0][:, :w & ~1] == ep[0][:, :w & ~1]).all()
        for k in (1, 2):
            assert (got[k][:, :w >> 1] == ep[k][:, :w >> 1]).all(), ("420p", w, h, ipal, cl, k)


@pytest.mark.parametrize("size", [(64, 12), (38, 6), (1920, 1080), (2, 2)])
def test_planar42x_to_packed422(eng, size):
    """YUV420P / YUV422P -> UYVY / YUYV (convert_yuv420_to_{uyvy,yuyv}_frame, convert_yuv422p_to_{uyvy,yuyv}_frame)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(130 + w)
    for ipal, opal in itertools.product((512, 522), (564, 565)):
        y, u, v = T.make_yuv_planar(rng, w, h, ipal == 522, True)
        wm = w >> 1
        exp = np.zeros((h, T.rowstride(wm, 4)), np.uint8)
        o.pe_or_yuv42xp_to_packed422(opal - 564, T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, int(ipal == 522), T.ptr(exp), exp.strides[0])
        lay = lb.Layer.from_host(eng, ipal, w, h, [y, u, v])
        assert lb.convert_layer_palette(lay, opal, 0)
        assert (lay.palette, lay.width, lay.height) == (opal, w, h)
        assert (payload(lay.to_host()[0], wm, 4) == payload(exp, wm, 4)).all(), (w, h, ipal, opal)


@pytest.mark.parametrize("size", [(64, 12), (38, 6), (1920, 1080), (2, 2), (6, 4)])
def test_yuv420p_to_yuv444p_quad_chroma(eng, size):
    """YUV420P / YVU420P -> YUV444P / YUVA4444P (luma copy + convert_quad_chroma), JPEG and MPEG sampling, both clampings"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(140 + w)
    for ipal, opal, samp, cl in itertools.product((512, 513), (544, 545), (0, 1), (0, 1)):
        y, u, v = T.make_yuv_planar(rng, w, h, False, cl == 0)
        su, sv = (u, v) if ipal == 512 else (v, u)  # YVU420P: plane 1 holds V
        st = T.rowstride(w, 1)
        ep = [np.zeros((h, st), np.uint8) for _ in range(4)]
        o.pe_or_quad_chroma(T.planes_arg(y, su, sv), T.strides_arg(y, su, sv), w, h, T.planes_arg(*ep), st, int(opal == 545), int(samp == 0), cl)
        lay = lb.Layer.from_host(eng, ipal, w, h, [y, u, v], yuv_clamping=cl, yuv_sampling=samp)
        assert lb.convert_layer_palette(lay, opal, cl)
        assert (lay.palette, lay.width, lay.height) == (opal, w, h)
        got = lay.to_host()
        assert len(got) == (4 if opal == 545 else 3)
        assert (got[0][:, :w] == y[:, :w]).all()
        for k in (1, 2):
            assert (got[k][:, :w] == ep[k][:, :w]).all(), ("quad", w, h, ipal, opal, samp, cl, k)
        if opal == 545:
            assert (got[3][:, :w] == 255).all()


@pytest.mark.parametrize("size", [(64, 12), (38, 6), (1920, 1080), (2, 2), (6, 3)])
def test_yuv888_subsample_and_alpha(eng, size):
    """YUV888 / YUVA8888 -> UYVY / YUYV / YUV422P / YUV420P (convert_yuv888_to_*_frame) and YUV888 <-> YUVA8888 (addpost / delpost)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(150 + w)
    for ipal, cl in itertools.product((588, 589), (0, 1)):
        ips = 4 if ipal == 589 else 3
        src = T.make_packed(rng, w, h, ips)
        we = w & ~1
        for opal, mode in ((564, 0), (565, 1), (522, 2), (512, 3)):
            he = h & ~1 if mode == 3 else h
            if he < 1:
                continue
            if mode <= 1:
                ep = [np.zeros((he, T.rowstride(we // 2, 4)), np.uint8)]
            else:
                ys = T.rowstride(we, 1)
                chh = he if mode == 2 else he >> 1
                ep = [np.zeros((he, ys), np.uint8), np.zeros((chh, ys >> 1), np.uint8), np.zeros((chh, ys >> 1), np.uint8)]
            pa = ep + [ep[0]] * (3 - len(ep))
            o.pe_or_yuv888_subsample(mode, T.ptr(src), src.strides[0], we, he, int(ipal == 589), T.planes_arg(*pa), T.strides_arg(*pa), cl)
            lay = packed_layer(eng, ipal, w, h, src, yuv_clamping=cl)
            assert lb.convert_layer_palette(lay, opal, cl)
            assert (lay.palette, lay.width, lay.height) == (opal, we, he)
            got = lay.to_host()
            assert len(got) == len(ep)
            for k, (g, e_) in enumerate(zip(got, ep)):
                nb = (we * 2 if mode <= 1 else (we if k == 0 else we >> 1))
                assert (g[:, :nb] == e_[:, :nb]).all(), ("subsample", w, h, ipal, cl, opal, k)
        # alpha add / drop on the packed YUV bytes
        opal = 589 if ipal == 588 else 588
        lay = packed_layer(eng, ipal, w, h, src, yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, opal, cl)
        got = lay.to_host()[0]
        px = src[:, :w * ips].reshape(h, w, ips)
        ops = 7 - ips
        gp = got[:, :w * ops].reshape(h, w, ops)
        assert (gp[:, :, :3] == px[:, :, :3]).all()
        if ops == 4:
            assert (gp[:, :, 3] == 255).all()


@pytest.mark.parametrize("size", [(64, 12), (34, 6), (1920, 1080), (2, 2)])
def test_packed422_to_yuv420p(eng, size):
    """UYVY / YUYV -> YUV420P (convert_{uyvy,yuyv}_to_yuv420_frame)"""
    o = T.oracle()
    w, h = size
    wm = w >> 1
    rng = np.random.default_rng(160 + w)
    for ipal, cl in itertools.product((564, 565), (0, 1)):
        src = T.make_packed(rng, wm, h, 4)
        ys = T.rowstride(w, 1)
        ep = [np.zeros((h, ys), np.uint8), np.zeros((h >> 1, ys >> 1), np.uint8), np.zeros((h >> 1, ys >> 1), np.uint8)]
        o.pe_or_packed422_to_yuv420p(ipal - 564, T.ptr(src), src.strides[0], wm, h, T.planes_arg(*ep), T.strides_arg(*ep), cl)
        lay = packed_layer(eng, ipal, w, h, src, yuv_clamping=cl)
        assert lb.convert_layer_palette(lay, 512, cl)
        assert (lay.palette, lay.width, lay.height) == (512, w, h)
        got = lay.to_host()
        assert (got[0][:, :w] == ep[0][:, :w]).all()
        for k in (1, 2):
            assert (got[k][:, :wm] == ep[k][:, :wm]).all(), ("packed422 -> 420p", w, h, ipal, cl, k)


@pytest.mark.parametrize("size", [(64, 12), (38, 6), (1920, 1080), (2, 2), (6, 4)])
def test_planar42x_to_packed_yuv888(eng, size):
    """YUV420P / YVU420P / YUV422P -> YUV888 / YUVA8888 (convert_quad_chroma_packed / convert_double_chroma_packed)"""
    o = T.oracle()
    w, h = size
    rng = np.random.default_rng(170 + w)
    for ipal, opal, samp, cl in itertools.product((512, 513, 522), (588, 589), (0, 1), (0, 1)):
        y, u, v = T.make_yuv_planar(rng, w, h, ipal == 522, cl == 0)
        su, sv = (v, u) if ipal == 513 else (u, v)
        ps = 4 if opal == 589 else 3
        exp = np.zeros((h, T.rowstride(w, ps)), np.uint8)
        o.pe_or_chroma_upsample_packed(int(ipal != 522), T.planes_arg(y, su, sv), T.strides_arg(y, su, sv), w, h, T.ptr(exp), exp.strides[0],
                                       int(opal == 589), int(samp == 0), cl)
        lay = lb.Layer.from_host(eng, ipal, w, h, [y, u, v], yuv_clamping=cl, yuv_sampling=samp)
        assert lb.convert_layer_palette(lay, opal, cl)
        assert (lay.palette, lay.width, lay.height) == (opal, w, h)
        assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), ("upsample packed", w, h, ipal, opal, samp, cl)


def test_yuv_clamping_switch(eng):
    """switch_yuv_clamping_and_subspace: convert_layer_palette_full with the same palette / subspace and the other clamping runs
    every sample through the clamped <-> unclamped tables in place; a palette change on top converts afterwards"""
    o = T.oracle()
    rng = np.random.default_rng(110)
    w, h = 66, 10
    ty = [np.zeros(256, np.uint8) for _ in range(4)]
    for k in range(4):
        o.pe_or_yy_table(k, T.ptr(ty[k]))
    for icl in (0, 1):
        ocl = 1 - icl
        tY, tC = (ty[0], ty[1]) if icl == 0 else (ty[2], ty[3])
        # planar
        for pal, is422 in ((512, False), (522, True)):
            y, u, v = T.make_yuv_planar(rng, w, h, is422, False)
            lay = lb.Layer.from_host(eng, pal, w, h, [y, u, v], yuv_clamping=icl, yuv_subspace=1)
            assert lb.convert_layer_palette_full(lay, pal, ocl, 0, 1, 0)
            assert (lay.palette, lay.yuv_clamping) == (pal, ocl)
            got = lay.to_host()
            assert (got[0][:, :w] == tY[y[:, :w]]).all() and (got[1][:, :w >> 1] == tC[u[:, :w >> 1]]).all()
            assert (got[2][:, :w >> 1] == tC[v[:, :w >> 1]]).all()
        pl = _planes444(rng, w, h, 4)
        lay = lb.Layer.from_host(eng, 545, w, h, pl, yuv_clamping=icl, yuv_subspace=1)
        assert lb.convert_layer_palette_full(lay, 545, ocl, 0, 1, 0)
        got = lay.to_host()
        assert (got[0][:, :w] == tY[pl[0][:, :w]]).all() and (got[1][:, :w] == tC[pl[1][:, :w]]).all()
        assert (got[2][:, :w] == tC[pl[2][:, :w]]).all() and (got[3][:, :w] == pl[3][:, :w]).all()
        # packed: the oracle walks the plane densely, the same way the kernel does
        for pal, ps, kind in ((588, 3, 2), (589, 4, 3), (564, 2, 4), (565, 2, 5)):
            src = T.make_packed(rng, w if ps != 2 else w // 2, h, ps if ps != 2 else 4)
            exp = src.copy()
            o.pe_or_switch_clamping_plane(T.ptr(exp), exp.size, kind, int(icl == 0))
            lay = packed_layer(eng, pal, w, h, src, yuv_clamping=icl, yuv_subspace=1)
            assert lb.convert_layer_palette_full(lay, pal, ocl, 0, 1, 0)
            assert lay.yuv_clamping == ocl
            nb = w * ps
            assert (lay.to_host()[0][:, :nb] == exp[:, :nb]).all(), (pal, icl)
        # clamping switch + palette change in one call: YUV888 clamped -> YUV444P unclamped
        # (the reference walks a YUV888 plane densely, Y U V Y U V ... across the row padding: with a rowstride that is not a
        # multiple of 3 the tables are misapplied from row 1 on -- replicated under ref_quirks, per-row phase without)
        src = T.make_packed(rng, w, h, 3)
        sw = src.copy()
        o.pe_or_switch_clamping_plane(T.ptr(sw), sw.size, 2, int(icl == 0))
        lay = packed_layer(eng, 588, w, h, src, yuv_clamping=icl, yuv_subspace=1)
        assert lb.convert_layer_palette_full(lay, 544, ocl, 0, 1, 0)
        assert (lay.palette, lay.yuv_clamping) == (544, ocl)
        got = lay.to_host()
        px = sw[:, :w * 3].reshape(h, w, 3)
        for k in range(3):
            assert (got[k][:, :w] == px[:, :, k]).all(), k
        e_nq = lb.Engine(ref_quirks=False)
        lay = packed_layer(e_nq, 588, w, h, src, yuv_clamping=icl, yuv_subspace=1)
        assert lb.convert_layer_palette_full(lay, 588, ocl, 0, 1, 0)
        px, got = src[:, :w * 3].reshape(h, w, 3), lay.to_host()[0][:, :w * 3].reshape(h, w, 3)
        assert (got[:, :, 0] == tY[px[:, :, 0]]).all() and (got[:, :, 1] == tC[px[:, :, 1]]).all() and (got[:, :, 2] == tC[px[:, :, 2]]).all()
        e_nq.close()


def test_convert_crossfade_batchv_operand_per_clip(eng):
    """pe_fx_convert_crossfade_batchv: clip i against operand i (the frames of one clip against successive frames of the other
    track) in one launch == the per-clip call == the oracle"""
    o = T.oracle()
    rng = np.random.default_rng(34)
    for (w, h), is422, n in (((256, 64), 1, 4), ((640, 360), 0, 3), ((1920, 1080), 1, 2)):
        clips, ops, exps = [], [], []
        for _ in range(n):
            y, u, v = T.make_yuv_planar(rng, w, h, bool(is422), True)
            operand = T.make_packed(rng, w, h, 3)
            exp = np.zeros((h, T.rowstride(w, 3)), np.uint8)
            o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), w, h, T.ptr(exp), exp.strides[0], 0, 0, is422, 0, 1,
                                   T.Q_HIGH, 1, None)
            o.pe_or_simple_blend(0, 1, T.ptr(exp), exp.strides[0], T.ptr(operand), operand.strides[0], T.ptr(exp), exp.strides[0], w, h,
                                 77, operand.size)
            clips.append(lb.Layer.from_host(eng, 522 if is422 else 512, w, h, [y, u, v], yuv_subspace=1))
            ops.append(packed_layer(eng, 1, w, h, operand))
            exps.append(exp)
        before = eng.launch_count
        assert lb.convert_crossfade_batchv(clips, ops, 1, 0, 77) == n
        assert eng.launch_count - before == 1
        for c, exp in zip(clips, exps):
            assert (payload(c.to_host()[0], w, 3) == payload(exp, w, 3)).all(), (w, h, is422)


def test_unhandled_conversion_fails_and_leaves_layer(eng):
    rng = np.random.default_rng(9)
    src = T.make_packed(rng, 32, 8, 4)
    lay = packed_layer(eng, 5, 32, 8, src)  # ARGB32 -> WEED_PALETTE_RGBFLOAT (64, libweed/weed-palettes.h:59): the reference has no converter for it either
    assert not lb.convert_layer_palette(lay, 64, 0)
    assert "not handled" in lb._capi.last_error()
    assert lay.palette == 5
    assert (lay.to_host()[0] == src).all()


# ------------------------------------------------------------------------------------------------ gamma / premult

def test_gamma_convert_layer_and_sub_layer(eng):
    o = T.oracle()
    rng = np.random.default_rng(10)
    w, h = 67, 19
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    for pal in RGB_PALS:
        ps = T.psize_of(pal)
        src = T.make_packed(rng, w, h, ps)
        exp = src.copy()
        o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], pal, 0, 0, w, h, T.ptr(lut))
        lay = packed_layer(eng, pal, w, h, src, gamma_type=T.G_LINEAR)
        assert lb.gamma_convert_layer(T.G_SRGB, lay)
        assert lay.gamma_type == T.G_SRGB
        assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), pal
        # sub rectangle
        exp = src.copy()
        o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], pal, 5, 3, 41, 11, T.ptr(lut))
        lay = packed_layer(eng, pal, w, h, src, gamma_type=T.G_LINEAR)
        assert lb.gamma_convert_sub_layer(T.G_SRGB, 1.0, lay, 5, 3, 41, 11)
        assert (payload(lay.to_host()[0], w, ps) == payload(exp, w, ps)).all(), pal
    # same gamma -> untouched TRUE; YUV -> FALSE (colourspace.c:14076-14080)
    lay = packed_layer(eng, 1, w, h, T.make_packed(rng, w, h, 3), gamma_type=T.G_SRGB)
    assert lb.gamma_convert_layer(T.G_SRGB, lay)
    ylay = packed_layer(eng, 588, w, h, T.make_packed(rng, w, h, 3), gamma_type=T.G_SRGB)
    assert not lb.gamma_convert_layer(T.G_LINEAR, ylay)


def test_alpha_premult(eng):
    o = T.oracle()
    rng = np.random.default_rng(11)
    w, h = 45, 9
    for pal, cl, direction in itertools.product((3, 4, 5, 589), (0, 1), (1, -1)):
        src = T.make_packed(rng, w, h, 4)
        exp = src.copy()
        o.pe_or_alpha_premult(T.ptr(exp), exp.strides[0], pal, cl, w, h, direction)
        lay = packed_layer(eng, pal, w, h, src, yuv_clamping=cl, flags=(1 if direction == -1 else 0))
        lb.alpha_premult(lay, direction)
        assert (payload(lay.to_host()[0], w, 4) == payload(exp, w, 4)).all(), (pal, cl, direction)
        assert lay.flags == (1 if direction == 1 else 0)


def test_alpha_premult_planar_yuva4444p(eng):
    o = T.oracle()
    rng = np.random.default_rng(4445)
    for (w, h), cl, direction in itertools.product(((45, 9), (1920, 270)), (0, 1), (1, -1)):
        st = T.rowstride(w, 1)
        pl = [np.zeros((h, st), np.uint8) for _ in range(4)]
        for p in pl:
            p[:, :w] = rng.integers(0, 256, (h, w), dtype=np.uint8)
        exp = [p.copy() for p in pl]
        o.pe_or_alpha_premult_planar(T.planes_arg(*exp), T.strides_arg(*exp), cl, w, h, direction)
        lay = lb.Layer.from_host(eng, 545, w, h, pl, yuv_clamping=cl)
        lb.alpha_premult(lay, direction)
        got = lay.to_host()
        for k in range(4):
            assert (got[k][:, :w] == exp[k][:, :w]).all(), (w, h, cl, direction, k)


# ------------------------------------------------------------------------------------------------ effects

@pytest.mark.parametrize("size", [(64, 32), (61, 7), (1920, 1080)])
def test_simple_blend_all_filters(eng, size):
    """simple_blend.c: chroma blend + luma overlays on the 5 RGB palettes, separate output and in place"""
    o = T.oracle()
    rng = np.random.default_rng(12)
    w, h = size
    bfs = (0, 1, 100, 128, 255) if w < 1000 else (100,)
    for pal, bf in itertools.product(RGB_PALS, bfs):
        ps = T.psize_of(pal)
        s1 = T.make_packed(rng, w, h, ps)
        s2 = T.make_packed(rng, w, h, ps)
        if ps == 4:
            al = s2[:, 3::4] if pal != 5 else s2[:, 0::4]
            al[rng.random(al.shape) < 0.4] = 255
        l1, l2 = packed_layer(eng, pal, w, h, s1), packed_layer(eng, pal, w, h, s2)
        for typ in (0, 1, 2, 3):
            d0 = np.full_like(s1, 9)
            exp = d0.copy()
            o.pe_or_simple_blend(typ, pal, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, h,
                                 bf, s2.strides[0] * (h - 1) + w * ps)
            lo = packed_layer(eng, pal, w, h, d0)
            lb.simple_blend(typ, l1, l2, lo, bf)
            assert (payload(lo.to_host()[0], w, ps) == payload(exp, w, ps)).all(), (pal, bf, typ)
            lo.free()
        exp = s1.copy()
        o.pe_or_simple_blend(0, pal, T.ptr(exp), exp.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, h, bf,
                             s2.strides[0] * (h - 1) + w * ps)
        lb.simple_blend("chroma blend", l1, l2, l1, bf)  # in place (effects-weed.c:2304-2314)
        assert (payload(l1.to_host()[0], w, ps) == payload(exp, w, ps)).all(), (pal, bf, "inplace")
        l1.free()
        l2.free()


def test_config4_chroma_blend_batch_chain(eng):
    """BASELINE config 4 (reduced batch for the oracle): 1080p RGB24 'chroma blend' bf=100, chain of 3 (64/128/192),
    a batch of independent frames in ONE launch"""
    o = T.oracle()
    n, w, h = 6, 1920, 1080
    ins1, ins2, outs, exps = [], [], [], []
    for i in range(n):
        rng = np.random.default_rng(5 + i)
        s1, s2 = T.make_packed(rng, w, h, 3), T.make_packed(rng, w, h, 3)
        exp = s1.copy()
        for bf in (100, 64, 128, 192):
            o.pe_or_simple_blend(0, 1, T.ptr(exp), exp.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, h, bf,
                                 s2.size)
        exps.append(exp)
        ins1.append(packed_layer(eng, 1, w, h, s1))
        ins2.append(packed_layer(eng, 1, w, h, s2))
    before = eng.launch_count
    for bf in (100, 64, 128, 192):
        lb.simple_blend_batch(0, ins1, ins2, ins1, bf)
    assert eng.launch_count - before == 4
    for i in range(n):
        assert (payload(ins1[i].to_host()[0], w, 3) == payload(exps[i], w, 3)).all(), i


def test_multi_blends(eng):
    o = T.oracle()
    rng = np.random.default_rng(13)
    w, h = 53, 12
    for pal, bf, typ in itertools.product((1, 2), (0, 17, 127, 128, 200, 255), range(7)):
        s1, s2 = T.make_packed(rng, w, h, 3), T.make_packed(rng, w, h, 3)
        exp = np.full_like(s1, 9)
        o.pe_or_multi_blend(typ, pal, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, h, bf)
        l1, l2, lo = packed_layer(eng, pal, w, h, s1), packed_layer(eng, pal, w, h, s2), packed_layer(eng, pal, w, h, np.full_like(s1, 9))
        lb.multi_blend(typ, l1, l2, lo, bf)
        assert (payload(lo.to_host()[0], w, 3) == payload(exp, w, 3)).all(), (pal, bf, typ)


@pytest.mark.parametrize("size", [(64, 32), (61, 7), (1920, 1080), (2, 2)])
def test_slide_over(eng, size):
    """slide_over.c: every direction x moving / fixed clips over the transition range, 3- and 4-byte and macropixel palettes"""
    o = T.oracle()
    rng = np.random.default_rng(14)
    w, h = size
    big = w > 1000
    for pal in ((1, 3) if big else (1, 2, 3, 5, 588, 589, 564, 565)):
        ps = T.psize_of(pal)
        wm = w // 2 if pal in (564, 565) else w  # the plugin's width: macropixels
        s1, s2 = T.make_packed(rng, wm, h, ps, stride=T.rowstride(wm, ps)), T.make_packed(rng, wm, h, ps, stride=T.rowstride(wm, ps))
        l1, l2 = packed_layer(eng, pal, w, h, s1), packed_layer(eng, pal, w, h, s2)
        lo = packed_layer(eng, pal, w, h, np.full_like(s1, 9))
        assert lo.desc.rowstrides[0] == s1.strides[0]
        tvs = (100,) if big else (0, 1, 64, 127, 128, 200, 254, 255)
        for direction, mvl, mvu, tv in itertools.product((1, 2, 3, 4), (0, 1), (0, 1), tvs):
            exp = np.full_like(s1, 9)
            o.pe_or_slide_over(direction, tv, mvl, mvu, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0],
                               wm, h, ps)
            lb.slide_over(l1, l2, lo, tv, direction, mvl, mvu)
            assert (payload(lo.to_host()[0], wm, ps) == payload(exp, wm, ps)).all(), (pal, direction, mvl, mvu, tv)
        with pytest.raises(lb.PixelEngineError):
            lb.slide_over(l1, l2, l1, 10, 1)  # not an in-place filter
        for l in (l1, l2, lo):
            l.free()


def _oracle_compositor(o, pal, w, h, layers, alphas, bg):
    ps = T.psize_of(pal)
    out = np.zeros((h, T.rowstride(w, ps)), np.uint8)
    o.pe_or_fill(T.ptr(out), out.strides[0], pal, w, h, *bg)
    for z in range(len(layers) - 1, -1, -1):
        if alphas[z] > 0:
            o.pe_or_alpha_over(T.ptr(out), out.strides[0], T.ptr(layers[z]), layers[z].strides[0], pal, w, h, alphas[z])
    return out


def test_compositor_alpha_over(eng):
    """compositor.c: bg fill + paint_pixel (double, truncation) per layer, last layer first"""
    o = T.oracle()
    rng = np.random.default_rng(14)
    w, h = 75, 21
    for pal in (1, 2, 3, 4):
        ps = T.psize_of(pal)
        srcs = [T.make_packed(rng, w, h, ps) for _ in range(3)]
        alphas = [0.5, 1.0 / 3.0, 0.9]
        exp = _oracle_compositor(o, pal, w, h, srcs, alphas, (10, 20, 30))
        out = lb.Layer.create(eng, pal, w, h)
        lb.compositor(out, [packed_layer(eng, pal, w, h, s) for s in srcs], alphas, (10, 20, 30))
        assert (payload(out.to_host()[0], w, ps) == payload(exp, w, ps)).all(), pal


def test_config3_4k_alpha_over_gamma(eng):
    """BASELINE config 3: two 3840x2160 RGBA32 layers, scalar alpha 0.5 alpha-over, then gamma (LUT8 on RGB)"""
    o = T.oracle()
    w, h = 3840, 2160
    bg = T.make_packed(np.random.default_rng(3), w, h, 4)
    fg = T.make_packed(np.random.default_rng(4), w, h, 4)
    exp = _oracle_compositor(o, 3, w, h, [fg, bg], [0.5, 1.0], (0, 0, 0))
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], 3, 0, 0, w, h, T.ptr(lut))
    out = lb.Layer.create(eng, 3, w, h, gamma_type=T.G_LINEAR)
    lb.compositor(out, [packed_layer(eng, 3, w, h, fg), packed_layer(eng, 3, w, h, bg)], [0.5, 1.0])
    assert lb.gamma_convert_layer(T.G_SRGB, out)
    got = out.to_host()[0]
    assert (got == exp).all()
    # ... and fused: compositor + gamma in one pass (pe_fx_compositor_gamma), dyadic and non-dyadic alpha
    for alpha in (0.5, 0.3):
        exp2 = _oracle_compositor(o, 3, w, h, [fg, bg], [alpha, 1.0], (0, 0, 0))
        o.pe_or_gamma_apply(T.ptr(exp2), exp2.strides[0], 3, 0, 0, w, h, T.ptr(lut))
        out2 = lb.Layer.create(eng, 3, w, h, gamma_type=T.G_LINEAR)
        before = eng.launch_count
        lb.compositor_gamma(out2, [packed_layer(eng, 3, w, h, fg), packed_layer(eng, 3, w, h, bg)], [alpha, 1.0], T.G_SRGB)
        assert eng.launch_count - before <= 2  # one paint (+ the one-off table build for alpha 0.3)
        assert out2.gamma_type == T.G_SRGB
        assert (out2.to_host()[0] == exp2).all(), alpha
    # the north-star's sRGB -> linear LUT is the identity table in the reference: gamma is then a no-op
    assert (eng.gamma_lut8(1.0, T.G_SRGB, T.G_LINEAR) == np.arange(256)).all()


def test_compositor_gamma_batch(eng):
    """pe_fx_compositor_gamma_batch == pe_fx_compositor_gamma frame by frame == the oracle: two layers (one launch for the batch),
    three layers (paint passes stay ordered), a non-dyadic alpha (table paint, frame-by-frame order), RGB24"""
    o = T.oracle()
    rng = np.random.default_rng(41)
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    for (w, h, pal), alphas, nframes in (((640, 360, 3), [0.5, 1.0], 5), ((130, 34, 3), [0.25, 0.5, 1.0], 4), ((64, 32, 3), [0.3, 1.0], 3),
                                         ((200, 50, 1), [0.75, 1.0], 3), ((64, 16, 3), [1.0], 2)):
        ps = T.psize_of(pal)
        frames = [[T.make_packed(rng, w, h, ps) for _ in alphas] for _ in range(nframes)]
        exps = []
        for ls in frames:
            exp = _oracle_compositor(o, pal, w, h, ls, alphas, (0, 0, 0))
            o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], pal, 0, 0, w, h, T.ptr(lut))
            exps.append(exp)
        outs = [lb.Layer.create(eng, pal, w, h, gamma_type=T.G_LINEAR) for _ in range(nframes)]
        lays = [[packed_layer(eng, pal, w, h, a) for a in ls] for ls in frames]
        before = eng.launch_count
        assert lb.compositor_gamma_batch(outs, lays, alphas, T.G_SRGB) == nframes
        if alphas == [0.5, 1.0]:
            assert eng.launch_count - before == 1
        for out, exp in zip(outs, exps):
            assert out.gamma_type == T.G_SRGB
            assert (payload(out.to_host()[0], w, ps) == payload(exp, w, ps)).all(), (w, h, pal, alphas)


# ------------------------------------------------------------------------------------------------ resize / letterbox

@pytest.mark.parametrize("case", [(64, 48, 32, 24, 3), (64, 48, 96, 72, 4), (130, 50, 77, 34, 4), (1920, 1080, 1280, 720, 3),
                                  (640, 360, 3840, 2160, 4), (100, 100, 100, 50, 3)])
def test_resize_packed_matches_contract(eng, case):
    o = T.oracle()
    rng = np.random.default_rng(15)
    sw, sh, dw, dh, ps = case
    pal = 1 if ps == 3 else 3
    src = T.make_packed(rng, sw, sh, ps)
    exp = np.zeros((dh, T.rowstride(dw, ps)), np.uint8)
    o.pe_or_resize_packed(T.ptr(src), src.strides[0], sw, sh, T.ptr(exp), exp.strides[0], dw, dh, ps)
    lay = packed_layer(eng, pal, sw, sh, src)
    assert lb.resize_layer(lay, dw, dh, lb.LIVES_INTERP_NORMAL, pal, 0)
    assert (lay.width, lay.height) == (dw, dh)
    assert (payload(lay.to_host()[0], dw, ps) == payload(exp, dw, ps)).all()


def test_resize_planar_and_noop(eng):
    o = T.oracle()
    rng = np.random.default_rng(16)
    w, h, dw, dh = 128, 96, 64, 48
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    lay = lb.Layer.from_host(eng, 512, w, h, [y, u, v])
    assert lb.resize_layer(lay, dw, dh, 1, lb.WEED_PALETTE_NONE, 0)
    got = lay.to_host()
    for src, g, (pw, ph, qw, qh) in zip((y, u, v), got, ((w, h, dw, dh), (w // 2, h // 2, dw // 2, dh // 2), (w // 2, h // 2, dw // 2, dh // 2))):
        exp = np.zeros((qh, g.shape[1]), np.uint8)
        o.pe_or_resize_packed(T.ptr(src), src.strides[0], pw, ph, T.ptr(exp), exp.strides[0], qw, qh, 1)
        assert (g[:, :qw] == exp[:, :qw]).all()
    # same size -> TRUE, nothing changes (colourspace.c:14860)
    src = T.make_packed(rng, 64, 48, 3)
    lay = packed_layer(eng, 1, 64, 48, src)
    before = eng.launch_count
    assert lb.resize_layer(lay, 64, 48, 1, 1, 0)
    assert eng.launch_count == before


def test_letterbox_packed_and_planar(eng):
    o = T.oracle()
    rng = np.random.default_rng(17)
    for pal in (1, 3, 5):
        ps = T.psize_of(pal)
        iw, ih, ow, oh = 64, 36, 64, 48
        src = T.make_packed(rng, iw, ih, ps)
        exp = np.zeros((oh, T.rowstride(ow, ps)), np.uint8)
        o.pe_or_letterbox_packed(T.ptr(src), src.strides[0], iw, ih, T.ptr(exp), exp.strides[0], ow, oh, pal)
        lay = packed_layer(eng, pal, iw, ih, src)
        assert lb.letterbox_layer(lay, ow, oh, iw, ih, 1, pal, 0)
        assert (lay.width, lay.height) == (ow, oh)
        assert (payload(lay.to_host()[0], ow, ps) == payload(exp, ow, ps)).all(), pal
    # resize + letterbox (pillarbox): inner 48x48 from 64x36, centred in 80x48
    src = T.make_packed(rng, 64, 36, 4)
    inner = np.zeros((48, T.rowstride(48, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(src), src.strides[0], 64, 36, T.ptr(inner), inner.strides[0], 48, 48, 4)
    exp = np.zeros((48, T.rowstride(81, 4)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], 48, 48, T.ptr(exp), exp.strides[0], 81, 48, 3)
    lay = packed_layer(eng, 3, 64, 36, src)
    assert lb.letterbox_layer(lay, 81, 48, 48, 48, 1, 3, 0)
    assert (payload(lay.to_host()[0], 81, 4) == payload(exp, 81, 4)).all()
    # planar: Y black = 16 (clamped), chroma 128, offsets halved on the chroma planes (:15538-15549)
    y, u, v = T.make_yuv_planar(rng, 64, 32, False, True)
    lay = lb.Layer.from_host(eng, 512, 64, 32, [y, u, v], yuv_clamping=0)
    assert lb.letterbox_layer(lay, 64, 48, 64, 32, 1, 512, 0)
    gy, gu, gv = lay.to_host()
    assert (gy[8:40, :64] == y[:, :64]).all() and (gy[:8, :64] == 16).all() and (gy[40:, :64] == 16).all()
    assert (gu[4:20, :32] == u[:, :32]).all() and (gu[:4, :32] == 128).all() and (gv[20:, :32] == 128).all()


# ------------------------------------------------------------------------------------------------ fused chain

def _oracle_chain(o, y, u, v, fw, fh, is422, bg, ow, oh, iw, ih, alpha, lut8, quirks=1):
    rgba = _oracle_planar(o, y, u, v, fw, fh, 3, is422, 0, 1, T.Q_HIGH, quirks=quirks)
    if (iw, ih) != (fw, fh):
        inner = np.zeros((ih, T.rowstride(iw, 4)), np.uint8)
        o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], fw, fh, T.ptr(inner), inner.strides[0], iw, ih, 4)
    else:
        inner = rgba
    boxed = np.zeros((oh, T.rowstride(ow, 4)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], iw, ih, T.ptr(boxed), boxed.strides[0], ow, oh, 3)
    out = bg.copy()
    o.pe_or_alpha_over(T.ptr(out), out.strides[0], T.ptr(boxed), boxed.strides[0], 3, ow, oh, alpha)
    out[:, 3:ow * 4:4] = 255  # compositor forces the alpha byte (compositor.c:184)
    if lut8 is not None:
        o.pe_or_gamma_apply(T.ptr(out), out.strides[0], 3, 0, 0, ow, oh, T.ptr(lut8))
    return out


@pytest.mark.parametrize("case", [
    # fw, fh, is422, ow, oh, iw, ih
    (64, 48, 0, 64, 48, 64, 48), (64, 48, 0, 128, 96, 128, 72), (128, 96, 0, 64, 64, 64, 48), (130, 34, 1, 200, 80, 150, 40),
    (640, 360, 0, 1280, 720, 1280, 536), (1920, 1080, 0, 1280, 720, 1280, 720),
])
def test_fused_chain_equals_unfused_oracle(eng, case):
    """the fused kernel == convert_layer_palette -> resize -> letterbox -> alpha-over -> gamma run one by one"""
    o = T.oracle()
    rng = np.random.default_rng(18)
    fw, fh, is422, ow, oh, iw, ih = case
    y, u, v = T.make_yuv_planar(rng, fw, fh, bool(is422), True)
    bg = T.make_packed(rng, ow, oh, 4)
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    exp = _oracle_chain(o, y, u, v, fw, fh, is422, bg, ow, oh, iw, ih, 0.5, lut)
    fg_l = lb.Layer.from_host(eng, 522 if is422 else 512, fw, fh, [y, u, v], yuv_subspace=1)
    bg_l = packed_layer(eng, 3, ow, oh, bg, gamma_type=T.G_LINEAR)
    out_l = lb.Layer.create(eng, 3, ow, oh)
    before = eng.launch_count
    lb.fused_convert_letterbox_over_gamma(fg_l, bg_l, out_l, iw, ih, 0.5, T.G_LINEAR, T.G_SRGB)
    got = out_l.to_host()[0]
    assert (payload(got, ow, 4) == payload(exp, ow, 4)).all()
    assert out_l.gamma_type == T.G_SRGB
    # ... and == the engine's own unfused ops
    lay = lb.Layer.from_host(eng, 522 if is422 else 512, fw, fh, [y, u, v], yuv_subspace=1)
    assert lb.convert_layer_palette(lay, 3, 0)
    assert lb.letterbox_layer(lay, ow, oh, iw, ih, 1, 3, 0)
    out2 = lb.Layer.create(eng, 3, ow, oh, gamma_type=T.G_LINEAR)
    lb.compositor(out2, [lay, bg_l], [0.5, 1.0])
    assert lb.gamma_convert_layer(T.G_SRGB, out2)
    assert (payload(out2.to_host()[0], ow, 4) == payload(exp, ow, 4)).all()
    del before


@pytest.mark.parametrize("case", [
    # fw, fh, is422, ow, oh, ih, alpha, gamma(from, to) or None   -- inner width == fw: the fast kernel (pe_kernels_fused2.cu)
    (64, 48, 0, 64, 48, 48, 0.5, (-1, 1)),          # no scaling at all
    (128, 96, 0, 128, 96, 72, 0.5, (-1, 1)),        # vertical squeeze 4:3, letterboxed
    (256, 270, 0, 256, 270, 202, 0.25, None),       # ratio 1.337 (the headline's), no gamma, alpha 1/4
    (128, 64, 0, 140, 100, 100, 0.75, (-1, 1)),     # vertical stretch + pillarbox offset 6 (not a multiple of 4)
    (132, 50, 1, 150, 90, 75, 0.5, (1, 2)),         # 4:2:2, stretch, offset 9
    (64, 34, 0, 64, 40, 30, 1.0, (-1, 1)),          # alpha 1
    (644, 362, 0, 700, 400, 300, 0.1, (-1, 1)),     # non-dyadic alpha: table blend
    (640, 360, 1, 640, 360, 240, 1.0 / 3.0, None),  # 4:2:2, ratio 1.5, table blend, no gamma
    (128, 400, 0, 128, 220, 200, 0.5, (-1, 1)),     # ratio 2: five taps -> generic kernel
    # the register-resident kernel's envelope (pe_kernels_fused3.cu): 4:2:0, inner width == outer width, alpha = k / 256
    (132, 50, 0, 132, 60, 45, 0.5, (-1, 1)),        # partial strip, chroma rows with padding (4:2:0 frames are even, :11603)
    (64, 30, 0, 64, 80, 72, 0.5, (-1, 1)),          # vertical stretch 2.4x: several output rows per step, 2 taps
    (644, 362, 0, 644, 400, 300, 0.5, (-1, 1)),     # six strips, the last one partial; squeeze 1.207
    (1920, 1080, 0, 1920, 1080, 804, 0.5, (-1, 1)), # the headline's ratio at 1080p; chroma stride == chroma width
    (256, 96, 0, 256, 96, 64, 0.375, None),         # ratio 1.5 (four full taps), alpha 3/8, no gamma
    (8, 6, 0, 8, 9, 7, 0.5, (-1, 1)),               # tiny frame: two lanes
    (4, 4, 0, 4, 4, 4, 0.5, (-1, 1)),               # one lane, identity
    (128, 64, 0, 160, 100, 100, 0.5, (-1, 1)),      # pillarbox + letterbox: inner offset 16, border lanes inside inner rows
    (640, 360, 0, 1000, 360, 360, 0.25, (-1, 1)),   # pure pillarbox (offset 180), strips left and right without inner lanes
    # ... and its 4:2:2 instantiation (a step = two luma rows with their own chroma rows; the seed slip :3600 on the lane of column 0)
    (132, 50, 1, 132, 60, 45, 0.5, (-1, 1)),        # partial strip, padded chroma rows
    (64, 30, 1, 64, 80, 72, 0.5, (-1, 1)),          # vertical stretch 2.4x
    (644, 362, 1, 644, 400, 300, 0.5, (-1, 1)),     # six strips, squeeze 1.207
    (1920, 1080, 1, 1920, 1080, 804, 0.5, (-1, 1)), # the headline's ratio at 1080p; chroma stride == chroma width (last rows: slow steps)
    (256, 96, 1, 256, 96, 64, 0.375, None),         # ratio 1.5, alpha 3/8, no gamma
    (8, 6, 1, 8, 9, 7, 0.5, (-1, 1)),               # tiny frame
    (4, 4, 1, 4, 4, 4, 0.5, (-1, 1)),               # one lane, identity
    (128, 64, 1, 160, 100, 100, 0.5, (-1, 1)),      # pillarbox + letterbox: the seed lane is not lane 0 of the strip
    (128, 33, 1, 128, 33, 33, 0.5, (-1, 1)),        # odd height (4:2:2 frames may be odd), identity
    (128, 35, 1, 128, 60, 52, 0.75, (-1, 1)),       # odd height, stretch
])
@pytest.mark.parametrize("variant", ["clamped", "unclamped_noquirks"])
def test_fused_fast_path_cases(case, variant):
    o = T.oracle()
    rng = np.random.default_rng(23)
    fw, fh, is422, ow, oh, ih, alpha, gam = case
    cl = 0 if variant == "clamped" else 1
    quirks = variant == "clamped"
    e = lb.Engine(ref_quirks=quirks)
    y, u, v = T.make_yuv_planar(rng, fw, fh, bool(is422), cl == 0)
    bg = T.make_packed(rng, ow, oh, 4)
    lut = None
    if gam:
        lut = np.zeros(256, np.uint8)
        assert o.pe_or_gamma_lut8(1.0, gam[0], gam[1], 1.4, T.ptr(lut)) == 0
    # oracle chain with this variant's clamping / quirks
    order, add_alpha = ORDER_OF[3]
    rgba = np.zeros((fh, T.rowstride(fw, 4)), np.uint8)
    o.pe_or_yuv420p_to_rgb(T.planes_arg(y, u, v), T.strides_arg(y, u, v), fw, fh, T.ptr(rgba), rgba.strides[0], order, add_alpha,
                           is422, cl, 1, T.Q_HIGH, int(quirks), None)
    inner = np.zeros((ih, T.rowstride(fw, 4)), np.uint8)
    o.pe_or_resize_packed(T.ptr(rgba), rgba.strides[0], fw, fh, T.ptr(inner), inner.strides[0], fw, ih, 4)
    boxed = np.zeros((oh, T.rowstride(ow, 4)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], fw, ih, T.ptr(boxed), boxed.strides[0], ow, oh, 3)
    exp = bg.copy()
    o.pe_or_alpha_over(T.ptr(exp), exp.strides[0], T.ptr(boxed), boxed.strides[0], 3, ow, oh, alpha)
    exp[:, 3:ow * 4:4] = 255
    if lut is not None:
        o.pe_or_gamma_apply(T.ptr(exp), exp.strides[0], 3, 0, 0, ow, oh, T.ptr(lut))
    fg_l = lb.Layer.from_host(e, 522 if is422 else 512, fw, fh, [y, u, v], yuv_clamping=cl, yuv_subspace=1)
    gf, gt = gam if gam else (0, 0)
    bg_l = packed_layer(e, 3, ow, oh, bg, gamma_type=gf)
    out_l = lb.Layer.create(e, 3, ow, oh)
    lb.fused_convert_letterbox_over_gamma(fg_l, bg_l, out_l, fw, ih, alpha, gf, gt)
    got = out_l.to_host()[0]
    bad = np.argwhere(payload(got, ow, 4) != payload(exp, ow, 4))
    assert len(bad) == 0, (len(bad), bad[:5])
    e.close()


def test_fused_random_geometries_equal_unfused_ops(eng):
    """40 random geometries inside k_fused3's envelope (and a few that fall out of it): the fused call == the engine's own
    convert_layer_palette -> letterbox_layer -> compositor -> gamma_convert_layer, which the tests above pin to the oracle"""
    rng = np.random.default_rng(77)
    for it in range(40):
        fw = int(rng.integers(1, 80)) * 4
        fh = int(rng.integers(2, 120)) * 2
        ih = max(4, int(fh * rng.uniform(0.67, 3.0)) & ~1)           # ratio <= 1.5 down (four taps) .. 3x up
        oh = ih + 2 * int(rng.integers(0, 20))
        ow = fw + 8 * int(rng.integers(0, 12)) if it % 3 else fw     # pillarbox offsets: multiples of 4
        alpha = float(rng.choice([0.0, 0.25, 0.5, 0.75, 1.0, 37 / 256]))
        cl = int(rng.integers(0, 2))
        is422 = it % 4 == 1
        ipal = 522 if is422 else 512
        y, u, v = T.make_yuv_planar(rng, fw, fh, is422, cl == 0)
        bg = T.make_packed(rng, ow, oh, 4)
        fg_l = lb.Layer.from_host(eng, ipal, fw, fh, [y, u, v], yuv_clamping=cl, yuv_subspace=1)
        bg_l = packed_layer(eng, 3, ow, oh, bg, gamma_type=T.G_LINEAR)
        out_l = lb.Layer.create(eng, 3, ow, oh)
        lb.fused_convert_letterbox_over_gamma(fg_l, bg_l, out_l, fw, ih, alpha, T.G_LINEAR, T.G_SRGB)
        got = out_l.to_host()[0]
        lay = lb.Layer.from_host(eng, ipal, fw, fh, [y, u, v], yuv_clamping=cl, yuv_subspace=1)
        assert lb.convert_layer_palette(lay, 3, 0)
        assert lb.letterbox_layer(lay, ow, oh, fw, ih, 1, 3, 0)
        out2 = lb.Layer.create(eng, 3, ow, oh, gamma_type=T.G_LINEAR)
        lb.compositor(out2, [lay, bg_l], [alpha, 1.0])
        assert lb.gamma_convert_layer(T.G_SRGB, out2)
        exp = out2.to_host()[0]
        bad = np.argwhere(payload(got, ow, 4) != payload(exp, ow, 4))
        assert len(bad) == 0, (it, fw, fh, is422, ih, ow, oh, alpha, cl, len(bad), bad[:4])
        for l in (fg_l, bg_l, out_l, lay, out2):
            l.free()


def test_fused_headline_4k(eng):
    """north-star headline: 3840x2160 YUV420P fg -> RGBA, letterboxed 3840x1608 inner in a 4K frame, alpha-over a 4K
    RGBA bg, gamma; batch of 2 in one launch"""
    o = T.oracle()
    fw, fh, ow, oh, iw, ih = 3840, 2160, 3840, 2160, 3840, 1608
    lut = np.zeros(256, np.uint8)
    o.pe_or_gamma_lut8(1.0, T.G_LINEAR, T.G_SRGB, 1.4, T.ptr(lut))
    fgs, bgs, outs, exps = [], [], [], []
    for i in range(2):
        y, u, v = T.make_yuv_planar(np.random.default_rng(20 + 2 * i), fw, fh, False, True)
        bg = T.make_packed(np.random.default_rng(21 + 2 * i), ow, oh, 4)
        exps.append(_oracle_chain(o, y, u, v, fw, fh, 0, bg, ow, oh, iw, ih, 0.5, lut))
        fgs.append(lb.Layer.from_host(eng, 512, fw, fh, [y, u, v], yuv_subspace=1))
        bgs.append(packed_layer(eng, 3, ow, oh, bg, gamma_type=T.G_LINEAR))
        outs.append(lb.Layer.create(eng, 3, ow, oh))
    before = eng.launch_count
    lb.fused_convert_letterbox_over_gamma_batch(fgs, bgs, outs, iw, ih, 0.5, T.G_LINEAR, T.G_SRGB)
    assert eng.launch_count - before <= 2  # the fused kernel (+ the one-off [bg][fg] table build)
    for i in range(2):
        assert (outs[i].to_host()[0] == exps[i]).all(), i


# ------------------------------------------------------------------------------------------------ diagnostics / host path

def test_frame_stats(eng):
    rng = np.random.default_rng(19)
    w, h = 333, 77
    for pal in (1, 3, 5):
        ps = T.psize_of(pal)
        src = T.make_packed(rng, w, h, ps, lo=3, hi=250)
        st = packed_layer(eng, pal, w, h, src).stats()
        px = src[:, :w * ps].reshape(h, w, ps)
        a_off = {1: -1, 3: 3, 5: 0}[pal]
        for k in range(ps):
            assert st["min"][k] == px[..., k].min() and st["max"][k] == px[..., k].max()
        col = [k for k in range(ps) if k != a_off]
        assert (st["hist"] == np.bincount(px[..., col].reshape(-1), minlength=256)).all()
        assert st["sum"] == int(px.astype(np.uint64).sum())
        assert not st["all_black_ish"] and not st["all_black"]
    blk = lb.Layer.create(eng, 3, 64, 64, black_fill=True)
    assert blk.stats()["all_black_ish"] == 1 and blk.stats()["all_black"] == 1


def test_host_dropins(eng):
    """pe_host_*: host buffers in, host buffers out (H2D and D2H inside the call)"""
    o = T.oracle()
    rng = np.random.default_rng(20)
    w, h = 96, 40
    y, u, v = T.make_yuv_planar(rng, w, h, False, True)
    exp = _oracle_planar(o, y, u, v, w, h, 3, 0, 0, 1, T.Q_HIGH)
    hl = lb.HostLayer(512, w, h, [y.copy(), u.copy(), v.copy()], yuv_subspace=1)
    assert lb.host_convert_layer_palette_full(eng, hl, 3, 0, 0, 1, 0)
    assert hl.d.palette == 3 and hl.d.rowstrides[0] == T.rowstride(w, 4)
    assert (payload(hl.planes[0], w, 4) == payload(exp, w, 4)).all()
    # RGB24 -> BGR24: same byte size, new palette
    src = T.make_packed(rng, w, h, 3)
    hl = lb.HostLayer(1, w, h, [src.copy()])
    assert lb.host_convert_layer_palette_full(eng, hl, 2, 0, 0, 0, 0)
    assert (hl.planes[0][:, 0:w * 3:3] == src[:, 2:w * 3:3]).all()
    # blend
    s1, s2 = T.make_packed(rng, w, h, 3), T.make_packed(rng, w, h, 3)
    exp = np.zeros_like(s1)
    o.pe_or_simple_blend(0, 1, T.ptr(s1), s1.strides[0], T.ptr(s2), s2.strides[0], T.ptr(exp), exp.strides[0], w, h, 100, s2.size)
    d = np.zeros_like(s1)
    lb.host_simple_blend(eng, 0, lb.HostLayer(1, w, h, [s1]), lb.HostLayer(1, w, h, [s2]), lb.HostLayer(1, w, h, [d]), 100)
    assert (payload(d, w, 3) == payload(exp, w, 3)).all()
    # resize + letterbox
    hl = lb.HostLayer(1, w, h, [s1.copy()])
    assert lb.host_letterbox_layer(eng, hl, 128, 64, 128, 52, 1, 1, 0)
    assert (hl.d.width, hl.d.height) == (128, 64)
    inner = np.zeros((52, T.rowstride(128, 3)), np.uint8)
    o.pe_or_resize_packed(T.ptr(s1), s1.strides[0], w, h, T.ptr(inner), inner.strides[0], 128, 52, 3)
    box = np.zeros((64, T.rowstride(128, 3)), np.uint8)
    o.pe_or_letterbox_packed(T.ptr(inner), inner.strides[0], 128, 52, T.ptr(box), box.strides[0], 128, 64, 1)
    assert (payload(hl.planes[0], 128, 3) == payload(box, 128, 3)).all()


def test_host_fused_batch_pipeline(eng):
    """pe_host_fused_convert_letterbox_over_gamma_batch: 7 host frames through the 3-slot H2D / kernel / D2H pipeline ==
    the same frames one by one"""
    rng = np.random.default_rng(21)
    fw, fh, ow, oh, ih = 128, 96, 128, 96, 72
    fgs, bgs, outs, exps = [], [], [], []
    for i in range(7):
        y, u, v = T.make_yuv_planar(rng, fw, fh, False, True)
        bg = T.make_packed(rng, ow, oh, 4)
        f = lb.HostLayer(512, fw, fh, [y, u, v], yuv_subspace=1)
        b = lb.HostLayer(3, ow, oh, [bg], gamma_type=T.G_LINEAR)
        one = lb.HostLayer(3, ow, oh, [np.zeros_like(bg)])
        lb.host_fused_convert_letterbox_over_gamma(eng, f, b, one, fw, ih, 0.5, T.G_LINEAR, T.G_SRGB)
        exps.append(one.planes[0].copy())
        fgs.append(f)
        bgs.append(b)
        outs.append(lb.HostLayer(3, ow, oh, [np.zeros_like(bg)]))
    lb.host_fused_convert_letterbox_over_gamma_batch(eng, fgs, bgs, outs, fw, ih, 0.5, T.G_LINEAR, T.G_SRGB)
    for i in range(7):
        assert (outs[i].planes[0] == exps[i]).all(), i
        assert outs[i].d.gamma_type == T.G_SRGB
